// lbp.cu -- gs_lbp_window / gs_lbp_detect (reference grayskull.h:769-835).
//
// A boosted cascade of 3x3 multi-block LBP features evaluated on an integral image, scanned over
// a fp32 scale ladder.  Not HBM-bound (about 1.1 B of compulsory traffic per window, but >= 6
// weak classifiers x 16 integral-image gathers): throughput is decided by gather rate and
// divergence.  Round-1 structure:
//   host    : the scale ladder (repeated fp32 multiply, reference :819), per-scale window size and
//             the per-scale truncated feature geometry (int)(int8 * scale) (:799-804) are computed
//             on the host with the reference's exact fp32 operations and cached on the device,
//             together with the cascade tables (keyed by content hash);
//   k_lbp_scan   : one lane per window, 32 x-adjacent windows per warp ("slot").  Each weak reads
//             the 4x4 corner lattice of its 3x3 cells (16 loads instead of the reference's 36),
//             stage sums are accumulated with sequential fp32 adds exactly as :808, a lane drops
//             out at the first failed stage and the warp leaves the cascade when its ballot is
//             empty.  The surviving-lane ballot of every slot is stored;
//   k_row_scan   : per-frame exclusive scan of the per-CTA hit counts;
//   k_lbp_emit   : rects written in the reference's (scale, y, x) order, truncated at max_rects
//             (the reference stops scanning there, :819-823).
#include <mutex>
#include <string.h>
#include <vector>

#include "common.cuh"
#include "scan.cuh"

namespace gsb {

struct ScaleInfo {
  int win_w, win_h, nx, ny;
  unsigned chunks;             // 32-window slots per scan row
  unsigned feat_off;           // first entry of this scale in the feature table
  unsigned long long slot0;    // first slot of this scale
};
struct Weak {
  float left, right;
  uint16_t fidx, sub_off, nsub, pad;
};
struct Stage {
  float thr;
  uint16_t start, n;
};

struct DevCascade {            // pointers into one device blob
  const ScaleInfo *scales;
  const short4 *feat;          // [nscales][nfeatures] = (fx, fy, fw, fh) after scaling/clamping
  const Weak *weaks;
  const int *subsets;
  const Stage *stages;
  int nscales, nfeatures, nstages;
  unsigned long long total_slots, total_windows;
  int step;
  bool safe_geometry;          // every feature stays inside its window at every scale
};

template <bool GUARD>
__device__ __forceinline__ uint32_t corner(const uint32_t *__restrict__ ii, unsigned iw, unsigned ih, int x, int y) {
  if (x < 0 || y < 0) return 0u;   // gs_integral_sum's x == 0 / y == 0 guards (reference :758-760)
  if (GUARD && ((unsigned)x >= iw || (unsigned)y >= ih)) return 0u;
  return __ldg(ii + (size_t)y * iw + (unsigned)x);
}

// gs_lbp_code (reference :769-783) from the 4x4 corner lattice
template <bool GUARD>
__device__ __forceinline__ int lbp_code(const uint32_t *__restrict__ ii, unsigned iw, unsigned ih, int x0, int y0,
                                        int fw, int fh) {
  uint32_t g[4][4];
#pragma unroll
  for (int j = 0; j < 4; j++)
#pragma unroll
    for (int i = 0; i < 4; i++) g[j][i] = corner<GUARD>(ii, iw, ih, x0 - 1 + i * fw, y0 - 1 + j * fh);
  uint32_t c[3][3];
#pragma unroll
  for (int j = 0; j < 3; j++)
#pragma unroll
    for (int i = 0; i < 3; i++) c[j][i] = g[j + 1][i + 1] + g[j][i] - g[j][i + 1] - g[j + 1][i];
  const uint32_t m = c[1][1];
  return ((c[0][0] >= m) << 7) | ((c[0][1] >= m) << 6) | ((c[0][2] >= m) << 5) | ((c[1][2] >= m) << 4) |
         ((c[2][2] >= m) << 3) | ((c[2][1] >= m) << 2) | ((c[2][0] >= m) << 1) | ((c[1][0] >= m) << 0);
}

// the cascade for one window per lane (reference gs_lbp_window :794-812); warp-collective
template <bool GUARD>
__device__ __forceinline__ bool cascade_eval(const uint32_t *__restrict__ ii, unsigned iw, unsigned ih, int x, int y,
                                             const short4 *__restrict__ feat, const Weak *__restrict__ weaks,
                                             const int *__restrict__ subsets, const Stage *__restrict__ stages,
                                             int nstages, bool alive) {
  for (int si = 0; si < nstages; si++) {
    if (!__any_sync(0xFFFFFFFFu, alive)) break;
    const Stage st = stages[si];
    if (alive) {
      float sum = 0.0f;
      for (int i = 0; i < st.n; i++) {
        const Weak wk = weaks[st.start + i];
        const short4 f = feat[wk.fidx];
        const int code = lbp_code<GUARD>(ii, iw, ih, x + f.x, y + f.y, f.z, f.w);
        const int idx = code >> 5;
        const bool match = idx < (int)wk.nsub && ((((unsigned)__ldg(subsets + wk.sub_off + idx)) >> (code & 31)) & 1u);
        sum = __fadd_rn(sum, match ? wk.left : wk.right);
      }
      if (sum < st.thr) alive = false;
    }
  }
  return alive;
}

template <bool GUARD>
__global__ void __launch_bounds__(256)
k_lbp_scan(const uint32_t *__restrict__ ii_all, unsigned iw, unsigned ih, DevCascade dc, unsigned *__restrict__ masks,
           unsigned *__restrict__ blockcount) {
  __shared__ unsigned wcnt[8];
  const unsigned f = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const unsigned long long slot = (unsigned long long)blockIdx.x * 8 + warp;
  const uint32_t *ii = ii_all + (size_t)f * iw * ih;
  unsigned mask = 0;
  if (slot < dc.total_slots) {
    int si = 0;
    while (si + 1 < dc.nscales && slot >= dc.scales[si + 1].slot0) si++;
    const ScaleInfo sc = dc.scales[si];
    const unsigned long long rel = slot - sc.slot0;
    const unsigned yi = (unsigned)(rel / sc.chunks), xi = (unsigned)(rel % sc.chunks) * 32 + lane;
    const bool valid = xi < (unsigned)sc.nx;
    const bool hit = cascade_eval<GUARD>(ii, iw, ih, (int)xi * dc.step, (int)yi * dc.step, dc.feat + sc.feat_off,
                                         dc.weaks, dc.subsets, dc.stages, dc.nstages, valid);
    mask = __ballot_sync(0xFFFFFFFFu, hit);
    if (lane == 0) masks[(size_t)f * dc.total_slots + slot] = mask;
  }
  if (lane == 0) wcnt[warp] = __popc(mask);
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned t = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) t += wcnt[i];
    blockcount[(size_t)f * gridDim.x + blockIdx.x] = t;
  }
}

__global__ void __launch_bounds__(256)
k_lbp_emit(DevCascade dc, const unsigned *__restrict__ masks, const unsigned *__restrict__ blockoff,
           struct gs_rect *__restrict__ rects, unsigned max_rects) {
  const unsigned f = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const unsigned off = blockoff[(size_t)f * gridDim.x + blockIdx.x];
  if (off >= max_rects) return;
  const unsigned long long slot0 = (unsigned long long)blockIdx.x * 8;
  unsigned before = 0, mine = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const unsigned long long s = slot0 + i;
    const unsigned m = s < dc.total_slots ? masks[(size_t)f * dc.total_slots + s] : 0u;
    if (i < (int)warp) before += __popc(m);
    if (i == (int)warp) mine = m;
  }
  if (!((mine >> lane) & 1u)) return;
  const unsigned pos = off + before + __popc(mine & ((1u << lane) - 1u));
  if (pos >= max_rects) return;
  const unsigned long long slot = slot0 + warp;
  int si = 0;
  while (si + 1 < dc.nscales && slot >= dc.scales[si + 1].slot0) si++;
  const ScaleInfo sc = dc.scales[si];
  const unsigned long long rel = slot - sc.slot0;
  const unsigned yi = (unsigned)(rel / sc.chunks), xi = (unsigned)(rel % sc.chunks) * 32 + lane;
  struct gs_rect r;
  r.x = xi * dc.step, r.y = yi * dc.step, r.w = (unsigned)sc.win_w, r.h = (unsigned)sc.win_h;
  *reinterpret_cast<uint4 *>(rects + (size_t)f * max_rects + pos) = make_uint4(r.x, r.y, r.w, r.h);
}

// gs_lbp_window for a single window: all 32 lanes evaluate the same window
__global__ void k_lbp_window_one(const uint32_t *ii, unsigned iw, unsigned ih, int x, int y, const short4 *feat,
                                 const Weak *weaks, const int *subsets, const Stage *stages, int nstages,
                                 unsigned *out) {
  const bool hit = cascade_eval<true>(ii, iw, ih, x, y, feat, weaks, subsets, stages, nstages, true);
  if (threadIdx.x == 0) *out = hit ? 1u : 0u;
}

// ---- host: plan + device cache ---------------------------------------------------------------
struct PlanKey {
  unsigned long long hash;
  unsigned iw, ih;
  float sf, mn, mx;
  int step, device;
  bool operator==(const PlanKey &o) const {
    return hash == o.hash && iw == o.iw && ih == o.ih && sf == o.sf && mn == o.mn && mx == o.mx && step == o.step &&
           device == o.device;
  }
};
struct PlanEntry {
  PlanKey key;
  void *blob;
  DevCascade dc;
  std::vector<ScaleInfo> scales;
};
static std::mutex g_plan_mutex;
static std::vector<PlanEntry> g_plans;

static unsigned long long fnv(unsigned long long h, const void *p, size_t n) {
  const unsigned char *b = static_cast<const unsigned char *>(p);
  for (size_t i = 0; i < n; i++) h = (h ^ b[i]) * 1099511628211ull;
  return h;
}

static unsigned subsets_len(const struct gs_lbp_cascade *c) {
  unsigned n = 0;
  for (unsigned i = 0; i < c->nweaks; i++) {
    unsigned e = (unsigned)c->weak_subset_offset[i] + c->weak_num_subsets[i];
    if (e > n) n = e;
  }
  return n;
}

static unsigned long long cascade_hash(const struct gs_lbp_cascade *c) {
  unsigned long long h = 1469598103934665603ull;
  h = fnv(h, &c->window_w, 2), h = fnv(h, &c->window_h, 2);
  h = fnv(h, &c->nfeatures, 2), h = fnv(h, &c->nweaks, 2), h = fnv(h, &c->nstages, 2);
  h = fnv(h, c->features, (size_t)c->nfeatures * 4);
  h = fnv(h, c->weak_feature_idx, (size_t)c->nweaks * 2);
  h = fnv(h, c->weak_left_val, (size_t)c->nweaks * 4), h = fnv(h, c->weak_right_val, (size_t)c->nweaks * 4);
  h = fnv(h, c->weak_subset_offset, (size_t)c->nweaks * 2), h = fnv(h, c->weak_num_subsets, (size_t)c->nweaks * 2);
  h = fnv(h, c->subsets, (size_t)subsets_len(c) * 4);
  h = fnv(h, c->stage_weak_start, (size_t)c->nstages * 2), h = fnv(h, c->stage_nweaks, (size_t)c->nstages * 2);
  h = fnv(h, c->stage_threshold, (size_t)c->nstages * 4);
  return h;
}

// the reference's scale ladder (:819-821), one entry per scale it would scan
static void build_scales(const struct gs_lbp_cascade *c, unsigned iw, unsigned ih, float scale_factor, float min_scale,
                         float max_scale, int step, std::vector<float> &scale_vals, std::vector<ScaleInfo> &out) {
  unsigned long long slot = 0;
  for (volatile float scale = min_scale; scale <= max_scale; scale = scale * scale_factor) {
    const float s = scale;
    const int win_w = (int)((float)c->window_w * s), win_h = (int)((float)c->window_h * s);
    if (win_w > (int)iw || win_h > (int)ih) break;
    ScaleInfo si;
    si.win_w = win_w, si.win_h = win_h;
    si.nx = ((int)iw - win_w) / step + 1, si.ny = ((int)ih - win_h) / step + 1;
    si.chunks = (unsigned)(si.nx + 31) / 32;
    si.feat_off = (unsigned)out.size() * c->nfeatures;
    si.slot0 = slot;
    slot += (unsigned long long)si.chunks * si.ny;
    out.push_back(si);
    scale_vals.push_back(s);
    if (out.size() > 4096 || !(scale_factor > 1.0f)) break;  // a non-growing ladder never ends in the reference
  }
}

static PlanEntry *get_plan(const struct gs_lbp_cascade *c, unsigned iw, unsigned ih, float sf, float mn, float mx,
                           int step) {
  int dev = 0;
  cudaGetDevice(&dev);
  PlanKey key = {cascade_hash(c), iw, ih, sf, mn, mx, step, dev};
  std::lock_guard<std::mutex> lock(g_plan_mutex);
  for (auto &e : g_plans)
    if (e.key == key) return &e;

  PlanEntry e;
  e.key = key;
  std::vector<float> svals;
  build_scales(c, iw, ih, sf, mn, mx, step, svals, e.scales);
  const int ns = (int)e.scales.size(), nf = c->nfeatures, nw = c->nweaks, nst = c->nstages;
  const unsigned nsub = subsets_len(c);
  std::vector<short4> feat((size_t)ns * nf);
  bool safe = true;
  for (int s = 0; s < ns; s++)
    for (int i = 0; i < nf; i++) {
      const float sc = svals[s];
      int fx = (int)((float)c->features[i * 4 + 0] * sc), fy = (int)((float)c->features[i * 4 + 1] * sc);
      int fw = (int)((float)c->features[i * 4 + 2] * sc), fh = (int)((float)c->features[i * 4 + 3] * sc);
      if (fw < 1) fw = 1;
      if (fh < 1) fh = 1;
      feat[(size_t)s * nf + i] = make_short4((short)fx, (short)fy, (short)fw, (short)fh);
      if (fx < 0 || fy < 0 || fx + 3 * fw > e.scales[s].win_w || fy + 3 * fh > e.scales[s].win_h) safe = false;
    }
  std::vector<Weak> weaks(nw);
  for (int i = 0; i < nw; i++) {
    weaks[i].left = c->weak_left_val[i], weaks[i].right = c->weak_right_val[i];
    weaks[i].fidx = c->weak_feature_idx[i], weaks[i].sub_off = c->weak_subset_offset[i];
    weaks[i].nsub = c->weak_num_subsets[i], weaks[i].pad = 0;
  }
  std::vector<Stage> stages(nst);
  for (int i = 0; i < nst; i++) {
    stages[i].thr = c->stage_threshold[i];
    stages[i].start = c->stage_weak_start[i], stages[i].n = c->stage_nweaks[i];
  }
  auto align16 = [](size_t v) { return (v + 15) & ~(size_t)15; };
  const size_t o_sc = 0, o_ft = align16(o_sc + sizeof(ScaleInfo) * (ns ? ns : 1));
  const size_t o_wk = align16(o_ft + sizeof(short4) * feat.size()), o_sb = align16(o_wk + sizeof(Weak) * nw);
  const size_t o_st = align16(o_sb + 4 * (size_t)nsub), total = align16(o_st + sizeof(Stage) * nst);
  std::vector<unsigned char> host(total, 0);
  if (ns) memcpy(&host[o_sc], e.scales.data(), sizeof(ScaleInfo) * ns);
  if (!feat.empty()) memcpy(&host[o_ft], feat.data(), sizeof(short4) * feat.size());
  memcpy(&host[o_wk], weaks.data(), sizeof(Weak) * nw);
  memcpy(&host[o_sb], c->subsets, 4 * (size_t)nsub);
  memcpy(&host[o_st], stages.data(), sizeof(Stage) * nst);
  if (cudaMalloc(&e.blob, total) != cudaSuccess) return nullptr;
  if (cudaMemcpy(e.blob, host.data(), total, cudaMemcpyHostToDevice) != cudaSuccess) return nullptr;
  unsigned char *b = static_cast<unsigned char *>(e.blob);
  e.dc.scales = reinterpret_cast<const ScaleInfo *>(b + o_sc);
  e.dc.feat = reinterpret_cast<const short4 *>(b + o_ft);
  e.dc.weaks = reinterpret_cast<const Weak *>(b + o_wk);
  e.dc.subsets = reinterpret_cast<const int *>(b + o_sb);
  e.dc.stages = reinterpret_cast<const Stage *>(b + o_st);
  e.dc.nscales = ns, e.dc.nfeatures = nf, e.dc.nstages = nst, e.dc.step = step;
  e.dc.safe_geometry = safe;
  e.dc.total_slots = 0, e.dc.total_windows = 0;
  for (auto &s : e.scales) {
    e.dc.total_slots += (unsigned long long)s.chunks * s.ny;
    e.dc.total_windows += (unsigned long long)s.nx * s.ny;
  }
  if (g_plans.size() >= 16) {  // tiny cache: drop the oldest
    cudaFree(g_plans.front().blob);
    g_plans.erase(g_plans.begin());
  }
  g_plans.push_back(e);
  return &g_plans.back();
}

}  // namespace gsb

extern "C" {

unsigned long long gs_b200_lbp_window_count(const struct gs_lbp_cascade *c, unsigned iw, unsigned ih,
                                            float scale_factor, float min_scale, float max_scale, int step) {
  GSB_ASSERT(c && step > 0);
  std::vector<float> sv;
  std::vector<gsb::ScaleInfo> sc;
  gsb::build_scales(c, iw, ih, scale_factor, min_scale, max_scale, step, sv, sc);
  unsigned long long n = 0;
  for (auto &s : sc) n += (unsigned long long)s.nx * s.ny;
  return n;
}

int gs_b200_lbp_detect_batch(const struct gs_lbp_cascade *c, const uint32_t *ii, unsigned iw, unsigned ih,
                             unsigned n, struct gs_rect *rects, unsigned *counts, unsigned max_rects,
                             float scale_factor, float min_scale, float max_scale, int step, gs_b200_stream s) {
  GSB_ASSERT(c && ii && iw > 0 && ih > 0 && counts && step > 0);  // step <= 0 never terminates in the reference
  GSB_ASSERT(rects || max_rects == 0);
  if (n == 0) return 0;
  cudaStream_t st = static_cast<cudaStream_t>(s);
  gsb::PlanEntry *p = gsb::get_plan(c, iw, ih, scale_factor, min_scale, max_scale, step);
  if (!p) return gsb::record_error(cudaErrorMemoryAllocation, __FILE__, __LINE__);
  const gsb::DevCascade &dc = p->dc;
  if (dc.total_slots == 0 || max_rects == 0) {
    GSB_CHECK(cudaMemsetAsync(counts, 0, sizeof(unsigned) * n, st));
    return 0;
  }
  const unsigned long long nblocks = (dc.total_slots + 7) / 8;
  GSB_ASSERT(nblocks < 0x7FFFFFFFull && n <= 65535u);
  unsigned *masks = static_cast<unsigned *>(gsb::workspace(st, gsb::WS_LBP_A, 4 * (size_t)dc.total_slots * n));
  unsigned *bcount = static_cast<unsigned *>(gsb::workspace(st, gsb::WS_LBP_B, 4 * (size_t)nblocks * n));
  if (!masks || !bcount) return (int)cudaErrorMemoryAllocation;
  dim3 grid((unsigned)nblocks, n);
  if (dc.safe_geometry) gsb::k_lbp_scan<false><<<grid, 256, 0, st>>>(ii, iw, ih, dc, masks, bcount);
  else gsb::k_lbp_scan<true><<<grid, 256, 0, st>>>(ii, iw, ih, dc, masks, bcount);
  GSB_LAUNCHED(1);
  gsb::k_row_scan<<<n, 1024, 0, st>>>(bcount, (unsigned)nblocks, counts, max_rects);
  GSB_LAUNCHED(1);
  gsb::k_lbp_emit<<<grid, 256, 0, st>>>(dc, masks, bcount, rects, max_rects);
  GSB_LAUNCHED(1);
  return 0;
}

// single-window hook for api.cu (gs_lbp_window): scale-specific geometry is built on the fly
int gsb_lbp_window_single(const struct gs_lbp_cascade *c, const uint32_t *ii, unsigned iw, unsigned ih, int x, int y,
                          float scale, unsigned *out_dev, cudaStream_t s) {
  const int nf = c->nfeatures, nw = c->nweaks, nst = c->nstages;
  const unsigned nsub = gsb::subsets_len(c);
  std::vector<short4> feat(nf);
  for (int i = 0; i < nf; i++) {
    int fx = (int)((float)c->features[i * 4 + 0] * scale), fy = (int)((float)c->features[i * 4 + 1] * scale);
    int fw = (int)((float)c->features[i * 4 + 2] * scale), fh = (int)((float)c->features[i * 4 + 3] * scale);
    feat[i] = make_short4((short)fx, (short)fy, (short)(fw < 1 ? 1 : fw), (short)(fh < 1 ? 1 : fh));
  }
  std::vector<gsb::Weak> weaks(nw);
  for (int i = 0; i < nw; i++) {
    weaks[i].left = c->weak_left_val[i], weaks[i].right = c->weak_right_val[i];
    weaks[i].fidx = c->weak_feature_idx[i], weaks[i].sub_off = c->weak_subset_offset[i];
    weaks[i].nsub = c->weak_num_subsets[i], weaks[i].pad = 0;
  }
  std::vector<gsb::Stage> stages(nst);
  for (int i = 0; i < nst; i++) {
    stages[i].thr = c->stage_threshold[i];
    stages[i].start = c->stage_weak_start[i], stages[i].n = c->stage_nweaks[i];
  }
  const size_t o_ft = 0, o_wk = (sizeof(short4) * nf + 15) & ~(size_t)15;
  const size_t o_sb = (o_wk + sizeof(gsb::Weak) * nw + 15) & ~(size_t)15;
  const size_t o_st = (o_sb + 4 * (size_t)nsub + 15) & ~(size_t)15, total = o_st + sizeof(gsb::Stage) * nst;
  unsigned char *blob = static_cast<unsigned char *>(gsb::workspace(s, gsb::WS_LBP_C, total));
  if (!blob) return (int)cudaErrorMemoryAllocation;
  GSB_CHECK(cudaMemcpyAsync(blob + o_ft, feat.data(), sizeof(short4) * nf, cudaMemcpyHostToDevice, s));
  GSB_CHECK(cudaMemcpyAsync(blob + o_wk, weaks.data(), sizeof(gsb::Weak) * nw, cudaMemcpyHostToDevice, s));
  GSB_CHECK(cudaMemcpyAsync(blob + o_sb, c->subsets, 4 * (size_t)nsub, cudaMemcpyHostToDevice, s));
  GSB_CHECK(cudaMemcpyAsync(blob + o_st, stages.data(), sizeof(gsb::Stage) * nst, cudaMemcpyHostToDevice, s));
  GSB_CHECK(cudaStreamSynchronize(s));  // the host vectors above go out of scope
  gsb::k_lbp_window_one<<<1, 32, 0, s>>>(ii, iw, ih, x, y, reinterpret_cast<const short4 *>(blob + o_ft),
                                         reinterpret_cast<const gsb::Weak *>(blob + o_wk),
                                         reinterpret_cast<const int *>(blob + o_sb),
                                         reinterpret_cast<const gsb::Stage *>(blob + o_st), nst, out_dev);
  GSB_LAUNCHED(1);
  return 0;
}
}
