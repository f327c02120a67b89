// box.cu -- gs_blur and gs_adaptive_threshold (reference grayskull.h:230-247, 268-283).
//
// Both need mean = floor(S / count), S = sum of the (2r+1)^2 window clipped to the image,
// count = number of in-image taps = (#in-image columns) * (#in-image rows).  2 B/pixel of
// compulsory HBM traffic regardless of the radius, so the kernel must stay under ~10 issue
// slots per pixel.  Fast path (radius 1..7, width % 16 == 0), one 256 x 64 output tile per CTA:
//   0. one TMA box loads the (256+32) x (64+2r) byte tile (16-byte column halo: a TMA box must
//      start on a 16-byte boundary, see tools/probe/tma_probe.cu); outside the image reads as 0,
//      which is exactly the contribution of a clipped tap to the SUM;
//   1. vertical pass: 68 word-columns x 4 row bands; a thread keeps the running column sums of
//      its 4 columns as two u16x2 words and rolls them down the band with one IADD3 per word
//      (sum += entering row - leaving row on both 16-bit lanes at once: no lane can borrow
//      because every true lane value is in [0, 65535]); column sums (<= 15*255) go to shared
//      memory as u16;
//   2. horizontal pass: a warp per row, a lane per 8 pixels; the (2r+1)-wide window sums for
//      two adjacent pixels are formed on 16-bit lane pairs from a rolling sum of pair words
//      (4 integer ops per 2 pixels, radius independent); max (2*7+1)^2*255 = 57375 < 65536;
//   3. exact division without integer divide: with m = ceil(2^24 / count),
//        fma_rd(2^23 + S, m * 2^-24, 2^23 - m/2) = 2^23 + floor(S * m / 2^24)
//      is computed exactly before its single round-down, and floor(S*m/2^24) == S / count for
//      all S <= 255 * count, count <= 225 (checked exhaustively in tests/test_host_logic.py);
//      the quotient is the low byte of the float's bit pattern.  Interior pixels use the
//      compile-time constants for count = (2r+1)^2, clipped pixels a 226-entry table.
// Other radii / widths take the generic kernel (one thread per pixel, any radius).
#include "common.cuh"

namespace gsb {

constexpr int BX_TW = 256;                 // output tile width
constexpr int BX_TH = 64;                  // output tile height
constexpr int BX_PW = 72;                  // tile pitch in 32-bit words: image bytes [x0-16, x0+272)
constexpr int BX_NC = 68;                  // word-columns that carry column sums: bytes [x0-8, x0+264)
constexpr int BX_SP = 272;                 // column-sum row pitch in u16 (columns x0-8 .. x0+263)
constexpr int BX_RMAX = 7;
constexpr int BX_BANDS = 4, BX_BAND_H = BX_TH / BX_BANDS;   // 16
constexpr int BX_THREADS = 288;            // 272 phase-1 items (68 word-columns x 4 bands) -> 9 warps
constexpr int BX_TILE_WORDS = BX_PW * (BX_TH + 2 * BX_RMAX);
constexpr int BX_SMEM = BX_TILE_WORDS * 4 + BX_TH * BX_SP * 2 + 226 * 8 + 16;

struct DivMagic {
  float inv, k;
};
__host__ __device__ inline DivMagic div_magic(unsigned count) {
  unsigned m = (16777216u + count - 1u) / count;          // ceil(2^24 / count) <= 2^24: exact float
  DivMagic d;
  d.inv = (float)m * 5.9604644775390625e-08f;             // m * 2^-24, exact
  d.k = 8388608.0f - 0.5f * (float)m;                     // 2^23 - m/2, exact (|.| < 2^23, ulp 0.5)
  return d;
}

// floor(s / count) for a 16-bit s, as the low byte of the returned bit pattern
__device__ __forceinline__ uint32_t div_lo(uint32_t t, float inv, float k) {
  return __float_as_uint(__fmaf_rd(__uint_as_float((t & 0xFFFFu) | 0x4B000000u), inv, k));
}
__device__ __forceinline__ uint32_t div_hi(uint32_t t, float inv, float k) {
  return __float_as_uint(__fmaf_rd(__uint_as_float(prmt(t, 0x4B000000u, 0x7632)), inv, k));
}

// window sums for 4 output pixel pairs from 12 pair words V[m] = (s_2m, s_2m+1), s_i = column
// sum of image column x - 8 + i; output pair p = pixels (x+2p, x+2p+1).
template <int R>
__device__ __forceinline__ void window_sums(const uint32_t (&V)[12], uint32_t (&T)[4]) {
  constexpr bool ODD = (R & 1) != 0;
  constexpr int NP = ODD ? R : R + 1;
  constexpr int A0 = 8 - R;
  constexpr int M0 = ODD ? (A0 + 1) / 2 : A0 / 2;
  uint32_t ps = V[M0];
#pragma unroll
  for (int i = 1; i < NP; i++) ps += V[M0 + i];
#pragma unroll
  for (int p = 0; p < 4; p++) {
    const int m = M0 + p;
    const uint32_t tot = ps + prmt(ps, ps, 0x1032);   // both lanes = lane0 + lane1
    if (ODD) T[p] = tot + prmt(V[m - 1], V[m + NP], 0x5432);   // + (s_a, s_{a+2R+1})
    else T[p] = tot - prmt(V[m + R], V[m], 0x5432);            // - (s_{a+2R+1}, s_a)
    if (p < 3) ps = ps + V[m + NP] - V[m];
  }
}

template <int R, bool ADAPTIVE>
__global__ void __launch_bounds__(BX_THREADS)
k_box_tma(const __grid_constant__ CUtensorMap tmap, uint8_t *__restrict__ dst, unsigned w, unsigned h,
          unsigned tiles_x, unsigned tiles_y, int cparam) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint32_t *tile = reinterpret_cast<uint32_t *>(smem_raw);
  uint16_t *colsum = reinterpret_cast<uint16_t *>(smem_raw + BX_TILE_WORDS * 4);
  float2 *magic = reinterpret_cast<float2 *>(smem_raw + BX_TILE_WORDS * 4 + BX_TH * BX_SP * 2);
  uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw + BX_TILE_WORDS * 4 + BX_TH * BX_SP * 2 + 226 * 8);

  unsigned bid = blockIdx.x;
  const unsigned tx = bid % tiles_x;
  bid /= tiles_x;
  const unsigned ty = bid % tiles_y;
  const unsigned frame = bid / tiles_y;
  const int x0 = tx * BX_TW, y0 = ty * BX_TH;
  constexpr int ROWS = BX_TH + 2 * R;
  constexpr int FULL = (2 * R + 1);

  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    mbar_fence_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar, BX_PW * 4 * ROWS);
    tma_load_3d(tile, &tmap, x0 / 4 - 4, y0 - R, frame, bar);
  }
  if (threadIdx.x >= 1 && threadIdx.x < 226) {   // clipped-count division table
    DivMagic d = div_magic(threadIdx.x);
    magic[threadIdx.x] = make_float2(d.inv, d.k);
  }
  mbar_wait(bar, 0);

  // ---- phase 1: vertical rolling sums ------------------------------------------------------
  if (threadIdx.x < BX_NC * BX_BANDS) {
    const int c = threadIdx.x % BX_NC, band = threadIdx.x / BX_NC;
    const uint32_t *in = tile + (band * BX_BAND_H) * BX_PW + c + 2;   // word of byte x0 - 8 + 4c
    uint32_t se = 0, so = 0;
#pragma unroll
    for (int i = 0; i < 2 * R; i++) {
      uint32_t v = in[i * BX_PW];
      se += v & 0x00FF00FFu;
      so += prmt(v, 0, 0x4341);
    }
    uint16_t *out = colsum + (band * BX_BAND_H) * BX_SP + 4 * c;
#pragma unroll
    for (int i = 0; i < BX_BAND_H; i++) {
      uint32_t vin = in[(i + 2 * R) * BX_PW];
      se += vin & 0x00FF00FFu;
      so += prmt(vin, 0, 0x4341);
      uint2 o;
      o.x = prmt(se, so, 0x5410);   // (s0, s1)
      o.y = prmt(se, so, 0x7632);   // (s2, s3)
      *reinterpret_cast<uint2 *>(out + i * BX_SP) = o;
      uint32_t vout = in[i * BX_PW];
      se -= vout & 0x00FF00FFu;
      so -= prmt(vout, 0, 0x4341);
    }
  }
  __syncthreads();

  // ---- phase 2: horizontal window sums, division, store ------------------------------------
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int x = x0 + 8 * lane;
  const bool live = x < (int)w;
  // lanes whose 8 windows never leave the image horizontally
  const bool xin = !live || (x - R >= 0 && x + 7 + R <= (int)w - 1);
  const bool warp_xin = __all_sync(0xFFFFFFFFu, xin);
  int cw[8];
#pragma unroll
  for (int j = 0; j < 8; j++) cw[j] = min(x + j + R, (int)w - 1) - max(x + j - R, 0) + 1;
  constexpr DivMagic FM = {
      (float)((16777216u + FULL * FULL - 1u) / (FULL * FULL)) * 5.9604644775390625e-08f,
      8388608.0f - 0.5f * (float)((16777216u + FULL * FULL - 1u) / (FULL * FULL))};
  uint8_t *outp = dst + (size_t)frame * w * h + x;

  for (int yo = warp; yo < BX_TH; yo += BX_THREADS / 32) {
    const int y = y0 + yo;
    if (y >= (int)h) break;
    const uint4 *sp = reinterpret_cast<const uint4 *>(colsum + yo * BX_SP + 8 * lane);
    uint32_t V[12], T[4];
    uint4 a = sp[0], b = sp[1], c4 = sp[2];
    V[0] = a.x, V[1] = a.y, V[2] = a.z, V[3] = a.w;
    V[4] = b.x, V[5] = b.y, V[6] = b.z, V[7] = b.w;
    V[8] = c4.x, V[9] = c4.y, V[10] = c4.z, V[11] = c4.w;
    window_sums<R>(V, T);
    const int ch = min(y + R, (int)h - 1) - max(y - R, 0) + 1;
    uint32_t q[8];
    if (warp_xin && ch == FULL) {
#pragma unroll
      for (int p = 0; p < 4; p++) {
        q[2 * p] = div_lo(T[p], FM.inv, FM.k);
        q[2 * p + 1] = div_hi(T[p], FM.inv, FM.k);
      }
    } else {
#pragma unroll
      for (int p = 0; p < 4; p++) {
        float2 m0 = magic[cw[2 * p] * ch], m1 = magic[cw[2 * p + 1] * ch];
        q[2 * p] = div_lo(T[p], m0.x, m0.y);
        q[2 * p + 1] = div_hi(T[p], m1.x, m1.y);
      }
    }
    uint2 o;
    if (ADAPTIVE) {
      // dst = src > (int)mean - c ? 255 : 0   (reference :244-245)
      const uint2 sv = *reinterpret_cast<const uint2 *>(tile + (yo + R) * BX_PW + 2 * lane + 4);
      uint32_t r0 = 0, r1 = 0;
#pragma unroll
      for (int j = 0; j < 4; j++) {
        int s0 = (sv.x >> (8 * j)) & 0xFF, s1 = (sv.y >> (8 * j)) & 0xFF;
        int t0 = (int)(q[j] & 0xFF) - cparam, t1 = (int)(q[4 + j] & 0xFF) - cparam;
        r0 |= (s0 > t0 ? 0xFFu : 0u) << (8 * j);
        r1 |= (s1 > t1 ? 0xFFu : 0u) << (8 * j);
      }
      o.x = r0, o.y = r1;
    } else {
      o.x = prmt(prmt(q[0], q[1], 0x0040), prmt(q[2], q[3], 0x0040), 0x5410);
      o.y = prmt(prmt(q[4], q[5], 0x0040), prmt(q[6], q[7], 0x0040), 0x5410);
    }
    if (live) st_cs_u2(outp + (size_t)y * w, o);
  }
}

// ---- generic: any radius, width, alignment; one thread per pixel ----------------------------
template <bool ADAPTIVE>
__global__ void k_box_generic(uint8_t *__restrict__ dst, const uint8_t *__restrict__ src, unsigned w,
                              unsigned h, unsigned n, unsigned r, int cparam) {
  const unsigned x = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= w || y >= h) return;
  const unsigned xa = x > r ? x - r : 0, xb = min(x + (unsigned long long)r, (unsigned long long)w - 1);
  const unsigned ya = y > r ? y - r : 0, yb = min(y + (unsigned long long)r, (unsigned long long)h - 1);
  const unsigned count = (xb - xa + 1) * (yb - ya + 1);
  for (unsigned f = blockIdx.z; f < n; f += gridDim.z) {
    const uint8_t *s = src + (size_t)f * w * h;
    unsigned sum = 0;
    for (unsigned yy = ya; yy <= yb; yy++)
      for (unsigned xx = xa; xx <= xb; xx++) sum += s[(size_t)yy * w + xx];
    const unsigned mean = sum / count;
    uint8_t v;
    if (ADAPTIVE) v = ((int)s[(size_t)y * w + x] > (int)(mean - (unsigned)cparam)) ? 255 : 0;
    else v = (uint8_t)mean;
    dst[(size_t)f * w * h + (size_t)y * w + x] = v;
  }
}

template <int R, bool ADAPTIVE>
static int launch_box_r(const CUtensorMap &tmap, uint8_t *dst, unsigned w, unsigned h, unsigned n,
                        int cparam, cudaStream_t s) {
  static bool configured = false;  // per instantiation; one device per process
  if (!configured) {
    GSB_CHECK(cudaFuncSetAttribute(k_box_tma<R, ADAPTIVE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   BX_SMEM));
    configured = true;
  }
  const unsigned tiles_x = (w + BX_TW - 1) / BX_TW, tiles_y = (h + BX_TH - 1) / BX_TH;
  const unsigned long long blocks = (unsigned long long)tiles_x * tiles_y * n;
  GSB_ASSERT(blocks < 0x7FFFFFFFull);
  k_box_tma<R, ADAPTIVE><<<(unsigned)blocks, BX_THREADS, BX_SMEM, s>>>(tmap, dst, w, h, tiles_x, tiles_y,
                                                                        cparam);
  GSB_LAUNCHED(1);
  return 0;
}

template <bool ADAPTIVE>
static int launch_box(uint8_t *dst, const uint8_t *src, unsigned w, unsigned h, unsigned n, unsigned r,
                      int cparam, cudaStream_t s) {
  if (n == 0) return 0;
  CUtensorMap tmap;
  if (r >= 1 && r <= BX_RMAX && tma_ok(src, w) && tma_ok(dst, w) &&
      make_tmap_u8frames(&tmap, src, w, h, n, BX_PW, BX_TH + 2 * r)) {
    switch (r) {
      case 1: return launch_box_r<1, ADAPTIVE>(tmap, dst, w, h, n, cparam, s);
      case 2: return launch_box_r<2, ADAPTIVE>(tmap, dst, w, h, n, cparam, s);
      case 3: return launch_box_r<3, ADAPTIVE>(tmap, dst, w, h, n, cparam, s);
      case 4: return launch_box_r<4, ADAPTIVE>(tmap, dst, w, h, n, cparam, s);
      case 5: return launch_box_r<5, ADAPTIVE>(tmap, dst, w, h, n, cparam, s);
      case 6: return launch_box_r<6, ADAPTIVE>(tmap, dst, w, h, n, cparam, s);
      default: return launch_box_r<7, ADAPTIVE>(tmap, dst, w, h, n, cparam, s);
    }
  }
  dim3 block(32, 8), grid((w + 31) / 32, (h + 7) / 8, n < 65535u ? n : 65535u);
  k_box_generic<ADAPTIVE><<<grid, block, 0, s>>>(dst, src, w, h, n, r, cparam);
  GSB_LAUNCHED(1);
  return 0;
}

}  // namespace gsb

extern "C" {
int gs_b200_blur_batch(uint8_t *dst, const uint8_t *src, unsigned w, unsigned h, unsigned n,
                       unsigned radius, gs_b200_stream s) {
  GSB_ASSERT(dst && src && w > 0 && h > 0);  // reference :269
  return gsb::launch_box<false>(dst, src, w, h, n, radius, 0, static_cast<cudaStream_t>(s));
}
int gs_b200_adaptive_threshold_batch(uint8_t *dst, const uint8_t *src, unsigned w, unsigned h,
                                     unsigned n, unsigned radius, int c, gs_b200_stream s) {
  GSB_ASSERT(dst && src && w > 0 && h > 0);  // reference :232
  return gsb::launch_box<true>(dst, src, w, h, n, radius, c, static_cast<cudaStream_t>(s));
}
}
