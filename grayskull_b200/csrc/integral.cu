// integral.cu -- gs_integral (reference grayskull.h:744-752): inclusive u32 summed-area table,
// ii[y][x] = sum of src over [0..x] x [0..y], modular 32-bit arithmetic (any association order
// is bit-identical).  Compulsory traffic: 1 B read + 4 B written per pixel.
//
// Round-1 implementation: two passes.
//   k_integral_rows : a warp per row; each lane takes 4 (vectorised) or 1 pixels per step,
//                     lane-local prefix + warp shuffle scan + running carry; writes row prefixes.
//   k_integral_cols : a thread per column; running sum down the rows, in place (coalesced over x).
// Traffic is 13 B/pixel instead of 5 (the table is written, re-read and re-written); the
// single-pass chained-band kernel that removes the second trip is the next optimisation
// ("integral").
#include "common.cuh"

namespace gsb {

template <bool VEC>
__global__ void k_integral_rows(uint32_t *__restrict__ ii, const uint8_t *__restrict__ src, unsigned w,
                                unsigned h, unsigned n) {
  const unsigned lane = threadIdx.x & 31;
  const unsigned long long row = (unsigned long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= (unsigned long long)h * n) return;
  const uint8_t *s = src + row * w;
  uint32_t *d = ii + row * w;
  uint32_t carry = 0;
  if (VEC) {
    for (unsigned x = lane * 4; x < ((w + 127) / 128) * 128; x += 128) {
      uint32_t v = x < w ? __ldg(reinterpret_cast<const uint32_t *>(s + x)) : 0u;
      uint32_t p0 = v & 0xFF, p1 = p0 + ((v >> 8) & 0xFF), p2 = p1 + ((v >> 16) & 0xFF), p3 = p2 + (v >> 24);
      uint32_t incl = p3;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if (lane >= (unsigned)o) incl += t;
      }
      const uint32_t base = carry + incl - p3;
      if (x < w) *reinterpret_cast<uint4 *>(d + x) = make_uint4(base + p0, base + p1, base + p2, base + p3);
      carry += __shfl_sync(0xFFFFFFFFu, incl, 31);
    }
  } else {
    for (unsigned x = lane; x < ((w + 31) / 32) * 32; x += 32) {
      uint32_t p = x < w ? s[x] : 0u;
      uint32_t incl = p;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if (lane >= (unsigned)o) incl += t;
      }
      if (x < w) d[x] = carry + incl;
      carry += __shfl_sync(0xFFFFFFFFu, incl, 31);
    }
  }
}

__global__ void k_integral_cols(uint32_t *__restrict__ ii, unsigned w, unsigned h, unsigned n) {
  const unsigned x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= w) return;
  for (unsigned f = blockIdx.y; f < n; f += gridDim.y) {
    uint32_t *p = ii + (size_t)f * w * h + x;
    uint32_t acc = 0;
    unsigned y = 0;
    for (; y + 4 <= h; y += 4) {
      uint32_t a = p[(size_t)y * w], b = p[(size_t)(y + 1) * w], c = p[(size_t)(y + 2) * w],
               d = p[(size_t)(y + 3) * w];
      a += acc, b += a, c += b, d += c;
      p[(size_t)y * w] = a, p[(size_t)(y + 1) * w] = b, p[(size_t)(y + 2) * w] = c, p[(size_t)(y + 3) * w] = d;
      acc = d;
    }
    for (; y < h; y++) acc += p[(size_t)y * w], p[(size_t)y * w] = acc;
  }
}

}  // namespace gsb

extern "C" int gs_b200_integral_batch(uint32_t *ii, const uint8_t *src, unsigned w, unsigned h, unsigned n,
                                      gs_b200_stream s) {
  GSB_ASSERT(src && ii && w > 0 && h > 0);  // reference :745
  if (n == 0) return 0;
  cudaStream_t st = static_cast<cudaStream_t>(s);
  const unsigned long long rows = (unsigned long long)h * n;
  const unsigned blocks = (unsigned)((rows + 7) / 8);
  const bool vec = (w % 4 == 0) && reinterpret_cast<uintptr_t>(src) % 4 == 0 &&
                   reinterpret_cast<uintptr_t>(ii) % 16 == 0;
  if (vec) gsb::k_integral_rows<true><<<blocks, 256, 0, st>>>(ii, src, w, h, n);
  else gsb::k_integral_rows<false><<<blocks, 256, 0, st>>>(ii, src, w, h, n);
  GSB_LAUNCHED(1);
  dim3 grid((w + 127) / 128, n < 65535u ? n : 65535u);
  gsb::k_integral_cols<<<grid, 128, 0, st>>>(ii, w, h, n);
  GSB_LAUNCHED(1);
  return 0;
}
