// scan.cuh -- per-frame exclusive scan of small count arrays (ordered-compaction helper).
#pragma once
#include "common.cuh"

namespace gsb {

// one CTA per frame: exclusive scan of the row counts (in place), counts[f] = min(total, cap)
static __global__ void __launch_bounds__(1024)
k_row_scan(unsigned *__restrict__ rowcount, unsigned rows, unsigned *__restrict__ counts, unsigned cap) {
  __shared__ unsigned wsum[32], wexcl[32];
  __shared__ unsigned carry_s, chunk_total;
  unsigned *rc = rowcount + (size_t)blockIdx.x * rows;
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (unsigned base = 0; base < rows; base += 1024) {
    const unsigned i = base + threadIdx.x;
    const unsigned v = i < rows ? rc[i] : 0;
    unsigned incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      unsigned t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
      if (lane >= (unsigned)o) incl += t;
    }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      const unsigned ws = wsum[lane];
      unsigned wi = ws;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        unsigned t = __shfl_up_sync(0xFFFFFFFFu, wi, o);
        if (lane >= (unsigned)o) wi += t;
      }
      wexcl[lane] = wi - ws;
      if (lane == 31) chunk_total = wi;
    }
    __syncthreads();
    const unsigned carry = carry_s;
    if (i < rows) rc[i] = carry + wexcl[warp] + incl - v;
    __syncthreads();
    if (threadIdx.x == 0) carry_s = carry + chunk_total;
    __syncthreads();
  }
  if (threadIdx.x == 0) counts[blockIdx.x] = min(carry_s, cap);
}

}  // namespace gsb
