// mix_probe.cu -- what HBM delivers for gs_integral's traffic mix (1 byte read : 4 bytes written) with an ideal
// streaming kernel: every thread reads 8 pixels (one 64-bit load) and writes 8 u32 (two 128-bit stores), no
// arithmetic to speak of.  Compared with a 1:1 copy and a write-only fill of the same kernel shape.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probe/mix_probe tools/probe/mix_probe.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_expand(uint32_t *__restrict__ out, const uint8_t *__restrict__ in, size_t n8, int cs) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (size_t)gridDim.x * blockDim.x) {
    const uint2 v = __ldg(reinterpret_cast<const uint2 *>(in) + i);
    uint4 a = make_uint4(v.x & 0xFF, (v.x >> 8) & 0xFF, (v.x >> 16) & 0xFF, v.x >> 24);
    uint4 b = make_uint4(v.y & 0xFF, (v.y >> 8) & 0xFF, (v.y >> 16) & 0xFF, v.y >> 24);
    uint4 *o = reinterpret_cast<uint4 *>(out) + 2 * i;
    if (cs) {
      asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(o), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w) : "memory");
      asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(o + 1), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w) : "memory");
    } else {
      o[0] = a, o[1] = b;
    }
  }
}
__global__ void k_copy(uint4 *__restrict__ out, const uint4 *__restrict__ in, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out[i] = __ldg(in + i);
}
__global__ void k_fill(uint4 *__restrict__ out, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out[i] = make_uint4(1, 2, 3, 4);
}
static float time_best(int which, int blocks, int threads, uint32_t *out, const uint8_t *in, size_t npx) {
  cudaEvent_t a, b;
  cudaEventCreate(&a), cudaEventCreate(&b);
  float ms, mn = 1e9f;
  for (int r = 0; r < 12; r++) {
    cudaEventRecord(a);
    if (which == 0) k_expand<<<blocks, threads>>>(out, in, npx / 8, 0);
    else if (which == 1) k_expand<<<blocks, threads>>>(out, in, npx / 8, 1);
    else if (which == 2) k_copy<<<blocks, threads>>>((uint4 *)out, (const uint4 *)(out + npx / 2), npx / 8);
    else k_fill<<<blocks, threads>>>((uint4 *)out, npx / 4);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    cudaEventElapsedTime(&ms, a, b);
    if (r >= 2 && ms < mn) mn = ms;
  }
  cudaEventDestroy(a), cudaEventDestroy(b);
  return mn;
}
int main() {
  const size_t npx = (size_t)1 << 30;     // 1 Gi pixels: 1 GiB in, 4 GiB out
  uint8_t *in = nullptr;
  uint32_t *out = nullptr;
  if (cudaMalloc(&in, npx) != cudaSuccess || cudaMalloc(&out, npx * 4) != cudaSuccess) {
    printf("allocation failed\n");
    return 1;
  }
  cudaMemset(in, 7, npx);
  const int blocks = 148 * 16;
  const int tl[3] = {128, 256, 512};
  for (int t = 0; t < 3; t++) {
    const int threads = tl[t];
    const double e0 = 5.0 * npx / (time_best(0, blocks, threads, out, in, npx) * 1e-3) / 1e9;
    const double e1 = 5.0 * npx / (time_best(1, blocks, threads, out, in, npx) * 1e-3) / 1e9;
    const double cp = 4.0 * npx / (time_best(2, blocks, threads, out, in, npx) * 1e-3) / 1e9;   // npx/8 uint4 read + written
    const double fl = 4.0 * npx / (time_best(3, blocks, threads, out, in, npx) * 1e-3) / 1e9;
    printf("threads %d: expand 1:4 plain %.0f GB/s, st.cs %.0f GB/s | copy 1:1 %.0f GB/s | fill %.0f GB/s\n", threads, e0, e1, cp, fl);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
