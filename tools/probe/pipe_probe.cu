// pipe_probe.cu -- issue rate and dependent-issue latency of the integer / conversion instructions the box and
// integral kernels lean on (IDP.2A, IDP.4A, IMAD, IADD3, I2FP, FFMA, PRMT), one SM's worth of warps at a time.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_probe pipe_probe.cu && ./pipe_probe
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define N_IT 4096
template <int OP, int ILP>
__global__ void k(uint32_t *out, uint32_t seed, long long *cyc) {
  uint32_t a[ILP];
  float f[ILP];
#pragma unroll
  for (int i = 0; i < ILP; i++) a[i] = seed + threadIdx.x + i, f[i] = (float)(seed + i);
  uint32_t b = seed | 0x0101u;
  const long long t0 = clock64();
  for (int it = 0; it < N_IT; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) {
      if (OP == 0) asm volatile("dp2a.lo.u32.u32 %0, %1, %2, %0;" : "+r"(a[i]) : "r"(b), "r"(0x0101u));
      if (OP == 1) asm volatile("dp4a.u32.u32 %0, %1, %2, %0;" : "+r"(a[i]) : "r"(b), "r"(0x01010101u));
      if (OP == 2) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(seed));
      if (OP == 3) asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b));
      if (OP == 4) asm volatile("{ .reg .f32 t; cvt.rn.f32.u32 t, %0; mov.b32 %0, t; }" : "+r"(a[i]));
      if (OP == 5) asm volatile("fma.rm.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(1.0001f), "f"(8388608.0f));
      if (OP == 6) asm volatile("prmt.b32 %0, %0, %1, 0x5410;" : "+r"(a[i]) : "r"(b));
      if (OP == 7) asm volatile("dp2a.lo.u32.s32 %0, %1, %2, %0;" : "+r"(a[i]) : "r"(b), "r"(0x00FFu));
      if (OP == 8) asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b));
      if (OP == 9) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b), "r"(seed));
      if (OP == 10) asm volatile("shf.r.wrap.b32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(seed));
      if (OP == 11) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(1.0001f));
      if (OP == 12) asm volatile("{ .reg .b16 lo, hi; .reg .f32 t; mov.b32 {lo, hi}, %0; cvt.rn.f32.u16 t, hi; mov.b32 %0, t; }" : "+r"(a[i]));
      if (OP == 13) asm volatile("{ .reg .f32 t; mov.b32 t, %0; cvt.rzi.u32.f32 %0, t; }" : "+r"(a[i]));
      if (OP == 14) asm volatile("popc.b32 %0, %0;" : "+r"(a[i]));
      if (OP == 15) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(f[i]));
      if (OP == 16) asm volatile("vadd2.u32.u32.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(seed));
      if (OP == 17) asm volatile("max.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b));
    }
  }
  const long long t1 = clock64();
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s += a[i] + __float_as_uint(f[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int OP, int ILP>
static void run(const char *name, int warps) {
  uint32_t *out;
  long long *cyc, h;
  cudaMalloc(&out, 148 * 1024 * 4);
  cudaMalloc(&cyc, 8);
  k<OP, ILP><<<1, warps * 32>>>(out, 12345u, cyc);
  k<OP, ILP><<<1, warps * 32>>>(out, 12345u, cyc);
  cudaDeviceSynchronize();
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  const double per = (double)h / N_IT;   // cycles per loop iteration of one warp
  printf("%-10s ilp %d warps %2d : %7.2f cyc/iter  -> %.2f cyc per instr per warp, %.1f lane-instr/clk/SM\n", name, ILP, warps, per,
         per / ILP, 32.0 * warps * ILP / per);
  cudaFree(out), cudaFree(cyc);
}
#define ALL(OP, NAME) run<OP, 1>(NAME, 1); run<OP, 8>(NAME, 4); run<OP, 8>(NAME, 16); run<OP, 8>(NAME, 32);
int main() {
  ALL(0, "IDP.2A") ALL(7, "IDP.2A.S8") ALL(1, "IDP.4A") ALL(2, "IMAD") ALL(3, "IADD") ALL(4, "I2FP") ALL(5, "FFMA.RM") ALL(6, "PRMT") ALL(8, "IMAD.HI") ALL(9, "LOP3") ALL(10, "SHF") ALL(11, "FADD") ALL(12, "I2F.U16") ALL(13, "F2I") ALL(14, "POPC") ALL(15, "MUFU.RCP")
  ALL(16, "vadd2") ALL(17, "IMNMX")
  return 0;
}
