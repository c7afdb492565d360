// Microbenchmark: per-SM throughput of MUFU.EX2, F2FP.BF16 pack, FFMA and FMNMX on sm_100a (ops / clock / SM).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_rates pipe_rates.cu && ./pipe_rates
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
template <int OP>
__global__ void k(float* out, int iters, long long* cyc) {
  float a[8];
  for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 0.001f + i * 0.1f;
  uint32_t acc = 0;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (OP == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (OP == 1) { uint32_t w; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(w) : "f"(a[i]), "f"(a[(i + 1) & 7])); acc ^= w; }
      if (OP == 2) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(1.0001f), "f"(0.5f));
      if (OP == 3) asm volatile("max.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(a[(i + 3) & 7]));
      if (OP == 4) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i])); asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(0.3f), "f"(-0.5f)); }
    }
  }
  long long t1 = clock64();
  float s = 0;
  for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int OP>
void run(const char* name, int warps) {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  int iters = 4096;
  k<OP><<<148, warps * 32>>>(out, iters, cyc);
  k<OP><<<148, warps * 32>>>(out, iters, cyc);
  cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
  double ops = (double)iters * 8 * warps * 32 * (OP == 4 ? 1 : 1);
  printf("%-22s warps/SM %2d: %.2f ops/clk/SM (%.1f clk per warp-instruction per SMSP)\n", name, warps, ops / c, c / ((double)iters * 8 * warps / 4));
  cudaFree(out); cudaFree(cyc);
}
int main() {
  for (int w : {4, 8, 16}) {
    if (w == 4) { run<0>("MUFU.EX2", 4); run<1>("F2FP.BF16.PACK", 4); run<2>("FFMA", 4); run<3>("FMNMX", 4); run<4>("EX2+FFMA pair", 4); }
    if (w == 8) { run<0>("MUFU.EX2", 8); run<1>("F2FP.BF16.PACK", 8); run<2>("FFMA", 8); run<3>("FMNMX", 8); run<4>("EX2+FFMA pair", 8); }
    if (w == 16) { run<0>("MUFU.EX2", 16); run<1>("F2FP.BF16.PACK", 16); run<2>("FFMA", 16); run<3>("FMNMX", 16); run<4>("EX2+FFMA pair", 16); }
  }
  return 0;
}
