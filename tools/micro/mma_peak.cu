// Microbenchmark: peak rate of the legacy warp-level mma.sync.m16n8k16 bf16 path on sm_100a (register operands only).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(float* out, int iters) {
  unsigned a[4] = {threadIdx.x, 2, 3, 4}, b[2] = {5, 6};
  float c[8][4] = {};
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(c[j][0]), "+f"(c[j][1]), "+f"(c[j][2]), "+f"(c[j][3])
                   : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
  }
  float s = 0;
  for (int j = 0; j < 8; ++j) s += c[j][0] + c[j][1] + c[j][2] + c[j][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  float* d; cudaMalloc(&d, 148 * 8 * 1024 * 4);
  for (int warps = 4; warps <= 32; warps *= 2) {
    for (int ctas = 1; ctas <= 2; ++ctas) {
      if (warps * ctas > 64) continue;
      const int iters = 20000;
      k<<<148 * ctas, warps * 32>>>(d, 100);
      cudaDeviceSynchronize();
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      cudaEventRecord(e0);
      k<<<148 * ctas, warps * 32>>>(d, iters);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      double flops = 2.0 * 16 * 8 * 16 * 8.0 * iters * warps * ctas * 148;
      printf("warps/CTA %d CTAs/SM %d: %.1f TFLOP/s\n", warps, ctas, flops / ms / 1e9);
    }
  }
  return 0;
}
