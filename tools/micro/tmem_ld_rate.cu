// Microbenchmark: tensor-memory read-out bandwidth (tcgen05.ld -> registers) per SM on sm_100a, by shape and warp count.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_ld_rate tmem_ld_rate.cu && ./tmem_ld_rate
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
template <int N>
__device__ __forceinline__ uint32_t ld(uint32_t taddr) {
  uint32_t r[N];
  if constexpr (N == 8)
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr));
  if constexpr (N == 16)
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(taddr));
  if constexpr (N == 32)
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
                   "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]),
                   "=r"(r[30]), "=r"(r[31]) : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  uint32_t x = 0;
#pragma unroll
  for (int i = 0; i < N; ++i) x ^= r[i];
  return x;
}
// 16 lanes x 256 bits: a warp reads 16 lanes; .x4 = 16 registers per thread (16 lanes x 32 columns)
__device__ __forceinline__ uint32_t ld_16x256(uint32_t taddr) {
  uint32_t r[16];
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                 "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  uint32_t x = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) x ^= r[i];
  return x;
}
template <int N, int MODE>
__global__ void k(uint32_t* out, int iters, long long* cyc) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = slot + (static_cast<uint32_t>((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const uint32_t col = ((it * 4 + c) * N) & 255;
      if (MODE == 0) acc ^= ld<N>(base + col);
      else acc ^= ld_16x256(base + (col & 255));
    }
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(slot) : "memory");
}
template <int N, int MODE>
void run(const char* name, int warps) {
  uint32_t* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  const int iters = 2048;
  k<N, MODE><<<148, warps * 32>>>(out, iters, cyc);
  k<N, MODE><<<148, warps * 32>>>(out, iters, cyc);
  cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyHostToDevice == 0 ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToHost);
  double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
  const double bytes = (double)iters * 4 * warps * (MODE == 0 ? 32.0 * N * 4 : 16.0 * 32 * 4);
  printf("%-18s warps %2d: %7.1f B/clk/SM  (%5.1f clk per load per warp)  %s\n", name, warps, bytes / c, c / (iters * 4.0), cudaGetErrorString(cudaGetLastError()));
  cudaFree(out); cudaFree(cyc);
}
int main() {
  for (int w : {1, 4, 8, 16}) {
    if (w == 1) { run<8, 0>("32x32b.x8", 1); run<16, 0>("32x32b.x16", 1); run<32, 0>("32x32b.x32", 1); run<16, 1>("16x256b.x4", 1); }
    if (w == 4) { run<8, 0>("32x32b.x8", 4); run<16, 0>("32x32b.x16", 4); run<32, 0>("32x32b.x32", 4); run<16, 1>("16x256b.x4", 4); }
    if (w == 8) { run<8, 0>("32x32b.x8", 8); run<16, 0>("32x32b.x16", 8); run<32, 0>("32x32b.x32", 8); run<16, 1>("16x256b.x4", 8); }
    if (w == 16) { run<8, 0>("32x32b.x8", 16); run<16, 0>("32x32b.x16", 16); run<32, 0>("32x32b.x32", 16); run<16, 1>("16x256b.x4", 16); }
  }
  return 0;
}
