// Phase timing of the tcgen05 GEMM epilogue (clock64 in epilogue warp 0 of CTA 0): TMEM read-out wait, math, stores, wait for MMA.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --extended-lambda -lcuda -o t gemm_epi_timing.cu
#define A4R_GEMM_TIMING 1
#include <cstdarg>
#include <cstdio>
#include <vector>
#include "../../adapter4rec_b200/csrc/gemm_sm100.cu"
int a4r_set_error(int code, const char* fmt, ...) { va_list ap; va_start(ap, fmt); vfprintf(stderr, fmt, ap); va_end(ap); fprintf(stderr, "\n"); return code; }
int a4r_num_sms() { return 148; }
void a4r_count_launch(int) {}
extern "C" int a4r_device_check(void) { return 0; }
static void run(const char* name, int64_t M, int64_t N, int64_t K, int epi) {
  __nv_bfloat16 *A, *B, *C, *aux, *res; float* bias;
  cudaMalloc(&A, M * K * 2); cudaMalloc(&B, N * K * 2); cudaMalloc(&C, M * N * 2); cudaMalloc(&aux, M * N * 2); cudaMalloc(&res, M * N * 2); cudaMalloc(&bias, N * 4);
  cudaMemset(A, 0x3c, M * K * 2); cudaMemset(B, 0x3c, N * K * 2); cudaMemset(aux, 0x3c, M * N * 2); cudaMemset(res, 0x3c, M * N * 2); cudaMemset(bias, 0, N * 4);
  a4r_gemm_args a{}; a.A = A; a.B = B; a.C = C; a.M = M; a.N = N; a.K = K; a.lda = K; a.ldb = K; a.ldc = N; a.alpha = 1.0f; a.epilogue = epi;
  if (epi == A4R_EPI_GELU) { a.aux = aux; a.ldaux = N; a.bias = bias; }
  if (epi == A4R_EPI_DGELU) { a.aux = aux; a.ldaux = N; }
  if (epi == A4R_EPI_LINEAR) { a.residual = res; a.ldr = N; a.bias = bias; }
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 3; ++i) a4r_gemm_bf16_tn(&a, 0);
  cudaEventRecord(e0);
  for (int i = 0; i < 10; ++i) a4r_gemm_bf16_tn(&a, 0);
  cudaEventRecord(e1); cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long t[8]; cudaMemcpyFromSymbol(t, g_gemm_timing, sizeof(t));
  const double chunks = t[3] ? t[3] : 1, tiles = t[5] ? t[5] : 1;
  printf("%-28s M=%lld N=%lld K=%lld: %.1f us = %.0f TFLOP/s | per tile %.0f clk: wait-MMA %.0f, per chunk (%.1f chunks/tile): TMEM ld+wait %.0f, math %.0f, store %.0f  (%s)\n",
         name, (long long)M, (long long)N, (long long)K, ms * 100, 2.0 * M * N * K / (ms / 10 * 1e-3) / 1e12, t[6] / tiles, t[4] / tiles,
         chunks / tiles, t[0] / chunks, t[1] / chunks, t[2] / chunks, cudaGetErrorString(cudaGetLastError()));
  cudaFree(A); cudaFree(B); cudaFree(C); cudaFree(aux); cudaFree(res); cudaFree(bias);
}
int main() {
  run("FFN1 GELU (+aux)", 161280, 3072, 768, A4R_EPI_GELU);
  run("dFFN2 GELU'", 161280, 3072, 768, A4R_EPI_DGELU);
  run("attention.output (+res)", 161280, 768, 768, A4R_EPI_LINEAR);
  run("FFN2 (+res)", 161280, 768, 3072, A4R_EPI_LINEAR);
  return 0;
}
