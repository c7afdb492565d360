// Phase timing of the tcgen05 attention backward (clock64 per phase of compute warp 0, CTA 0).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --extended-lambda -lcuda -o t attn_tc_bwd_timing.cu
#define A4R_ATTN_TIMING 1
#include <cstdarg>
#include <cstdio>
#include <vector>
#include "../../adapter4rec_b200/csrc/attention_tc_bwd_sm100.cu"
int a4r_set_error(int code, const char* fmt, ...) { va_list ap; va_start(ap, fmt); vfprintf(stderr, fmt, ap); va_end(ap); fprintf(stderr, "\n"); return code; }
int a4r_num_sms() { return 148; }
void a4r_count_launch(int) {}
int main(int argc, char** argv) {
  const int N = argc > 1 ? atoi(argv[1]) : 704, L = argc > 2 ? atoi(argv[2]) : 197, heads = 12, H = heads * 64;
  std::vector<__nv_bfloat16> h(static_cast<size_t>(N) * L * 3 * H);
  unsigned s = 12345;
  for (auto& x : h) { s = s * 1664525u + 1013904223u; x = __float2bfloat16(((s >> 8) & 0xFFFF) / 32768.0f - 1.0f); }
  __nv_bfloat16 *qkv, *dqkv, *out, *dout; float* lse;
  cudaMalloc(&qkv, h.size() * 2); cudaMalloc(&dqkv, h.size() * 2);
  cudaMalloc(&out, static_cast<size_t>(N) * L * H * 2); cudaMalloc(&dout, static_cast<size_t>(N) * L * H * 2); cudaMalloc(&lse, static_cast<size_t>(N) * L * heads * 4);
  cudaMemcpy(qkv, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(out, h.data(), static_cast<size_t>(N) * L * H * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dout, h.data() + 1000, static_cast<size_t>(N) * L * H * 2, cudaMemcpyHostToDevice);
  std::vector<float> hl(static_cast<size_t>(N) * L * heads, 5.5f);
  cudaMemcpy(lse, hl.data(), hl.size() * 4, cudaMemcpyHostToDevice);
  a4r_attn_args a{}; a.qkv = qkv; a.out = dqkv; a.dout = dout; a.ctx = out; a.lse = lse; a.ld_qkv = 3 * H; a.ld_out = H; a.N = N; a.L = L; a.heads = heads; a.head_dim = 64; a.scale = 0.125f;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 2; ++i) a4r_attn_vit_tc_bwd(&a, 0);
  cudaEventRecord(e0);
  for (int i = 0; i < 5; ++i) a4r_attn_vit_tc_bwd(&a, 0);
  cudaEventRecord(e1); cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  printf("bwd N=%d L=%d: %.1f us per launch (%s)\n", N, L, ms * 200, cudaGetErrorString(cudaGetLastError()));
  long long t[16]; cudaMemcpyFromSymbol(t, g_attn_bwd_timing, sizeof(t));
  const char* names[8] = {"wait S,dP (M1 M2)", "TMEM load + math", "wait P/dS tiles free (M3-M5 of previous step)", "store P, dS", "deferred dV/dK read-out",
                          "read + store dQ", "-", "wait dQ"};
  long long units = t[8] ? t[8] : 1, tot = 0;
  for (int i = 0; i < 8; ++i) tot += t[i];
  printf("compute warp 0 of 16, %lld units, %.0f cycles per unit (+ dQ read-out):\n", units, (double)tot / units);
  for (int i = 0; i < 8; ++i) printf("   %-46s %8.0f cycles per unit (%.1f%%)\n", names[i], (double)t[i] / units, 100.0 * t[i] / tot);
  return 0;
}
