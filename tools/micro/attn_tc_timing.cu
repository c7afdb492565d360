// Phase timing of the tcgen05 attention forward (clock64 per phase of one softmax warp per query tile, CTA 0).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --extended-lambda -lcuda -o attn_tc_timing attn_tc_timing.cu
#define A4R_ATTN_TIMING 1
#include <cstdarg>
#include <cstdio>
#include <vector>
#include "../../adapter4rec_b200/csrc/attention_tc_sm100.cu"
int a4r_set_error(int code, const char* fmt, ...) { va_list ap; va_start(ap, fmt); vfprintf(stderr, fmt, ap); va_end(ap); fprintf(stderr, "\n"); return code; }
int a4r_num_sms() { return 148; }
void a4r_count_launch(int) {}
int main(int argc, char** argv) {
  const int N = argc > 1 ? atoi(argv[1]) : 704, L = argc > 2 ? atoi(argv[2]) : 197, heads = 12, H = heads * 64;
  std::vector<__nv_bfloat16> h(static_cast<size_t>(N) * L * 3 * H);
  unsigned s = 12345;
  for (auto& x : h) { s = s * 1664525u + 1013904223u; x = __float2bfloat16(((s >> 8) & 0xFFFF) / 32768.0f - 1.0f); }
  __nv_bfloat16 *qkv, *out; float* lse;
  cudaMalloc(&qkv, h.size() * 2); cudaMalloc(&out, static_cast<size_t>(N) * L * H * 2); cudaMalloc(&lse, static_cast<size_t>(N) * L * heads * 4);
  cudaMemcpy(qkv, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
  a4r_attn_args a{}; a.qkv = qkv; a.out = out; a.lse = lse; a.ld_qkv = 3 * H; a.ld_out = H; a.N = N; a.L = L; a.heads = heads; a.head_dim = 64; a.scale = 0.125f;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  long long zero[8][8] = {};
  for (int i = 0; i < 3; ++i) a4r_attn_vit_tc_fwd(&a, 0);
  cudaDeviceSynchronize();
  cudaMemcpyToSymbol(g_attn_timing, zero, sizeof(zero));
  a4r_attn_vit_tc_fwd(&a, 0);
  cudaDeviceSynchronize();
  {
    long long tr[16][10][6]; cudaMemcpyFromSymbol(tr, g_attn_trace, sizeof(tr));
    const long long base = tr[2][0][0];
    printf("timeline of CTA 0 (cycles since unit 2, tile 0 quad 0 began waiting for S):\n");
    for (int u = 2; u < 6; ++u) {
      printf(" unit %d issuer: scores t0 %lld t1 %lld | P V t0 seen %lld issued %lld t1 seen %lld issued %lld\n", u, tr[u][8][0] - base, tr[u][8][1] - base, tr[u][9][2] - base, tr[u][9][0] - base, tr[u][9][3] - base, tr[u][9][1] - base);
      for (int w = 0; w < 8; ++w) {
        if (tr[u][w][0] == 0) continue;
        printf("   tile %d quad %d: wait S %lld, S ready %lld, P ready (arrive) %lld, O ready %lld, O read %lld, stored %lld\n", w >> 2, w & 3,
               tr[u][w][0] - base, tr[u][w][1] - base, tr[u][w][2] - base, tr[u][w][3] - base, tr[u][w][4] - base, tr[u][w][5] - base);
      }
    }
  }
  {
    long long t1[8][8]; cudaMemcpyFromSymbol(t1, g_attn_timing, sizeof(t1));
    printf("issuer (CTA 0, one launch): score MMA group (4 x N=%d) %.0f cycles, P V group (%d x N=64, A in TMEM) %.0f cycles (per tile, issue -> commit arrival)\n",
           (L + 15) / 16 * 16, (double)t1[0][7] / (2.0 * t1[0][6]), (L + 15) / 16, (double)t1[1][7] / (2.0 * t1[0][6]));
  }
  cudaEventRecord(e0);
  for (int i = 0; i < 10; ++i) a4r_attn_vit_tc_fwd(&a, 0);
  cudaEventRecord(e1); cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  printf("fwd N=%d L=%d: %.1f us per launch (%s)\n", N, L, ms * 100, cudaGetErrorString(cudaGetLastError()));
  long long t[8][8]; cudaMemcpyFromSymbol(t, g_attn_timing, sizeof(t));
  const char* names[6] = {"wait S (MMA)", "first block load", "softmax blocks", "wait O (P V MMA)", "O read-out", "store"};
  for (int tile = 0; tile < 8; ++tile) {
    if (t[tile][6] == 0) continue;
    long long units = t[tile][6] ? t[tile][6] : 1, tot = 0;
    for (int i = 0; i < 6; ++i) tot += t[tile][i];
    printf("tile %d quad %d, %lld units, %.0f cycles per unit:\n", tile >> 2, tile & 3, units, (double)tot / units);
    for (int i = 0; i < 6; ++i) printf("   %-18s %8.0f cycles (%.1f%%)\n", names[i], (double)t[tile][i] / units, 100.0 * t[tile][i] / tot);
  }
  return 0;
}
