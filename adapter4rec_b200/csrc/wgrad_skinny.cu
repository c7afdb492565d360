// wgrad_skinny.cu — weight gradients of the TRAINABLE low-rank modules:  dW[N,K] = alpha * Aᵀ[N,M]·B[M,K]
//
// With a frozen backbone the only weight gradients on the path are those of AdapterBlock.fc_down/fc_up
// (r = 64 / 16; Downstream/Text/model/modules.py:131-134) and of loralib's lora_A / lora_B (r = 8..16;
// Downstream/Text/run.py:414-428): one of N, K is <= 64, the reduction runs over all M tokens, so the
// arithmetic intensity is ~r FLOP/B and the kernel is HBM-bound.  Each CTA owns one 64x64
// output tile and one slice of M, streams its two operand slices through a 3-stage cp.async ring, multiplies
// on mma.sync (both operands are "transposed" in memory: ldmatrix.trans does that for free) and writes an
// fp32 partial; a second kernel reduces the partials over the M-slices in fixed order (deterministic).
#include "a4r_common.cuh"
#include "mma_sync.cuh"

namespace {

constexpr int TN = 64, TK = 64, TM = 64;  // output tile 64x64, 64 reduction rows per stage
constexpr int STAGES = 3;
constexpr int WG_THREADS = 128;
constexpr int STAGE_BYTES = 2 * TM * 64 * 2;  // A slice [64 m][64 n] + B slice [64 m][64 k], bf16

struct WgradParams {
  const __nv_bfloat16* A;  // [M, lda], columns 0..N-1 used
  const __nv_bfloat16* B;  // [M, ldb], columns 0..K-1 used
  float* partial;          // [splits][tiles_n*tiles_k][64*64]
  int64_t lda, ldb, M;
  int N, K, tiles_n, tiles_k, splits;
  int64_t rows_per_split;  // multiple of TM
};

A4R_DEVICE uint32_t tile_off(int row, int chunk) { return static_cast<uint32_t>(row * 128 + (((chunk ^ row) & 7) << 4)); }

A4R_DEVICE void cp_async16(uint32_t dst, const void* src, bool pred) {
  const int sz = pred ? 16 : 0;  // src-size 0 => zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
A4R_DEVICE void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
A4R_DEVICE void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__global__ void __launch_bounds__(WG_THREADS) wgrad_kernel(const WgradParams p) {
  extern __shared__ __align__(128) uint8_t smem_wg[];
  const int tile = blockIdx.x % (p.tiles_n * p.tiles_k);
  const int split = blockIdx.x / (p.tiles_n * p.tiles_k);
  const int n0 = (tile / p.tiles_k) * TN, k0 = (tile % p.tiles_k) * TK;
  const int64_t m_begin = static_cast<int64_t>(split) * p.rows_per_split;
  int64_t m_end = m_begin + p.rows_per_split;
  if (m_end > p.M) m_end = p.M;
  const int nsteps = m_end > m_begin ? static_cast<int>((m_end - m_begin + TM - 1) / TM) : 0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t sbase = smem_u32(smem_wg);

  auto issue = [&](int step) {
    const int st = step % STAGES;
    const uint32_t sa = sbase + st * STAGE_BYTES, sb = sa + TM * 128;
    const int64_t m0 = m_begin + static_cast<int64_t>(step) * TM;
#pragma unroll
    for (int i = threadIdx.x; i < TM * 8; i += WG_THREADS) {
      const int r = i >> 3, ch = i & 7;
      const int64_t m = m0 + r;
      const bool row_ok = m < m_end;
      const int64_t mm = row_ok ? m : 0;
      const bool a_ok = row_ok && (n0 + ch * 8 < p.N), b_ok = row_ok && (k0 + ch * 8 < p.K);
      cp_async16(sa + tile_off(r, ch), p.A + (a_ok ? mm * p.lda + n0 + ch * 8 : 0), a_ok);
      cp_async16(sb + tile_off(r, ch), p.B + (b_ok ? mm * p.ldb + k0 + ch * 8 : 0), b_ok);
    }
  };

  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[i][e] = 0.0f;

  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < nsteps) issue(s);
    cp_async_commit();
  }
  for (int step = 0; step < nsteps; ++step) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    if (step + STAGES - 1 < nsteps) issue(step + STAGES - 1);
    cp_async_commit();
    const int st = step % STAGES;
    const uint32_t sa = sbase + st * STAGE_BYTES, sb = sa + TM * 128;
#pragma unroll
    for (int kk = 0; kk < TM; kk += 16) {
      uint32_t a[4];
      // A operand = (A slice)ᵀ: stored [m (reduction)][n]; this warp owns output rows warp*16..+16
      {
        const int mi = lane >> 3, r = lane & 7;
        const int row = kk + r + ((mi >> 1) << 3);
        const int chunk = (warp * 16 >> 3) + (mi & 1);
        ldsm_x4_t(a, sa + tile_off(row, chunk));
      }
#pragma unroll
      for (int np = 0; np < 4; ++np) {
        uint32_t b[4];
        {
          const int mi = lane >> 3, r = lane & 7;
          const int row = kk + r + ((mi & 1) << 3);
          const int chunk = (np * 16 >> 3) + (mi >> 1);
          ldsm_x4_t(b, sb + tile_off(row, chunk));
        }
        const uint32_t b0[2] = {b[0], b[1]}, b1[2] = {b[2], b[3]};
        mma_bf16_16816(acc[np * 2], a, b0);
        mma_bf16_16816(acc[np * 2 + 1], a, b1);
      }
    }
  }
  cp_async_wait<0>();
  // write the 64x64 fp32 partial of this (split, tile)
  float* out = p.partial + (static_cast<int64_t>(split) * p.tiles_n * p.tiles_k + tile) * (TN * TK);
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    const int r0 = warp * 16 + g, c = nt * 8 + 2 * t;
    *reinterpret_cast<float2*>(out + r0 * TK + c) = make_float2(acc[nt][0], acc[nt][1]);
    *reinterpret_cast<float2*>(out + (r0 + 8) * TK + c) = make_float2(acc[nt][2], acc[nt][3]);
  }
}

__global__ void wgrad_reduce_kernel(const float* __restrict__ partial, float* __restrict__ dW, int64_t ldw, int N, int K,
                                    int tiles_n, int tiles_k, int splits, float alpha, int accumulate) {
  const int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (idx >= static_cast<int64_t>(N) * K) return;
  const int n = static_cast<int>(idx / K), k = static_cast<int>(idx % K);
  const int tile = (n / TN) * tiles_k + (k / TK);
  const int off = (n % TN) * TK + (k % TK);
  float acc = 0.0f;
  for (int s = 0; s < splits; ++s)
    acc += partial[(static_cast<int64_t>(s) * tiles_n * tiles_k + tile) * (TN * TK) + off];
  float* o = dW + static_cast<int64_t>(n) * ldw + k;
  *o = accumulate ? *o + alpha * acc : alpha * acc;
}

void plan(int64_t M, int64_t N, int64_t K, int* tiles_n, int* tiles_k, int* splits, int64_t* rows_per_split) {
  *tiles_n = static_cast<int>((N + TN - 1) / TN);
  *tiles_k = static_cast<int>((K + TK - 1) / TK);
  const int tiles = *tiles_n * *tiles_k;
  int64_t want = (2 * 148 + tiles - 1) / tiles;  // ~2 CTAs per SM in flight
  const int64_t max_splits = (M + 4 * TM - 1) / (4 * TM);  // at least 256 rows per split
  if (want > max_splits) want = max_splits;
  if (want < 1) want = 1;
  int64_t rps = (M + want - 1) / want;
  rps = ((rps + TM - 1) / TM) * TM;
  if (rps < TM) rps = TM;
  *rows_per_split = rps;
  *splits = static_cast<int>((M + rps - 1) / rps);
  if (*splits < 1) *splits = 1;
}

}  // namespace

extern "C" size_t a4r_wgrad_workspace_bytes(int64_t M, int64_t N, int64_t K) {
  int tn, tk, sp;
  int64_t rps;
  plan(M, N, K, &tn, &tk, &sp, &rps);
  return static_cast<size_t>(sp) * tn * tk * TN * TK * sizeof(float);
}

extern "C" int a4r_wgrad_bf16(const void* A, int64_t lda, const void* B, int64_t ldb, float* dW, int64_t ldw,
                              int64_t M, int64_t N, int64_t K, float alpha, int32_t accumulate, void* workspace,
                              size_t workspace_bytes, a4r_stream_t stream_) {
  A4R_CHECK_ARG(A && B && dW, "wgrad: NULL pointer");
  A4R_CHECK_ARG(M >= 0 && N > 0 && K > 0 && N % 8 == 0 && K % 8 == 0, "wgrad: N and K must be positive multiples of 8");
  A4R_CHECK_ARG(lda >= N && ldb >= K && lda % 8 == 0 && ldb % 8 == 0 && ldw >= K, "wgrad: bad leading dimensions");
  A4R_CHECK_ARG(a4r_aligned16(A) && a4r_aligned16(B), "wgrad: A and B must be 16B aligned");
  const size_t need = a4r_wgrad_workspace_bytes(M, N, K);
  if (workspace == nullptr || workspace_bytes < need)
    return a4r_set_error(A4R_EWORKSPACE, "wgrad: workspace too small (%zu < %zu)", workspace_bytes, need);
  int rc = a4r_device_check();
  if (rc != A4R_OK) return rc;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  WgradParams p;
  p.A = static_cast<const __nv_bfloat16*>(A);
  p.B = static_cast<const __nv_bfloat16*>(B);
  p.partial = static_cast<float*>(workspace);
  p.lda = lda;
  p.ldb = ldb;
  p.M = M;
  p.N = static_cast<int>(N);
  p.K = static_cast<int>(K);
  plan(M, N, K, &p.tiles_n, &p.tiles_k, &p.splits, &p.rows_per_split);
  static bool attr_done = false;
  if (!attr_done) {
    A4R_CUDA_OK(cudaFuncSetAttribute(wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, STAGES * STAGE_BYTES));
    attr_done = true;
  }
  const int grid = p.tiles_n * p.tiles_k * p.splits;
  wgrad_kernel<<<grid, WG_THREADS, STAGES * STAGE_BYTES, stream>>>(p);
  A4R_LAUNCH_OK();
  const int64_t total = N * K;
  wgrad_reduce_kernel<<<static_cast<int>((total + 255) / 256), 256, 0, stream>>>(
      p.partial, dW, ldw, p.N, p.K, p.tiles_n, p.tiles_k, p.splits, alpha, accumulate);
  A4R_LAUNCH_OK();
  a4r_count_launch(2);
  return A4R_OK;
}
