// wgrad_sm100.cu — weight gradients on tcgen05:  dW[N, K] (+)= alpha * dYᵀ[N, M] · X[M, K]   (M = tokens, the reduction)
//
// The contraction runs over the TOKEN axis, which is the slow axis of both operands in memory (dY [M, N] and X [M, K] are
// row-major activations).  The K-major kernel of gemm_sm100.cu would need both operands transposed first (two extra HBM
// passes); tcgen05 does not: its shared-memory descriptors take MN-major operands, i.e. tiles whose CONTIGUOUS axis is the
// output axis.  A TMA box of [64 tokens x 64 columns] of a row-major activation with SWIZZLE_128B lands as 64 rows of
// 128 bytes — exactly the canonical MN-major SW128 layout ((8,8,m),(8,k)):((1,8,LBO),(64,SBO)) in elements: 8 token
// rows form a 1024-byte swizzle atom (SBO = 1024 B), the next 64-column chunk of the tile starts LBO = 8192 B later, and
// one MMA (16 tokens) advances the start address by 2048 B.
//
// Used for (a) full fine-tuning (Pretraining/Text/run.py:241-253, fine_tune_to=all — every nn.Linear
// weight of the encoder: 768x768, 3072x768, 768x3072 at M = 161,280 tokens per pass: tensor-bound) and (b) the skinny
// LoRA / adapter gradients (N or K <= 64: HBM-bound, one pass over the wide operand).
//
// Mapping: CTA = one (128 x BN) tile of dW x one SPLIT of the token range; warp 0 TMA producer, warp 1 MMA issuer, warp 2
// TMEM allocator, warps 4-7 epilogue (TMEM -> registers -> fp32 partial tile in the workspace).  A second kernel sums the
// partials in split order (deterministic — no atomics), applies alpha and the optional accumulation into dW.
// Ragged edges: rows beyond M, columns beyond N / K are zero-filled by TMA and contribute nothing; stores are guarded.
#include "a4r_common.cuh"

namespace {

constexpr int BT = 64;            // tokens per pipeline stage (4 MMAs of 16)
constexpr int TILE_N = 128;       // dW rows per CTA (the MMA's M)
constexpr int CHUNK_BYTES = 8192; // one [64 tokens x 64 columns] bf16 box

template <int BN>
struct WCfg {
  static constexpr int kStageA = (TILE_N / 64) * CHUNK_BYTES;
  static constexpr int kStageB = (BN / 64) * CHUNK_BYTES;
  static constexpr int kStage = kStageA + kStageB;
  static constexpr int kStages = BN == 256 ? 4 : 6;
  static constexpr int kSmem = kStages * kStage + 1024 + 256;
  static constexpr uint32_t kTmemCols = BN < 32 ? 32 : BN;
};

struct WParams {
  float* partial;        // [splits, Npad, Kpad] fp32, Npad = tiles_n * 128, Kpad = tiles_k * BN
  int N, K;
  int tiles_n, tiles_k, splits;
  int kb_total;          // ceil(M / 64)
  int kb_per_split;
  int64_t Npad, Kpad;
};

// MN-major, 128-byte swizzle shared-memory descriptor (see the header comment)
A4R_DEVICE uint64_t umma_desc_mn_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);      // start address
  d |= static_cast<uint64_t>(CHUNK_BYTES >> 4) << 16;           // leading byte offset: next 64-wide MN chunk
  d |= static_cast<uint64_t>(1024 >> 4) << 32;                  // stride byte offset: next group of 8 k-rows
  d |= static_cast<uint64_t>(1) << 46;                          // descriptor version (Blackwell)
  d |= static_cast<uint64_t>(2) << 61;                          // SWIZZLE_128B
  return d;
}

template <int BN>
__global__ void __launch_bounds__(256, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const WParams p) {
  using C = WCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + C::kStages * C::kStage);
  uint64_t* empty_bar = full_bar + C::kStages;
  uint64_t* tmem_full_bar = empty_bar + C::kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);   // (warp-uniform for the compiler)
  const int lane = threadIdx.x & 31;
  const int tile = blockIdx.x % (p.tiles_n * p.tiles_k);
  const int split = blockIdx.x / (p.tiles_n * p.tiles_k);
  const int n0 = (tile / p.tiles_k) * TILE_N;
  const int k0 = (tile % p.tiles_k) * BN;
  const int kb_begin = split * p.kb_per_split;
  const int kb_end = min(p.kb_total, kb_begin + p.kb_per_split);
  const int nkb = max(0, kb_end - kb_begin);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < C::kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    mbar_fence_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, C::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    {                                   // the WHOLE warp, converged: TMA / tcgen05 issue with the election inside the PTX (a4r_common.cuh)
      int stage = 0;
      uint32_t phase = 0;
      for (int i = 0; i < nkb; ++i) {
        const int t0 = (kb_begin + i) * BT;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        __syncwarp();
        uint8_t* sa = smem + stage * C::kStage;
        uint8_t* sb = sa + C::kStageA;
        mbar_expect_tx_elect(&full_bar[stage], C::kStage);
#pragma unroll
        for (int j = 0; j < TILE_N / 64; ++j) tma_load_2d_elect_noarm(&tmA, sa + j * CHUNK_BYTES, &full_bar[stage], n0 + 64 * j, t0);
#pragma unroll
        for (int j = 0; j < BN / 64; ++j) tma_load_2d_elect_noarm(&tmB, sb + j * CHUNK_BYTES, &full_bar[stage], k0 + 64 * j, t0);
        if (++stage == C::kStages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    {
      // both operands MN-major: bits 15 (A) and 16 (B) of the instruction descriptor
      constexpr uint32_t idesc = umma_idesc_bf16(TILE_N, BN) | (1u << 15) | (1u << 16);
      int stage = 0;
      uint32_t phase = 0;
      for (int i = 0; i < nkb; ++i) {
        mbar_wait(&full_bar[stage], phase);
        __syncwarp();
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + stage * C::kStage);
        const uint64_t adesc = umma_desc_mn_sw128(sa);
        const uint64_t bdesc = umma_desc_mn_sw128(sa + C::kStageA);
#pragma unroll
        for (int k = 0; k < BT / 16; ++k)   // 16 token rows = 2048 B = 128 descriptor units
          umma_bf16_ss_elect(tmem_base, adesc + static_cast<uint64_t>(k * 128), bdesc + static_cast<uint64_t>(k * 128), idesc,
                             (i | k) != 0 ? 1u : 0u);
        umma_commit_elect(&empty_bar[stage]);
        if (++stage == C::kStages) {
          stage = 0;
          phase ^= 1;
        }
      }
      umma_commit_elect(tmem_full_bar);
    }
  } else if (warp >= 4) {
    const int quad = warp & 3;
    const int row = n0 + quad * 32 + lane;                 // dW row of this thread
    float* dst = p.partial + (static_cast<int64_t>(split) * p.Npad + row) * p.Kpad + k0;
    if (nkb > 0) {
      mbar_wait(tmem_full_bar, 0);
      tc_fence_after();
    }
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      uint32_t acc[32];
      if (nkb > 0) {
        tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + static_cast<uint32_t>(c * 32), acc);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[j] = 0u;           // an empty split contributes zeros
      }
#pragma unroll
      for (int j = 0; j < 8; ++j)
        st_na_v4(dst + c * 32 + 4 * j, make_uint4(acc[4 * j], acc[4 * j + 1], acc[4 * j + 2], acc[4 * j + 3]));
    }
    tc_fence_before();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 2) tmem_dealloc(tmem_base, C::kTmemCols);
}

// dW[n, k] = alpha * sum_s partial[s, n, k] (+ dW[n, k]); fixed summation order
__global__ void wgrad_tc_reduce_kernel(const float* __restrict__ partial, float* __restrict__ dW, int64_t ldw, int N, int K,
                                       int64_t Npad, int64_t Kpad, int splits, float alpha, int accumulate) {
  const int k4 = K >> 2;
  const int64_t total = static_cast<int64_t>(N) * k4;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int n = static_cast<int>(i / k4), k = static_cast<int>(i % k4) * 4;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int sp = 0; sp < splits; ++sp) {
      const float4 v = *reinterpret_cast<const float4*>(partial + (static_cast<int64_t>(sp) * Npad + n) * Kpad + k);
      s.x += v.x, s.y += v.y, s.z += v.z, s.w += v.w;
    }
    float* o = dW + static_cast<int64_t>(n) * ldw + k;
    if (accumulate) {
      o[0] += alpha * s.x, o[1] += alpha * s.y, o[2] += alpha * s.z, o[3] += alpha * s.w;
    } else {
      o[0] = alpha * s.x, o[1] = alpha * s.y, o[2] = alpha * s.z, o[3] = alpha * s.w;
    }
  }
}

int pick_bn(int64_t K) { return K > 128 ? 256 : (K > 64 ? 128 : 64); }

void plan_tc(int64_t M, int64_t N, int64_t K, WParams* p, int* bn) {
  *bn = pick_bn(K);
  p->N = static_cast<int>(N);
  p->K = static_cast<int>(K);
  p->tiles_n = static_cast<int>((N + TILE_N - 1) / TILE_N);
  p->tiles_k = static_cast<int>((K + *bn - 1) / *bn);
  p->Npad = static_cast<int64_t>(p->tiles_n) * TILE_N;
  p->Kpad = static_cast<int64_t>(p->tiles_k) * *bn;
  p->kb_total = static_cast<int>((M + BT - 1) / BT);
  const int tiles = p->tiles_n * p->tiles_k;
  int splits = a4r_num_sms() / tiles;                    // about one wave of CTAs
  const int max_splits = (p->kb_total + 15) / 16;        // at least 16 k-blocks (1,024 tokens) per split
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  p->kb_per_split = (p->kb_total + splits - 1) / splits;
  if (p->kb_per_split < 1) p->kb_per_split = 1;
  p->splits = splits;
}

template <int BN>
int launch_tc(const CUtensorMap& tmA, const CUtensorMap& tmB, const WParams& p, cudaStream_t stream) {
  static bool attr_done = false;
  if (!attr_done) {
    A4R_CUDA_OK(cudaFuncSetAttribute(wgrad_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, WCfg<BN>::kSmem));
    attr_done = true;
  }
  const int grid = p.tiles_n * p.tiles_k * p.splits;
  wgrad_tc_kernel<BN><<<grid, 256, WCfg<BN>::kSmem, stream>>>(tmA, tmB, p);
  A4R_LAUNCH_OK();
  return A4R_OK;
}

}  // namespace

int a4r_make_tmap_bf16(CUtensorMap* m, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows);

extern "C" size_t a4r_wgrad_tc_workspace_bytes(int64_t M, int64_t N, int64_t K) {
  WParams p;
  int bn;
  plan_tc(M, N, K, &p, &bn);
  return static_cast<size_t>(p.splits) * p.Npad * p.Kpad * sizeof(float);
}

extern "C" int a4r_wgrad_tc_bf16(const void* A, int64_t lda, const void* B, int64_t ldb, float* dW, int64_t ldw, int64_t M,
                                 int64_t N, int64_t K, float alpha, int32_t accumulate, void* workspace,
                                 size_t workspace_bytes, a4r_stream_t stream_) {
  A4R_CHECK_ARG(M >= 0 && N > 0 && K > 0 && N % 8 == 0 && K % 8 == 0, "wgrad_tc: N and K must be positive multiples of 8");
  A4R_CHECK_ARG(dW && (M == 0 || (A && B)), "wgrad_tc: NULL pointer");   // empty operands may legitimately be NULL
  A4R_CHECK_ARG(M < (1ll << 31) && N < (1ll << 24) && K < (1ll << 24), "wgrad_tc: dims out of range");
  A4R_CHECK_ARG(lda >= N && ldb >= K && lda % 8 == 0 && ldb % 8 == 0 && ldw >= K && ldw % 4 == 0, "wgrad_tc: bad leading dimensions");
  A4R_CHECK_ARG(a4r_aligned16(A) && a4r_aligned16(B) && a4r_aligned16(dW), "wgrad_tc: A, B and dW must be 16B aligned");
  const size_t need = a4r_wgrad_tc_workspace_bytes(M, N, K);
  if (workspace == nullptr || workspace_bytes < need)
    return a4r_set_error(A4R_EWORKSPACE, "wgrad_tc: workspace too small (%zu < %zu)", workspace_bytes, need);
  A4R_CHECK_ARG(a4r_aligned16(workspace), "wgrad_tc: workspace must be 16B aligned");
  int rc = a4r_device_check();
  if (rc != A4R_OK) return rc;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  WParams p;
  int bn;
  plan_tc(M, N, K, &p, &bn);
  p.partial = static_cast<float*>(workspace);
  if (M > 0) {
    CUtensorMap tmA, tmB;
    if ((rc = a4r_make_tmap_bf16(&tmA, A, M, N, lda, BT)) != A4R_OK) return rc;
    if ((rc = a4r_make_tmap_bf16(&tmB, B, M, K, ldb, BT)) != A4R_OK) return rc;
    if (bn == 256) rc = launch_tc<256>(tmA, tmB, p, stream);
    else if (bn == 128) rc = launch_tc<128>(tmA, tmB, p, stream);
    else rc = launch_tc<64>(tmA, tmB, p, stream);
    if (rc != A4R_OK) return rc;
    a4r_count_launch(1);
  } else {
    p.splits = 0;   // nothing to sum: dW = 0 (or unchanged when accumulating)
  }
  const int64_t total = N * (K / 4);
  int64_t blocks = (total + 255) / 256;
  const int64_t cap = static_cast<int64_t>(a4r_num_sms()) * 8;
  if (blocks > cap) blocks = cap;
  wgrad_tc_reduce_kernel<<<static_cast<int>(blocks), 256, 0, stream>>>(p.partial, dW, ldw, p.N, p.K, p.Npad, p.Kpad, p.splits,
                                                                      alpha, accumulate);
  A4R_LAUNCH_OK();
  a4r_count_launch(1);
  return A4R_OK;
}
