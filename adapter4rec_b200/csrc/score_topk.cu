// score_topk.cu — K10..K13: the full-ranking evaluator.
//
//   a4r_gather_rows     K10  item-ID gather  E[ids] -> [B,S,D]        (BuildEvalDataset.__getitem__, dataset.py:65-78)
//   a4r_score_topk      K11 + K12  scores = U·Eᵀ on tcgen05 with the history mask and a per-row top-k selection
//                       fused into the TMEM epilogue: the [users x items] score matrix never exists in HBM
//                       (eval_model, metrics.py:105-111; the reference materialises [B, I+1] fp32 and argsorts it)
//   a4r_topk_merge      K13  merge of partial top-k lists (from item splits on one GPU and from the item shards
//                       of all GPUs after the all-gather) + HR@k / NDCG@k per user (metrics_topK, metrics.py:51-59)
//
// Total order everywhere: (score desc, item id asc).  Item id 0 (the padding item) and the user's history ids are
// excluded, as `score[history] = -inf; score = score[1:]` does in the reference.
//
// score_topk mapping: M = users (TMEM lanes: one epilogue thread owns one user), N = items.  A CTA PAIR (cta_group::2, the
// two SMs of a TPC) owns one block of 256 users and one contiguous slice of the item shard; it streams item tiles of 256
// rows through the same TMA -> smem -> tcgen05.mma -> TMEM pipeline as gemm_sm100.cu — each CTA stages its own 128 users
// and HALF of the item rows of a tile, the pair's MMA reads both halves, which cuts shared-memory and L2->SM operand
// traffic per CTA by a third against the single-CTA 128 x 256 tile — while its epilogue threads keep a sorted top-k
// list in registers.  Work per launch = ceil(U/256) x splits pairs; each (split, epilogue group) emits one partial list.
#include "a4r_common.cuh"

namespace {

constexpr int BM = 128, BN = 256, BK = 64, UMMA_K = 16;
constexpr int CG = 2;                // CTAs per MMA
constexpr int STAGES = 6;
constexpr int EPI_GROUPS = 2;
constexpr int THREADS = 128 + 128 * EPI_GROUPS;
constexpr int STAGE_A = BM * BK * 2, STAGE_B = (BN / CG) * BK * 2, STAGE_BYTES = STAGE_A + STAGE_B;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
constexpr int TMEM_COLS = 2 * BN;
constexpr int KMAX = 16;  // top-k capacity of the register list

struct TopkParams {
  const int32_t* history;  // [U, hist_len] item ids (0 = empty slot) or NULL
  float* out_scores;       // [P, U, k]
  int32_t* out_ids;        // [P, U, k]
  int64_t id_base;         // item id of row 0 of this shard
  int U, I, hist_len, k;
  int nk;                  // k-blocks (d / 64, rounded up)
  int m_blocks, splits, tiles_per_split;   // m_blocks = blocks of 256 users (one per CTA pair)
};

__global__ void __launch_bounds__(THREADS, 1)
score_topk_kernel(const __grid_constant__ CUtensorMap tmU, const __grid_constant__ CUtensorMap tmE, const TopkParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;   // (warp-uniform for the compiler)
  const uint32_t cta_rank = cluster_ctarank();
  const bool leader = cta_rank == 0;
  const int pair = static_cast<int>(blockIdx.x >> 1);
  const int m_blk = pair % p.m_blocks;
  const int split = pair / p.m_blocks;
  const int total_tiles = (p.I + BN - 1) / BN;
  const int t_begin = split * p.tiles_per_split;
  int t_end = t_begin + p.tiles_per_split;
  if (t_end > total_tiles) t_end = total_tiles;
  const int m0 = m_blk * (BM * CG) + static_cast<int>(cta_rank) * BM;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmU);
    tma_prefetch_desc(&tmE);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full_bar[s], 1);
      mbar_init(&tmem_empty_bar[s], 4 * EPI_GROUPS * CG);   // one arrive per epilogue warp of both CTAs (leader's copy)
    }
    mbar_fence_init();
  }
  if (warp == 2) {
    tmem_alloc2(tmem_slot, TMEM_COLS);
    tmem_relinquish2();
  }
  tc_fence_before();
  cluster_sync_all();   // barriers of BOTH CTAs are initialised past this point
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    {                                   // the WHOLE warp, converged: TMA / tcgen05 issue with the election inside the PTX (a4r_common.cuh)
      int stage = 0;
      uint32_t phase = 0;
      for (int t = t_begin; t < t_end; ++t) {
        for (int kb = 0; kb < p.nk; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          __syncwarp();
          uint8_t* sa = smem + stage * STAGE_BYTES;
          // both CTAs' bytes are credited to the LEADER's full barrier
          if (leader) mbar_expect_tx_elect(&full_bar[stage], STAGE_BYTES * CG);
          const uint32_t bar = mapa_u32(smem_u32(&full_bar[stage]), 0);
          tma_load_2d_2cta_elect(&tmU, sa, bar, kb * BK, m0);
          tma_load_2d_2cta_elect(&tmE, sa + STAGE_A, bar, kb * BK, t * BN + static_cast<int>(cta_rank) * (BN / CG));
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (leader) {
      constexpr uint32_t idesc = umma_idesc_bf16(BM * CG, BN);
      int stage = 0, as = 0;
      uint32_t phase = 0, aphase = 0;
      for (int t = t_begin; t < t_end; ++t) {
        mbar_wait(&tmem_empty_bar[as], aphase ^ 1);
        __syncwarp();
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(as * BN);
        for (int kb = 0; kb < p.nk; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          __syncwarp();
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
          const uint64_t adesc = umma_desc_k_sw128(sa), bdesc = umma_desc_k_sw128(sa + STAGE_A);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k)
            umma_bf16_ss_2cta_elect(tmem_d, adesc + static_cast<uint64_t>(k * 2), bdesc + static_cast<uint64_t>(k * 2), idesc,
                                    (kb | k) != 0 ? 1u : 0u);
          umma_commit_2cta_elect(&empty_bar[stage], 3);   // frees the slot in both CTAs of the pair
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit_2cta_elect(&tmem_full_bar[as], 3);
        if (++as == 2) {
          as = 0;
          aphase ^= 1;
        }
      }
    }
  } else if (warp >= 4) {
    const int quad = warp & 3, group = (warp - 4) >> 2;
    const int user = m0 + quad * 32 + lane;
    const bool user_ok = user < p.U;
    // sorted (score desc, id asc) list in registers; slots >= k stay at -inf and are never written out
    float ls[KMAX];
    int32_t li[KMAX];
#pragma unroll
    for (int i = 0; i < KMAX; ++i) {
      ls[i] = -INFINITY;
      li[i] = 0;
    }
    float thr = -INFINITY;  // score of the current k-th entry
    const int32_t* hist = (p.history != nullptr && user_ok) ? p.history + static_cast<int64_t>(user) * p.hist_len : nullptr;
    int as = 0;
    uint32_t aphase = 0;
    for (int t = t_begin; t < t_end; ++t) {
      mbar_wait(&tmem_full_bar[as], aphase);
      tc_fence_after();
#pragma unroll 1
      for (int c = group; c < BN / 32; c += EPI_GROUPS) {
        const int col0 = t * BN + c * 32;
        if (col0 >= p.I) break;
        uint32_t acc[32];
        tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + static_cast<uint32_t>(as * BN + c * 32), acc);
        tmem_ld_wait();
        // cheap pre-filter: which of the 32 scores beat the current threshold?  (registers only: the common case ends here)
        uint32_t cand = 0;
#pragma unroll
        for (int j = 0; j < 32; ++j) cand |= (__uint_as_float(acc[j]) > thr ? 1u : 0u) << j;
        if (!user_ok) cand = 0;
        if (__any_sync(0xffffffffu, cand != 0)) {
          // rare path: only now do the scores go to (thread-local) memory for dynamic indexing; each lane then visits
          // only its own candidates, in ascending column order
          float sc[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) sc[j] = __uint_as_float(acc[j]);
          while (cand != 0) {
            const int j = __ffs(cand) - 1;
            cand &= cand - 1;
            const float s = sc[j];
            if (!(s > thr)) continue;  // ids arrive in ascending order: an equal score never displaces an entry
            const int col = col0 + j;
            if (col >= p.I) break;
            const int64_t id64 = p.id_base + col;
            if (id64 == 0) continue;  // padding item
            const int32_t id = static_cast<int32_t>(id64);
            bool seen = false;
            if (hist != nullptr)
              for (int h = 0; h < p.hist_len; ++h) seen |= (__ldg(hist + h) == id);
            if (seen) continue;
            // insert after every entry with score >= s (stable for ties), dropping the last
#pragma unroll
            for (int i = KMAX - 1; i >= 0; --i) {
              if (i >= p.k) continue;
              const bool shift = i > 0 && ls[i - 1] < s;  // entry i-1 moves down to i
              const bool place = (ls[i] < s) && !(i > 0 && ls[i - 1] < s);
              if (shift) {
                ls[i] = ls[i - 1];
                li[i] = li[i - 1];
              } else if (place) {
                ls[i] = s;
                li[i] = id;
              }
            }
            thr = ls[KMAX - 1];
#pragma unroll
            for (int i = 0; i < KMAX; ++i)
              if (i == p.k - 1) thr = ls[i];
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(mapa_u32(smem_u32(&tmem_empty_bar[as]), 0));   // the leader's barrier
      if (++as == 2) {
        as = 0;
        aphase ^= 1;
      }
    }
    if (user_ok) {
      const int64_t part = static_cast<int64_t>(split) * EPI_GROUPS + group;
      float* os = p.out_scores + (part * p.U + user) * p.k;
      int32_t* oi = p.out_ids + (part * p.U + user) * p.k;
#pragma unroll
      for (int i = 0; i < KMAX; ++i)
        if (i < p.k) {
          os[i] = ls[i];
          oi[i] = li[i];
        }
    }
  }

  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  if (warp == 2) tmem_dealloc2(tmem_base, TMEM_COLS);
}

// K13: one thread per user merges P partial lists under (score desc, id asc); optional HR/NDCG against `target`.
__global__ void topk_merge_kernel(const float* __restrict__ in_scores, const int32_t* __restrict__ in_ids, int P, int U,
                                  int k, float* __restrict__ out_scores, int32_t* __restrict__ out_ids,
                                  const int32_t* __restrict__ target, float* __restrict__ hit, float* __restrict__ ndcg) {
  const int u = blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= U) return;
  int head[64];  // next unread entry of each partial list (each is already sorted)
  for (int q = 0; q < P; ++q) head[q] = 0;
  const int32_t tgt = target != nullptr ? target[u] : -1;
  float h = 0.0f, nd = 0.0f;
  for (int r = 0; r < k; ++r) {
    int best = -1;
    float bs = -INFINITY;
    int32_t bi = 0;
    for (int q = 0; q < P; ++q) {
      if (head[q] >= k) continue;
      const int64_t off = (static_cast<int64_t>(q) * U + u) * k + head[q];
      const float s = in_scores[off];
      const int32_t id = in_ids[off];
      if (s == -INFINITY) continue;
      if (best < 0 || s > bs || (s == bs && id < bi)) {
        best = q;
        bs = s;
        bi = id;
      }
    }
    if (best >= 0) head[best]++;
    out_scores[static_cast<int64_t>(u) * k + r] = best >= 0 ? bs : -INFINITY;
    out_ids[static_cast<int64_t>(u) * k + r] = best >= 0 ? bi : 0;
    if (best >= 0 && bi == tgt) {
      h = 1.0f;
      nd = 1.0f / log2f(static_cast<float>(r) + 2.0f);
    }
  }
  if (hit != nullptr) hit[u] = h;
  if (ndcg != nullptr) ndcg[u] = nd;
}

// K10: out[r, :] = table[ids[r], :]   (bf16 rows of D elements, 16-byte vectors)
__global__ void gather_rows_kernel(const __nv_bfloat16* __restrict__ table, const int64_t* __restrict__ ids,
                                   __nv_bfloat16* __restrict__ out, int64_t rows, int D) {
  const int chunks = D >> 3;
  const int64_t total = rows * chunks;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = i / chunks;
    const int c = static_cast<int>(i % chunks);
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(table + ids[r] * D + c * 8));
    *reinterpret_cast<uint4*>(out + r * D + c * 8) = v;
  }
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int make_tmap(CUtensorMap* m, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
  static PFN_encodeTiled fn = nullptr;
  if (fn == nullptr) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      return a4r_set_error(A4R_ECUDA, "cuTensorMapEncodeTiled entry point not found");
    fn = reinterpret_cast<PFN_encodeTiled>(ptr);
  }
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 2};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(BK), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return a4r_set_error(A4R_ECUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  return A4R_OK;
}

void plan(int64_t U, int64_t I, int* m_blocks, int* splits, int* tiles_per_split) {
  *m_blocks = static_cast<int>((U + BM * CG - 1) / (BM * CG));
  const int tiles = static_cast<int>((I + BN - 1) / BN);
  int s = (a4r_num_sms() / CG) / *m_blocks;   // CTA pairs resident at once
  if (s < 1) s = 1;
  if (s > 16) s = 16;  // the merge kernel handles at most 64 partial lists (splits x 2 groups x shards)
  if (s > tiles) s = tiles;
  *tiles_per_split = (tiles + s - 1) / s;
  *splits = (tiles + *tiles_per_split - 1) / *tiles_per_split;
}

}  // namespace

extern "C" int a4r_score_topk_partials(int64_t U, int64_t I) {
  int mb, sp, tps;
  plan(U, I, &mb, &sp, &tps);
  return sp * EPI_GROUPS;
}

extern "C" int a4r_score_topk(const void* users, int64_t ld_users, const void* items, int64_t ld_items, int64_t U,
                              int64_t I, int64_t d, int64_t id_base, const int32_t* history, int64_t hist_len,
                              int32_t k, float* out_scores, int32_t* out_ids, a4r_stream_t stream_) {
  A4R_CHECK_ARG(users && items && out_scores && out_ids, "score_topk: NULL pointer");
  A4R_CHECK_ARG(U >= 1 && I >= 1 && d >= 8 && d % 8 == 0, "score_topk: bad U/I/d");
  A4R_CHECK_ARG(U < (1ll << 31) && I < (1ll << 31) && id_base >= 0 && id_base + I < (1ll << 31), "score_topk: ids must fit int32");
  A4R_CHECK_ARG(k >= 1 && k <= KMAX, "score_topk: k must be in [1,%d]", KMAX);
  A4R_CHECK_ARG(ld_users >= d && ld_items >= d && ld_users % 8 == 0 && ld_items % 8 == 0, "score_topk: bad leading dims");
  A4R_CHECK_ARG(a4r_aligned16(users) && a4r_aligned16(items), "score_topk: pointers must be 16B aligned");
  A4R_CHECK_ARG(history == nullptr || hist_len >= 1, "score_topk: hist_len must be >= 1 when history is given");
  int rc = a4r_device_check();
  if (rc != A4R_OK) return rc;
  CUtensorMap tmU, tmE;
  if ((rc = make_tmap(&tmU, users, U, d, ld_users, BM)) != A4R_OK) return rc;
  if ((rc = make_tmap(&tmE, items, I, d, ld_items, BN / CG)) != A4R_OK) return rc;
  TopkParams p;
  p.history = history;
  p.out_scores = out_scores;
  p.out_ids = out_ids;
  p.id_base = id_base;
  p.U = static_cast<int>(U);
  p.I = static_cast<int>(I);
  p.hist_len = static_cast<int>(hist_len);
  p.k = k;
  p.nk = static_cast<int>((d + BK - 1) / BK);
  plan(U, I, &p.m_blocks, &p.splits, &p.tiles_per_split);
  static bool attr_done = false;
  if (!attr_done) {
    A4R_CUDA_OK(cudaFuncSetAttribute(score_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    attr_done = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(p.m_blocks * p.splits * CG);
  cfg.blockDim = dim3(THREADS);
  cfg.dynamicSmemBytes = SMEM_BYTES;
  cfg.stream = static_cast<cudaStream_t>(stream_);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  A4R_CUDA_OK(cudaLaunchKernelEx(&cfg, score_topk_kernel, tmU, tmE, p));
  A4R_LAUNCH_OK();
  a4r_count_launch(1);
  return A4R_OK;
}

extern "C" int a4r_topk_merge(const float* in_scores, const int32_t* in_ids, int32_t P, int64_t U, int32_t k,
                              float* out_scores, int32_t* out_ids, const int32_t* target, float* hit, float* ndcg,
                              a4r_stream_t stream_) {
  A4R_CHECK_ARG(in_scores && in_ids && out_scores && out_ids, "topk_merge: NULL pointer");
  A4R_CHECK_ARG(P >= 1 && P <= 64, "topk_merge: P must be in [1,64] (got %d)", P);
  A4R_CHECK_ARG(U >= 1 && U < (1ll << 31) && k >= 1 && k <= KMAX, "topk_merge: bad U/k");
  int rc = a4r_device_check();
  if (rc != A4R_OK) return rc;
  topk_merge_kernel<<<static_cast<int>((U + 127) / 128), 128, 0, static_cast<cudaStream_t>(stream_)>>>(
      in_scores, in_ids, P, static_cast<int>(U), k, out_scores, out_ids, target, hit, ndcg);
  A4R_LAUNCH_OK();
  a4r_count_launch(1);
  return A4R_OK;
}

extern "C" int a4r_gather_rows(const void* table, const int64_t* ids, void* out, int64_t rows, int64_t D,
                               a4r_stream_t stream_) {
  A4R_CHECK_ARG(table && ids && out, "gather_rows: NULL pointer");
  A4R_CHECK_ARG(rows >= 0 && D >= 8 && D % 8 == 0, "gather_rows: D must be a positive multiple of 8");
  A4R_CHECK_ARG(a4r_aligned16(table) && a4r_aligned16(out), "gather_rows: pointers must be 16B aligned");
  int rc = a4r_device_check();
  if (rc != A4R_OK) return rc;
  if (rows == 0) return A4R_OK;
  const int64_t total = rows * (D / 8);
  int64_t blocks = (total + 255) / 256;
  const int64_t cap = static_cast<int64_t>(a4r_num_sms()) * 16;
  if (blocks > cap) blocks = cap;
  gather_rows_kernel<<<static_cast<int>(blocks), 256, 0, static_cast<cudaStream_t>(stream_)>>>(
      static_cast<const __nv_bfloat16*>(table), ids, static_cast<__nv_bfloat16*>(out), rows, static_cast<int>(D));
  A4R_LAUNCH_OK();
  a4r_count_launch(1);
  return A4R_OK;
}
