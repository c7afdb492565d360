// attention_tc_bwd_sm100.cu — backward of the unmasked mid-length attention (32 < L <= 256, head_dim 64) on tcgen05.
//
//   P = exp(q kᵀ/sqrt(d) - lse)      dV = Pᵀ dO      dP = dO Vᵀ      dS = P ∘ (dP - D) / sqrt(d),  D = rowsum(dO ∘ O)
//   dQ = dS K      dK = dSᵀ Q         transformers' ViTSelfAttention backward (Downstream/CV/model/encoders.py:31-32 forward)
//
// One work unit = one (image, head); a persistent CTA walks the units.  A unit is processed in up to four STEPS (key tile j outer,
// query tile i inner, 128 x 128 each); per step five tensor-core contractions run, all operands in shared memory:
//   M1  S_ij  = Q_i K_jᵀ        (A, B K-major)                          -> tensor memory columns [0, 128)
//   M2  dP_ij = dO_i V_jᵀ       (A, B K-major)                          -> [128, 256)
//        the 8 compute warps (two threads per query row, 64 keys each) read S and dP once, form P and dS in registers
//        (one FFMA + MUFU.EX2 per score: the forward's log-sum-exp makes the softmax a pure elementwise map), and write both as
//        bf16 into two [128 q x 128 keys] SWIZZLE_128B shared-memory tiles
//   M3  dV_j += P_ijᵀ dO_i      (A = the P tile read MN-major, i.e. transposed by the descriptor; B = dO_i MN-major)  -> [256, 320)
//   M4  dK_j += dS_ijᵀ Q_i      (A = the dS tile MN-major; B = Q_i MN-major)                                        -> [320, 384)
//   M5  dQ_i += dS_ij K_j       (A = the dS tile K-major;  B = K_j MN-major)                                        -> [384 + 64 i, +64)
// No transposed copy of anything is ever made: the same TMA box serves as K-major and as MN-major operand.
// Inputs stream through two TMA rings — (Q_i | dO_i) pairs, 3 slots, and (K_j | V_j) pairs, 2 slots — so the next unit's tiles
// are in flight while the current unit computes; the issuer runs M1/M2 of step s+1 BEFORE M3-M5 of step s, so the tensor core
// works on one step while the compute warps work on the next.
#include "attn_tc_common.cuh"

namespace {

constexpr int CWARPS = 8;                           // compute warps: two threads per query row, 64 keys each
constexpr int BWD_THREADS = 128 + CWARPS * 32;
constexpr int QD_SLOTS = 3, KV_SLOTS = 2;
constexpr int PAIR = 2 * BOX;                       // 32 KB: two [128 x 64] tiles
constexpr int PDS_OFF = (QD_SLOTS + KV_SLOTS) * PAIR;       // P tile (32 KB) then dS tile (32 KB)
constexpr int BAR_OFF = PDS_OFF + 4 * BOX;
constexpr int BWD_SMEM = BAR_OFF + 32 * 8 + 16;
constexpr int COL_S = 0, COL_DP = 128, COL_DV = 256, COL_DK = 320, COL_DQ = 384;
constexpr float LOG2E_F = 1.4426950408889634f;

struct BwdParams {
  const __nv_bfloat16* dout;
  const __nv_bfloat16* ctx;
  const float* lse;
  __nv_bfloat16* dqkv;
  int64_t ld_qkv, ld_out;
  int N, L, Lk, heads, nq, nkj;
  float scale, c;
};

// MN-major SW128 operand spanning TWO 64-wide mn chunks 16 KB apart (M = 128 rows of the transposed tile)
A4R_DEVICE uint64_t umma_desc_mn_sw128_2chunk(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(BOX >> 4) << 16;                   // LBO: the second 64-wide mn chunk
  d |= static_cast<uint64_t>(1024 >> 4) << 32;                  // SBO: next group of 8 k-rows
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

// {lse * log2(e), scale * sum_d dO[r, d] O[r, d]} per (token row, head), written as one float2 into the first 8 bytes of the row's
// dq slot of dqkv (which the main kernel reads before it overwrites the slot with dQ): no workspace, 8 bytes per (row, head).
// One warp per token row; 8 lanes share a head (64 columns = 8 chunks of 8).
template <int ITERS>
__global__ void __launch_bounds__(256)
attn_row_info_kernel(const __nv_bfloat16* __restrict__ ctx, const __nv_bfloat16* __restrict__ dout, const float* __restrict__ lse,
                     __nv_bfloat16* __restrict__ dqkv, int64_t rows, int heads, int64_t ld_out, int64_t ld_qkv, float scale) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  const int nchunk = heads * 8;
  for (int64_t row = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; row < rows; row += warps) {
    // every load of the row is issued before the first use (the kernel is one pass over two [rows, H] tensors: HBM-bound, and
    // with a load-use-store chain per 32-chunk group it kept too few bytes in flight: 0.61 of the HBM peak)
    uint4 a[ITERS], b[ITERS];
    float l[ITERS];
#pragma unroll
    for (int i = 0; i < ITERS; ++i) {
      const int c = i * 32 + lane;                   // 8-column chunk of the row; head = c / 8
      a[i] = b[i] = make_uint4(0u, 0u, 0u, 0u);
      l[i] = 0.0f;
      if (c < nchunk) {
        a[i] = ld_nc_v4(ctx + row * ld_out + c * 8);
        b[i] = ld_nc_v4(dout + row * ld_out + c * 8);
        if ((lane & 7) == 0) l[i] = __ldg(lse + row * heads + (c >> 3));
      }
    }
#pragma unroll
    for (int i = 0; i < ITERS; ++i) {
      const int c = i * 32 + lane;
      const uint32_t aw[4] = {a[i].x, a[i].y, a[i].z, a[i].w}, bw[4] = {b[i].x, b[i].y, b[i].z, b[i].w};
      float acc = 0.0f;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 x = bf16x2_to_f2(aw[e]), y = bf16x2_to_f2(bw[e]);
        acc = fmaf(x.x, y.x, fmaf(x.y, y.y, acc));
      }
      acc += __shfl_xor_sync(0xffffffffu, acc, 1);
      acc += __shfl_xor_sync(0xffffffffu, acc, 2);
      acc += __shfl_xor_sync(0xffffffffu, acc, 4);
      if ((lane & 7) == 0 && c < nchunk)
        *reinterpret_cast<float2*>(dqkv + row * ld_qkv + (c >> 3) * DH) = make_float2(l[i] * LOG2E_F, acc * scale);
    }
  }
}

#ifdef A4R_ATTN_TIMING
__device__ long long g_attn_bwd_timing[16];
#define BSTAMP(var) const long long var = clock64()
#define BACC(i, a, b) acc_t[i] += (b) - (a)
#else
#define BSTAMP(var)
#define BACC(i, a, b)
#endif

__global__ void __launch_bounds__(BWD_THREADS, 1)
attn_vit_tc_bwd_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmDO, const BwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* s_qd = smem;                              // [3] x (Q_i | dO_i)
  uint8_t* s_kv = smem + QD_SLOTS * PAIR;            // [2] x (K_j | V_j)
  uint8_t* s_p = smem + PDS_OFF;                     // [2 key chunks][128 q][128 B]
  uint8_t* s_ds = s_p + 2 * BOX;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + BAR_OFF);
  uint64_t* qd_full = bars;             // [3]
  uint64_t* qd_empty = bars + 3;        // [3]
  uint64_t* kv_full = bars + 6;         // [2]
  uint64_t* kv_empty = bars + 8;        // [2]
  uint64_t* sdp_full = bars + 10;       // M1, M2 of a step have retired
  uint64_t* sdp_free = bars + 11;       // the compute warps have read S and dP out of tensor memory
  uint64_t* pds_ready = bars + 12;      // P and dS of a step are in shared memory
  uint64_t* pds_free = bars + 13;       // M3-M5 of a step have retired (the P / dS tiles may be overwritten)
  uint64_t* dvk_full = bars + 14;       // dV_j, dK_j are complete
  uint64_t* dvk_free = bars + 15;
  uint64_t* dq_full = bars + 16;        // dQ_0, dQ_1 are complete
  uint64_t* dq_free = bars + 17;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 18);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;   // (warp-uniform for the compiler)
  const int units = p.N * p.heads;
  const int H = p.heads * DH;
  const int my_units = (units - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQKV);
    tma_prefetch_desc(&tmDO);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < QD_SLOTS; ++s) {
      mbar_init(&qd_full[s], 1);
      mbar_init(&qd_empty[s], 1);
    }
    for (int s = 0; s < KV_SLOTS; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 1);
    }
    mbar_init(sdp_full, 1);
    mbar_init(sdp_free, CWARPS);
    mbar_init(pds_ready, CWARPS);
    mbar_init(pds_free, 1);
    mbar_init(dvk_full, 1);
    mbar_init(dvk_free, CWARPS);
    mbar_init(dq_full, 1);
    mbar_init(dq_free, CWARPS);
    mbar_fence_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (warp == 0) {
      // ============================== TMA producer (whole warp, converged): pairs in the order the steps consume them ==============================
      uint32_t nqd = 0, nkv = 0;
      auto load_qd = [&](int img, int head, int i) {
        const uint32_t slot = nqd % QD_SLOTS;
        mbar_wait(&qd_empty[slot], ((nqd / QD_SLOTS) & 1u) ^ 1u);
        __syncwarp();
        mbar_expect_tx_elect(&qd_full[slot], PAIR);
        tma_load_3d_elect(&tmQKV, s_qd + slot * PAIR, &qd_full[slot], head * DH, i * QT, img);
        tma_load_3d_elect(&tmDO, s_qd + slot * PAIR + BOX, &qd_full[slot], head * DH, i * QT, img);
        ++nqd;
      };
      auto load_kv = [&](int img, int head, int j) {
        const uint32_t slot = nkv % KV_SLOTS;
        mbar_wait(&kv_empty[slot], ((nkv / KV_SLOTS) & 1u) ^ 1u);
        __syncwarp();
        mbar_expect_tx_elect(&kv_full[slot], PAIR);
        tma_load_3d_elect(&tmQKV, s_kv + slot * PAIR, &kv_full[slot], H + head * DH, j * QT, img);
        tma_load_3d_elect(&tmQKV, s_kv + slot * PAIR + BOX, &kv_full[slot], 2 * H + head * DH, j * QT, img);
        ++nkv;
      };
      for (int u = blockIdx.x; u < units; u += gridDim.x) {
        const int img = u / p.heads, head = u - img * p.heads;
        load_qd(img, head, 0);
        load_kv(img, head, 0);
        if (p.nq > 1) load_qd(img, head, 1);
        if (p.nkj > 1) load_kv(img, head, 1);
      }
    } else if (warp == 1) {
      // ============================== MMA issuer: the WHOLE warp, converged (see umma_*_elect) ==============================
      // steps of a unit: (j, i) lexicographic; global step index g; M12(g + 1) is issued before M345(g)
      const int steps_per_unit = p.nq * p.nkj;
      const int total_steps = my_units * steps_per_unit;
      const uint32_t idesc_kk_full = umma_idesc_bf16(QT, 128);
      const uint32_t idesc_o = umma_idesc_bf16(QT, DH);
      auto step_ij = [&](int g, int& it, int& i, int& j) {
        it = g / steps_per_unit;
        const int r = g - it * steps_per_unit;
        j = r / p.nq;
        i = r - j * p.nq;
      };
      auto issue_m12 = [&](int g) {
        int it, i, j;
        step_ij(g, it, i, j);
        const uint32_t nqd = static_cast<uint32_t>(it * p.nq + i), nkv = static_cast<uint32_t>(it * p.nkj + j);
        if (j == 0) mbar_wait(&qd_full[nqd % QD_SLOTS], (nqd / QD_SLOTS) & 1u);       // first use of the (Q_i | dO_i) pair
        if (i == 0) mbar_wait(&kv_full[nkv % KV_SLOTS], (nkv / KV_SLOTS) & 1u);       // first use of the (K_j | V_j) pair
        if (g > 0) mbar_wait(sdp_free, static_cast<uint32_t>(g - 1) & 1u);            // S, dP of the previous step were read
        __syncwarp();
        tc_fence_after();
        const uint32_t qd = smem_u32(s_qd + (nqd % QD_SLOTS) * PAIR), kv = smem_u32(s_kv + (nkv % KV_SLOTS) * PAIR);
        const int ncols = min(128, p.Lk - j * QT);
        const uint32_t idesc = ncols == 128 ? idesc_kk_full : umma_idesc_bf16(QT, static_cast<uint32_t>(ncols));
        const uint64_t aq = umma_desc_k_sw128(qd), bk = umma_desc_k_sw128(kv);
        const uint64_t ado = umma_desc_k_sw128(qd + BOX), bv = umma_desc_k_sw128(kv + BOX);
#pragma unroll
        for (int k = 0; k < DH / 16; ++k)
          umma_bf16_ss_elect(tmem_base + COL_S, aq + static_cast<uint64_t>(2 * k), bk + static_cast<uint64_t>(2 * k), idesc, k != 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < DH / 16; ++k)
          umma_bf16_ss_elect(tmem_base + COL_DP, ado + static_cast<uint64_t>(2 * k), bv + static_cast<uint64_t>(2 * k), idesc, k != 0 ? 1u : 0u);
        umma_commit_elect(sdp_full);
      };
      auto issue_m345 = [&](int g) {
        int it, i, j;
        step_ij(g, it, i, j);
        const uint32_t nqd = static_cast<uint32_t>(it * p.nq + i), nkv = static_cast<uint32_t>(it * p.nkj + j);
        mbar_wait(pds_ready, static_cast<uint32_t>(g) & 1u);
        const int jg = it * p.nkj + j;                      // global key-group index: dV_j / dK_j accumulators
        if (i == 0 && jg > 0) mbar_wait(dvk_free, static_cast<uint32_t>(jg - 1) & 1u);
        if (j == 0 && i == 0 && it > 0) mbar_wait(dq_free, static_cast<uint32_t>(it - 1) & 1u);
        __syncwarp();
        tc_fence_after();
        const uint32_t qd = smem_u32(s_qd + (nqd % QD_SLOTS) * PAIR), kv = smem_u32(s_kv + (nkv % KV_SLOTS) * PAIR);
        const int rows = min(128, p.L - i * QT), ncols = min(128, p.Lk - j * QT);
        const int kq = (rows + 15) / 16;                    // k-steps over the query rows of this tile (rows past L are zero)
        const int kk = ncols / 16;                          // k-steps over the keys of this tile
        // M3: dV_j += P^T dO_i      M4: dK_j += dS^T Q_i      (A MN-major = transposed tile, B MN-major)
        const uint32_t idesc_t = idesc_o | (1u << 15) | (1u << 16);
        const uint64_t ap = umma_desc_mn_sw128_2chunk(smem_u32(s_p)), ads = umma_desc_mn_sw128_2chunk(smem_u32(s_ds));
        const uint64_t bdo = umma_desc_mn_sw128_1chunk(qd + BOX), bq = umma_desc_mn_sw128_1chunk(qd);
        for (int k = 0; k < kq; ++k)
          umma_bf16_ss_elect(tmem_base + COL_DV, ap + static_cast<uint64_t>(128 * k), bdo + static_cast<uint64_t>(128 * k), idesc_t,
                             (i | k) != 0 ? 1u : 0u);
        for (int k = 0; k < kq; ++k)
          umma_bf16_ss_elect(tmem_base + COL_DK, ads + static_cast<uint64_t>(128 * k), bq + static_cast<uint64_t>(128 * k), idesc_t,
                             (i | k) != 0 ? 1u : 0u);
        // M5: dQ_i += dS K_j        (A K-major: two 64-key chunks; B = K_j MN-major)
        const uint32_t idesc_q = idesc_o | (1u << 16);
        const uint64_t bkm = umma_desc_mn_sw128_1chunk(kv);
        for (int k = 0; k < kk; ++k) {
          const uint64_t a = umma_desc_k_sw128(smem_u32(s_ds) + (k >> 2) * BOX) + static_cast<uint64_t>(2 * (k & 3));
          umma_bf16_ss_elect(tmem_base + COL_DQ + i * DH, a, bkm + static_cast<uint64_t>(128 * k), idesc_q, (j | k) != 0 ? 1u : 0u);
        }
        umma_commit_elect(pds_free);
        if (i == p.nq - 1) {                                 // last query tile of this key group
          umma_commit_elect(dvk_full);
          umma_commit_elect(&kv_empty[nkv % KV_SLOTS]);
        }
        if (j == p.nkj - 1) umma_commit_elect(&qd_empty[nqd % QD_SLOTS]);   // last key group: the (Q_i | dO_i) pair is dead
        if (j == p.nkj - 1 && i == p.nq - 1) umma_commit_elect(dq_full);
      };
      if (total_steps > 0) issue_m12(0);
      for (int g = 0; g < total_steps; ++g) {
        if (g + 1 < total_steps) issue_m12(g + 1);
        issue_m345(g);
      }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
    // ============================== compute warps: two threads per query row ==============================
    const int cw = warp - 4;                        // 0..7
    const int quad = cw & 3, half = cw >> 2;        // TMEM lane quadrant; 64-key chunk of the 128-key tile
    const uint32_t lane_addr = static_cast<uint32_t>(quad * 32) << 16;
    const int rl = quad * 32 + lane;                // row within a 128-row tile = TMEM lane
    const uint32_t sp = smem_u32(s_p), sds = smem_u32(s_ds);
    const float nscale = p.scale;

    // {lse * log2(e), D * scale} of query row i * 128 + rl of unit u: left by attn_row_info_kernel in the (not yet written) dq
    // slot of that row and head.  Rows past L get lse = +inf (P = dS = 0).
    auto row_info = [&](int u, int i) {
      const int img = u / p.heads, head = u - img * p.heads;
      const int r = i * QT + rl;
      float2 v = make_float2(INFINITY, 0.0f);
      if (r < p.L) {
        const float* q = reinterpret_cast<const float*>(p.dqkv + (static_cast<int64_t>(img) * p.L + r) * p.ld_qkv + head * DH);
        asm volatile("ld.global.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(q) : "memory");   // (plain load: the slot is rewritten later)
      }
      return v;
    };
    // one 32-key pass: P = 2^(s c - lse2), dS = P (dP scale - D scale), both packed to bf16 pairs
    auto pass = [&](const float* sv, const float* dv, const float2 info, uint32_t* wp_, uint32_t* wd_, int first_col, int my_cols, bool full) {
      if (full) {
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          const float p0 = ex2_approx(fmaf(sv[2 * e], p.c, -info.x)), p1 = ex2_approx(fmaf(sv[2 * e + 1], p.c, -info.x));
          wp_[e] = pack_bf16x2(p0, p1);
          wd_[e] = pack_bf16x2(p0 * fmaf(dv[2 * e], nscale, -info.y), p1 * fmaf(dv[2 * e + 1], nscale, -info.y));
        }
      } else {
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          // keys past ncols hold stale tensor-memory columns: their P / dS land in tile columns no contraction reads (the
          // k-steps stop at ncols), but they are kept finite (zero)
          const bool live = first_col + 2 * e < my_cols;
          const float p0 = live ? ex2_approx(fmaf(sv[2 * e], p.c, -info.x)) : 0.0f;
          const float p1 = live ? ex2_approx(fmaf(sv[2 * e + 1], p.c, -info.x)) : 0.0f;
          wp_[e] = pack_bf16x2(p0, p1);
          wd_[e] = pack_bf16x2(live ? p0 * fmaf(dv[2 * e], nscale, -info.y) : 0.0f, live ? p1 * fmaf(dv[2 * e + 1], nscale, -info.y) : 0.0f);
        }
      }
    };
    // dV_j (warps 0-3) / dK_j (warps 4-7) of a finished key group: 32 key rows x 64 columns per warp
    auto read_dvk = [&](int img, int head, int j, int jgroup) {
      mbar_wait(dvk_full, static_cast<uint32_t>(jgroup) & 1u);
      __syncwarp();
      tc_fence_after();
      float o[64];
      tmem_ld64(tmem_base + lane_addr + (half == 0 ? COL_DV : COL_DK), o);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(dvk_free);
      __syncwarp();
      const int key = j * QT + rl;
      if (key < p.L) {
        __nv_bfloat16* dst = p.dqkv + (static_cast<int64_t>(img) * p.L + key) * p.ld_qkv + (half == 0 ? 2 * H : H) + head * DH;
#pragma unroll
        for (int c8 = 0; c8 < 4; ++c8) {
          uint32_t w[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) w[e] = pack_bf16x2(o[c8 * 16 + 2 * e], o[c8 * 16 + 2 * e + 1]);
          st_na_v8(dst + c8 * 16, w);
        }
      }
    };

    // dQ_0 (warps 0-3) / dQ_1 (warps 4-7) of a finished unit
    auto read_dq = [&](int img, int head, int itq) {
      mbar_wait(dq_full, static_cast<uint32_t>(itq) & 1u);
      __syncwarp();
      tc_fence_after();
      float o[64];
      const bool tile_on = half < p.nq;
      if (tile_on) tmem_ld64(tmem_base + lane_addr + COL_DQ + half * DH, o);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(dq_free);
      __syncwarp();
      const int qrow = half * QT + rl;
      if (tile_on && qrow < p.L) {
        __nv_bfloat16* dst = p.dqkv + (static_cast<int64_t>(img) * p.L + qrow) * p.ld_qkv + head * DH;
#pragma unroll
        for (int c8 = 0; c8 < 4; ++c8) {
          uint32_t w[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) w[e] = pack_bf16x2(o[c8 * 16 + 2 * e], o[c8 * 16 + 2 * e + 1]);
          st_na_v8(dst + c8 * 16, w);
        }
      }
    };

    float2 info0 = make_float2(INFINITY, 0.0f), info1 = info0, next0 = info0, next1 = info0;
    if (my_units > 0) {
      info0 = row_info(blockIdx.x, 0);
      if (p.nq > 1) info1 = row_info(blockIdx.x, 1);
    }
    int g = 0, jg = 0, it = 0;
    int pend_img = -1, pend_head = 0, pend_j = 0, pend_jg = 0;      // a finished key group whose dV / dK are still in tensor memory
    int pend_dq_img = -1, pend_dq_head = 0, pend_dq_it = 0;         // a finished unit whose dQ is still in tensor memory
#ifdef A4R_ATTN_TIMING
    long long acc_t[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#endif
    for (int u = blockIdx.x; u < units; u += gridDim.x, ++it) {
      const int img = u / p.heads, head = u - img * p.heads;
      // the next unit's row information: two 8-byte loads, in flight for the whole unit
      if (u + static_cast<int>(gridDim.x) < units) {
        next0 = row_info(u + gridDim.x, 0);
        if (p.nq > 1) next1 = row_info(u + gridDim.x, 1);
      }
      for (int j = 0; j < p.nkj; ++j) {
        for (int i = 0; i < p.nq; ++i, ++g) {
          const int ncols = min(128, p.Lk - j * QT);
          const int my_cols = max(0, min(64, ncols - half * 64));        // 64, 16..48, or 0
          const bool rows_on = i * QT + quad * 32 < p.L;
          const bool on = rows_on && my_cols > 0;
          uint32_t wp[32], wd[32];
          BSTAMP(b0);
          mbar_wait(sdp_full, static_cast<uint32_t>(g) & 1u);
          BSTAMP(b1);
          BACC(0, b0, b1);
          __syncwarp();
          tc_fence_after();
          if (on) {
            const uint32_t ts = tmem_base + lane_addr + COL_S + half * 64, td = tmem_base + lane_addr + COL_DP + half * 64;
            const float2 info = i == 0 ? info0 : info1;
            // two passes of 32 keys: the second pass's scores are in flight while the first is exponentiated
            float sa[32], da[32], sb[32], db[32];
            tmem_ld32(ts, sa);
            tmem_ld32(td, da);
            tmem_ld_wait();
            if (my_cols > 32) {
              tmem_ld32(ts + 32, sb);
              tmem_ld32(td + 32, db);
            }
            pass(sa, da, info, wp, wd, 0, my_cols, my_cols >= 32);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(sdp_free);
            __syncwarp();
            pass(sb, db, info, wp + 16, wd + 16, 32, my_cols, my_cols == 64);
          } else {
            if (lane == 0) mbar_arrive(sdp_free);
            __syncwarp();
#pragma unroll
            for (int e = 0; e < 32; ++e) wp[e] = wd[e] = 0u;
          }
          // the P / dS tiles of the previous step must have been consumed by M3-M5
          BSTAMP(b2);
          BACC(1, b1, b2);
          mbar_wait(pds_free, (static_cast<uint32_t>(g) & 1u) ^ 1u);
          BSTAMP(b3);
          BACC(2, b2, b3);
          if (rows_on) {      // my 64 keys of row rl: one 128-byte row of key chunk `half` (8 x 16 B, 128-byte swizzle)
#pragma unroll
            for (int c16 = 0; c16 < 8; ++c16) {
              const uint32_t off = half * BOX + rl * 128 + ((c16 ^ (rl & 7)) << 4);
              sts_v4(sp + off, wp[4 * c16], wp[4 * c16 + 1], wp[4 * c16 + 2], wp[4 * c16 + 3]);
              sts_v4(sds + off, wd[4 * c16], wd[4 * c16 + 1], wd[4 * c16 + 2], wd[4 * c16 + 3]);
            }
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(pds_ready);
          __syncwarp();
          BSTAMP(b4);
          BACC(3, b3, b4);
          // The dV / dK of the key group that ended with the PREVIOUS step are read out now, after this step's tiles have been
          // handed to the issuer: their M3 / M4 have long retired, and the issuer has M1 / M2 of the next step to run meanwhile.
          if (pend_img >= 0) {
            read_dvk(pend_img, pend_head, pend_j, pend_jg);
            pend_img = -1;
          }
          if (pend_dq_img >= 0) {
            read_dq(pend_dq_img, pend_dq_head, pend_dq_it);
            pend_dq_img = -1;
          }
          BSTAMP(b5);
          BACC(4, b4, b5);
          if (i == p.nq - 1) {
            pend_img = img;
            pend_head = head;
            pend_j = j;
            pend_jg = jg;
            ++jg;
          }
        }
      }
      // dQ of this unit is read out during the NEXT unit's first step (like dV / dK): its last M5 retires while the compute warps
      // already exponentiate that step, instead of the whole CTA draining its pipeline at every unit boundary.
      pend_dq_img = img;
      pend_dq_head = head;
      pend_dq_it = it;
      info0 = next0;
      info1 = next1;
    }
    if (pend_img >= 0) read_dvk(pend_img, pend_head, pend_j, pend_jg);
    if (pend_dq_img >= 0) read_dq(pend_dq_img, pend_dq_head, pend_dq_it);
#ifdef A4R_ATTN_TIMING
    if (blockIdx.x == 0 && cw == 0 && lane == 0) {
      for (int i = 0; i < 8; ++i) g_attn_bwd_timing[i] = acc_t[i];
      g_attn_bwd_timing[8] = it;
    }
#endif
  }
#ifdef A4R_ATTN_TIMING
  if (blockIdx.x == 0 && warp == 4 && lane == 0) {
    // (declared in the compute branch: re-read through the global array written there)
  }
#endif
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

}  // namespace

// arguments are validated by the caller (a4r_attn_mid_bwd); unmasked, head_dim 64, L <= 256, dqkv row stride = ld_qkv
int a4r_attn_vit_tc_bwd(const a4r_attn_args* a, cudaStream_t stream) {
  BwdParams p;
  p.dout = static_cast<const __nv_bfloat16*>(a->dout);
  p.ctx = static_cast<const __nv_bfloat16*>(a->ctx);
  p.lse = a->lse;
  p.dqkv = static_cast<__nv_bfloat16*>(a->out);
  p.ld_qkv = a->ld_qkv;
  p.ld_out = a->ld_out;
  p.N = static_cast<int>(a->N);
  p.L = static_cast<int>(a->L);
  p.Lk = (p.L + 15) & ~15;
  p.heads = static_cast<int>(a->heads);
  p.nq = (p.L + QT - 1) / QT;
  p.nkj = (p.Lk + QT - 1) / QT;
  p.scale = a->scale;
  p.c = a->scale * LOG2E_F;
  CUtensorMap tmQKV, tmDO;
  int rc = make_tmap_tokens(&tmQKV, a->qkv, a->N, a->L, 3 * a->heads * DH, a->ld_qkv);
  if (rc != A4R_OK) return rc;
  if ((rc = make_tmap_tokens(&tmDO, a->dout, a->N, a->L, a->heads * DH, a->ld_out)) != A4R_OK) return rc;
  const size_t smem = static_cast<size_t>(BWD_SMEM) + 1024;
  const int64_t units = a->N * a->heads;
  const int grid = static_cast<int>(units < a4r_num_sms() ? units : a4r_num_sms());
  A4R_CUDA_OK(cudaFuncSetAttribute(attn_vit_tc_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  const int64_t rows = a->N * a->L;
  const int info_blocks = static_cast<int>(rows < 8 * 8 * a4r_num_sms() ? (rows + 7) / 8 : 8 * a4r_num_sms());
  const int info_iters = (p.heads * 8 + 31) / 32;                 // 32 eight-column chunks per warp pass
  if (info_iters <= 1)
    attn_row_info_kernel<1><<<info_blocks, 256, 0, stream>>>(p.ctx, p.dout, p.lse, p.dqkv, rows, p.heads, p.ld_out, p.ld_qkv, p.scale);
  else if (info_iters <= 3)
    attn_row_info_kernel<3><<<info_blocks, 256, 0, stream>>>(p.ctx, p.dout, p.lse, p.dqkv, rows, p.heads, p.ld_out, p.ld_qkv, p.scale);
  else if (info_iters <= 8)
    attn_row_info_kernel<8><<<info_blocks, 256, 0, stream>>>(p.ctx, p.dout, p.lse, p.dqkv, rows, p.heads, p.ld_out, p.ld_qkv, p.scale);
  else
    return a4r_set_error(A4R_EINVAL, "attention backward: more than 32 heads are not supported by the row-information kernel");
  A4R_LAUNCH_OK();
  attn_vit_tc_bwd_kernel<<<grid, BWD_THREADS, smem, stream>>>(tmQKV, tmDO, p);
  A4R_LAUNCH_OK();
  a4r_count_launch(2);
  return A4R_OK;
}
