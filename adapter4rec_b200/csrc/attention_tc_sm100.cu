// attention_tc_sm100.cu — unmasked mid-length attention (32 < L <= 256, head_dim 64) on tcgen05: ViT-B/16's 197 (+ prompt) tokens.
//
//   ctx = softmax(q kᵀ / sqrt(d)) v        transformers' ViTSelfAttention, reached from Vit_Encoder.forward
//                                          (Downstream/CV/model/encoders.py:31-32)
//
// One work unit = one (image, head).  A persistent CTA walks the units; per unit the whole K and V of the head (Lk = L rounded up
// to 16 keys) and the Q rows (one or two 128-row tiles) sit in shared memory, loaded by TMA from the fused-QKV activation
// [N*L, 3H] through a 3-D tensor map (column, token, image) whose out-of-range tokens read as zeros.
//
//   warp 0      TMA producer: Q | K, then V of unit u+1 while unit u is being computed (two 96 KB stages)
//   warp 1      tcgen05 issuer:  S_t = Q_t Kᵀ   (A, B K-major in shared memory;  M = 128 queries, N = Lk keys, K = 64)
//                                O_t = P_t V    (A = P_t in TENSOR MEMORY, B = V MN-major in shared memory;  N = 64, K = Lk)
//   warp 2      TMEM allocator
//   warps 4-7 / 8-11   softmax + epilogue of query tile 0 / 1: one query row per thread (= one TMEM lane).  The thread pulls its
//               whole score row out of TMEM ONCE (TMEM read-out, 64 B/cycle/SM, is the scarcest resource of this kernel: a second
//               pass for the row maximum would double it), keeps it in registers (setmaxnreg: 232 registers for these warps),
//               takes the maximum, exponentiates with one FFMA + one MUFU.EX2 per score, and writes P back as packed bf16 over
//               the first columns of the dead score row; the issuer then runs P V straight out of tensor memory, so P never
//               touches shared memory.  O lands in columns [128, 192) of the same (dead) score region.
//               (The row is consumed in 64-column register blocks, the next block in flight while the current one is exponentiated.)
// The per-row log-sum-exp (natural log, scaled-score domain) is stored for the backward pass.
// Tensor-memory map (512 columns): tile t owns [256 t, 256 t + 256): S = [0, Lk) fp32, P = [0, Lk/2) packed bf16, O = [128, 192).
#include "attn_tc_common.cuh"

namespace {

constexpr int STAGE = 6 * BOX;           // Q (2 boxes) | K (2) | V (2)
constexpr int NSTAGE = 2;
constexpr int THREADS = 384;
constexpr int BLK_CH = 4;                 // score chunks of 16 columns per register block (two blocks in flight)
constexpr int TILE_COLS = 256;
constexpr int O_COL = 128;
constexpr float LOG2E = 1.4426950408889634f;

struct TcParams {
  __nv_bfloat16* out;
  float* lse;
  int64_t ld_out;
  int N, L, Lk, heads, nq, nk;
  int act[2];                            // active softmax warps per query tile
  float scale, c;                        // c = scale * log2(e)
};

#ifdef A4R_ATTN_TIMING
__device__ long long g_attn_timing[2][8];
#define TSTAMP(var) const long long var = clock64()
#else
#define TSTAMP(var)
#endif

__global__ void __launch_bounds__(THREADS, 1)
attn_vit_tc_fwd_kernel(const __grid_constant__ CUtensorMap tmQKV, const TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + NSTAGE * STAGE);
  uint64_t* qk_full = bars;            // [2] stage
  uint64_t* v_full = bars + 2;         // [2]
  uint64_t* qk_free = bars + 4;        // [2]
  uint64_t* v_free = bars + 6;         // [2]
  uint64_t* s_full = bars + 8;         // [2] query tile
  uint64_t* p_ready = bars + 10;       // [2]
  uint64_t* o_full = bars + 12;        // [2]
  uint64_t* s_free = bars + 14;        // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int units = p.N * p.heads;
  const int H = p.heads * DH;

  if (warp == 0 && lane == 0) tma_prefetch_desc(&tmQKV);
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(&qk_full[s], 1);
      mbar_init(&v_full[s], 1);
      mbar_init(&qk_free[s], 1);
      mbar_init(&v_free[s], 1);
      mbar_init(&s_full[s], 1);
      mbar_init(&o_full[s], 1);
      mbar_init(&p_ready[s], p.act[s] > 0 ? p.act[s] : 1);
      mbar_init(&s_free[s], p.act[s] > 0 ? p.act[s] : 1);
    }
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
    if (warp == 0 && lane == 0) {
      // ============================== TMA producer ==============================
      int it = 0;
      for (int u = blockIdx.x; u < units; u += gridDim.x, ++it) {
        const int st = it & 1;
        const uint32_t ph = static_cast<uint32_t>(it >> 1) & 1u;
        const int img = u / p.heads, head = u - img * p.heads;
        uint8_t* sq = smem + st * STAGE;
        mbar_wait(&qk_free[st], ph ^ 1u);
        mbar_expect_tx(&qk_full[st], static_cast<uint32_t>((p.nq + p.nk) * BOX));
        for (int t = 0; t < p.nq; ++t) tma_load_3d(&tmQKV, sq + t * BOX, &qk_full[st], head * DH, t * QT, img);
        for (int j = 0; j < p.nk; ++j) tma_load_3d(&tmQKV, sq + (2 + j) * BOX, &qk_full[st], H + head * DH, j * QT, img);
        mbar_wait(&v_free[st], ph ^ 1u);
        mbar_expect_tx(&v_full[st], static_cast<uint32_t>(p.nk * BOX));
        for (int j = 0; j < p.nk; ++j) tma_load_3d(&tmQKV, sq + (4 + j) * BOX, &v_full[st], 2 * H + head * DH, j * QT, img);
      }
    } else if (warp == 1 && lane == 0) {
      // ============================== MMA issuer ==============================
      const uint32_t idesc_s = umma_idesc_bf16(QT, static_cast<uint32_t>(p.Lk));
      const uint32_t idesc_o = umma_idesc_bf16(QT, DH) | (1u << 16);      // B (= V) MN-major
      const int ksteps = p.Lk / 16;
      const int my_units = (units - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
      // Two in-order queues — score MMAs (unit i, tile t) and P V MMAs (unit j, tile t') — merged by readiness: the scores of the
      // next unit are issued as soon as their tile's tensor-memory region has been read out, without waiting for the other
      // tile's P V of the current unit.
      int si = 0, st_ = 0, pj = 0, pt = 0;
      while (pj < my_units) {
        bool progressed = false;
        if (si < my_units) {
          const int stg = si & 1;
          if (mbar_try_wait(&qk_full[stg], static_cast<uint32_t>(si >> 1) & 1u) &&
              mbar_try_wait(&s_free[st_], (static_cast<uint32_t>(si) & 1u) ^ 1u)) {
            tc_fence_after();
            const uint32_t sq = smem_u32(smem + stg * STAGE);
            const uint64_t adesc = umma_desc_k_sw128(sq + st_ * BOX), bdesc = umma_desc_k_sw128(sq + 2 * BOX);
#pragma unroll
            for (int k = 0; k < DH / 16; ++k)
              umma_bf16_ss(tmem_base + st_ * TILE_COLS, adesc + static_cast<uint64_t>(2 * k), bdesc + static_cast<uint64_t>(2 * k),
                           idesc_s, k != 0 ? 1u : 0u);
            umma_commit(&s_full[st_]);
#ifdef A4R_ATTN_TIMING
            if (blockIdx.x == 0) {
              const long long q0 = clock64();
              mbar_wait(&s_full[st_], static_cast<uint32_t>(si) & 1u);
              g_attn_timing[0][7] += clock64() - q0;
            }
#endif
            if (++st_ == p.nq) {
              umma_commit(&qk_free[stg]);                  // Q and K of this stage are dead once the score MMAs retire
              st_ = 0;
              ++si;
            }
            progressed = true;
          }
        }
        if (!progressed && pj < si + (st_ > pt ? 1 : 0)) {   // the P V of (pj, pt) follows the scores of (pj, pt)
          const int stg = pj & 1;
          if (mbar_try_wait(&v_full[stg], static_cast<uint32_t>(pj >> 1) & 1u) && mbar_try_wait(&p_ready[pt], static_cast<uint32_t>(pj) & 1u)) {
            tc_fence_after();
            const uint32_t sq = smem_u32(smem + stg * STAGE);
            const uint64_t vdesc = umma_desc_mn_sw128_1chunk(sq + 4 * BOX);
            for (int k = 0; k < ksteps; ++k)
              umma_bf16_ts(tmem_base + pt * TILE_COLS + O_COL, tmem_base + pt * TILE_COLS + 8 * k, vdesc + static_cast<uint64_t>(128 * k),
                           idesc_o, k != 0 ? 1u : 0u);
            umma_commit(&o_full[pt]);
#ifdef A4R_ATTN_TIMING
            if (blockIdx.x == 0) {
              const long long q0 = clock64();
              mbar_wait(&o_full[pt], static_cast<uint32_t>(pj) & 1u);
              g_attn_timing[1][7] += clock64() - q0;
            }
#endif
            if (++pt == p.nq) {
              umma_commit(&v_free[stg]);
              pt = 0;
              ++pj;
            }
          }
        }
      }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
    // ============================== softmax + epilogue: one query row per thread ==============================
    const int t = (warp - 4) >> 2, quad = warp & 3;
    const int row = t * QT + quad * 32 + lane;
    const bool warp_on = t * QT + quad * 32 < p.L;
    const uint32_t tbase = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + t * TILE_COLS;
    const int nch = p.Lk / 16;
    const int nblk = (nch + BLK_CH - 1) / BLK_CH;
    if (warp_on) {
      int it = 0;
#ifdef A4R_ATTN_TIMING
      long long acc_t[6] = {0, 0, 0, 0, 0, 0};
#endif
      for (int u = blockIdx.x; u < units; u += gridDim.x, ++it) {
        const uint32_t tp = static_cast<uint32_t>(it) & 1u;
        const int img = u / p.heads, head = u - img * p.heads;
        TSTAMP(ts0);
        mbar_wait(&s_full[t], tp);
        TSTAMP(ts1);
        __syncwarp();
        tc_fence_after();
        // The score row is consumed in blocks of BLK_CH * 16 = 64 columns, each read from tensor memory ONCE; block i + 1 is in
        // flight while block i is exponentiated (tcgen05.wait::ld waits for every outstanding load, so the next block is issued
        // right AFTER the wait for the current one).  Softmax is shift-invariant: the shift is the running maximum, but it is
        // only RAISED when a block beats it by more than 2^20 (then the P columns written so far and the running sum are
        // rescaled in tensor memory: exact, and rare), so the common path reads every score exactly once.
        float sum = 0.0f, mc = 0.0f;
        float sA[BLK_CH * 16], sB[BLK_CH * 16];
        auto load_block = [&](int blk, float* dst) {
#ifdef A4R_ATTN_LDTM_OFF
          if (blk != 0) return;
#endif
          if (blk * BLK_CH + BLK_CH <= nch) {          // a whole block: one 64-column load
            tmem_ld64(tbase + blk * BLK_CH * 16, dst);
          } else {
#pragma unroll
            for (int i = 0; i < BLK_CH; ++i)
              if (blk * BLK_CH + i < nch) tmem_ld16(tbase + (blk * BLK_CH + i) * 16, dst + i * 16);
          }
        };
        auto process = [&](int blk, float* sv) {
          if (blk == nblk - 1 && (p.L & 15) != 0) {          // the partial last chunk: keys >= L must not count (exp2(-inf) = 0)
#pragma unroll
            for (int i = 0; i < BLK_CH; ++i)
#pragma unroll
              for (int e = 0; e < 16; ++e)
                if ((blk * BLK_CH + i) * 16 + e >= p.L) sv[i * 16 + e] = -INFINITY;
          }
          float bm = -INFINITY;
#pragma unroll
          for (int i = 0; i < BLK_CH; ++i)
            if (blk * BLK_CH + i < nch) {
#pragma unroll
              for (int e = 0; e < 16; ++e) bm = fmaxf(bm, sv[i * 16 + e]);
            }
          const float bmc = bm * p.c;
          if (blk == 0) {
            mc = bmc;
          } else if (__any_sync(0xffffffffu, bmc > mc + 20.0f)) {
            // rare: raise the shift of the rows that need it; rescale their P columns [0, 8 * BLK_CH * blk) and running sums
            const float nmc = bmc > mc + 20.0f ? bmc : mc;
            const float f = ex2_approx(mc - nmc);
            mc = nmc;
            sum *= f;
            tmem_st_wait();
            for (int cc = 0; cc < blk * BLK_CH; ++cc) {
              uint32_t w[8];
              asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                           : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7])
                           : "r"(tbase + cc * 8)
                           : "memory");
              tmem_ld_wait();
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const float2 v = bf16x2_to_f2(w[e]);
                w[e] = pack_bf16x2(v.x * f, v.y * f);
              }
              tmem_st_32x32b_x8(tbase + cc * 8, w);
            }
          }
#pragma unroll
          for (int i = 0; i < BLK_CH; ++i) {
            const int ch = blk * BLK_CH + i;
            if (ch < nch) {
              uint32_t w[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) {
#ifdef A4R_ATTN_EXP_OFF      // timing experiment only (tools/micro/attn_tc_timing.cu)
                const float p0 = fmaf(sv[i * 16 + 2 * e], p.c, -mc), p1 = fmaf(sv[i * 16 + 2 * e + 1], p.c, -mc);
#else
                const float p0 = ex2_approx(fmaf(sv[i * 16 + 2 * e], p.c, -mc));
                const float p1 = ex2_approx(fmaf(sv[i * 16 + 2 * e + 1], p.c, -mc));
#endif
                sum += p0 + p1;
                w[e] = pack_bf16x2(p0, p1);
              }
#ifdef A4R_ATTN_STTM_OFF
              if (w[0] == 0x12345678u)
#endif
              tmem_st_32x32b_x8(tbase + ch * 8, w);     // P chunk ch lands on score columns of chunk ch / 2: already in registers
            }
          }
        };
        load_block(0, sA);
        tmem_ld_wait();
        TSTAMP(ts2);
        for (int blk = 0; blk < nblk; blk += 2) {
          if (blk + 1 < nblk) load_block(blk + 1, sB);
          process(blk, sA);
          tmem_ld_wait();
          if (blk + 1 < nblk) {
            if (blk + 2 < nblk) load_block(blk + 2, sA);
            process(blk + 1, sB);
            tmem_ld_wait();
          }
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_ready[t]);
        __syncwarp();
        TSTAMP(ts3);
        // ---- O = P V is on its way: normalise and store the row ----
        mbar_wait(&o_full[t], tp);
        TSTAMP(ts4);
        __syncwarp();
        tc_fence_after();
        float o[DH];
#pragma unroll
        for (int c4 = 0; c4 < DH / 16; ++c4) tmem_ld16(tbase + O_COL + c4 * 16, o + c4 * 16);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_free[t]);
        __syncwarp();
        TSTAMP(ts5);
        if (row < p.L) {
          const float inv = 1.0f / sum;
          __nv_bfloat16* dst = p.out + (static_cast<int64_t>(img) * p.L + row) * p.ld_out + head * DH;
#pragma unroll
          for (int c8 = 0; c8 < DH / 16; ++c8) {
            uint32_t w[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) w[e] = pack_bf16x2(o[c8 * 16 + 2 * e] * inv, o[c8 * 16 + 2 * e + 1] * inv);
            st_na_v8(dst + c8 * 16, w);
          }
          if (p.lse != nullptr)
            p.lse[(static_cast<int64_t>(img) * p.L + row) * p.heads + head] = mc * 0.6931471805599453f + __logf(sum);   // mc = shift * scale * log2(e)
        }
#ifdef A4R_ATTN_TIMING
        const long long ts6 = clock64();
        acc_t[0] += ts1 - ts0; acc_t[1] += ts2 - ts1; acc_t[2] += ts3 - ts2; acc_t[3] += ts4 - ts3; acc_t[4] += ts5 - ts4; acc_t[5] += ts6 - ts5;
#endif
      }
#ifdef A4R_ATTN_TIMING
      if (blockIdx.x == 0 && lane == 0 && quad == 0) {
        for (int i = 0; i < 6; ++i) g_attn_timing[t][i] = acc_t[i];
        g_attn_timing[t][6] = it;
      }
#endif
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

}  // namespace

// arguments are validated by the caller (a4r_attn_mid_fwd); unmasked, head_dim 64, L <= 256
int a4r_attn_vit_tc_fwd(const a4r_attn_args* a, cudaStream_t stream) {
  TcParams p;
  p.out = static_cast<__nv_bfloat16*>(a->out);
  p.lse = a->lse;
  p.ld_out = a->ld_out;
  p.N = static_cast<int>(a->N);
  p.L = static_cast<int>(a->L);
  p.Lk = (p.L + 15) & ~15;
  p.heads = static_cast<int>(a->heads);
  p.nq = (p.L + QT - 1) / QT;
  p.nk = (p.Lk + QT - 1) / QT;
  for (int t = 0; t < 2; ++t) {
    const int rows = p.L - t * QT;
    p.act[t] = rows <= 0 ? 0 : (rows >= QT ? 4 : (rows + 31) / 32);
  }
  p.scale = a->scale;
  p.c = a->scale * LOG2E;
  CUtensorMap tm;
  int rc = make_tmap_tokens(&tm, a->qkv, a->N, a->L, 3 * a->heads * DH, a->ld_qkv);
  if (rc != A4R_OK) return rc;
  const size_t smem = static_cast<size_t>(NSTAGE) * STAGE + 16 * sizeof(uint64_t) + 16 + 1024;
  const int64_t units = a->N * a->heads;
  const int grid = static_cast<int>(units < a4r_num_sms() ? units : a4r_num_sms());
  A4R_CUDA_OK(cudaFuncSetAttribute(attn_vit_tc_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  attn_vit_tc_fwd_kernel<<<grid, THREADS, smem, stream>>>(tm, p);
  A4R_LAUNCH_OK();
  a4r_count_launch(1);
  return A4R_OK;
}
