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
constexpr int OUT_BOX = 32 * 128;           // one warp's result box
constexpr float LOG2E = 1.4426950408889634f;
#ifdef A4R_ATTN_POLL_TRY            // (A/B build of tools/micro only: the blocking probe this kernel used to poll with)
#define A4R_POLL mbar_try_wait
#else
#define A4R_POLL mbar_test_wait
#endif

template <bool B>
struct FullBlock {
  static constexpr bool value = B;
};

struct TcParams {
  __nv_bfloat16* out;
  float* lse;
  int64_t ld_out;
  int N, L, Lk, heads, nq, nk;
  int act[2];                            // active softmax warps per query tile
  float scale, c;                        // c = scale * log2(e)
};

#ifdef A4R_ATTN_TIMING
__device__ long long g_attn_timing[8][8];      // [tile * 4 + quad][phase]; [0][7], [1][7]: issuer
#define TSTAMP(var) const long long var = clock64()
__device__ long long g_attn_trace[16][10][6];   // [unit][0-7: tile*4+quad, 8: issuer scores, 9: issuer P V][event]
#else
#define TSTAMP(var)
#endif

__global__ void __launch_bounds__(THREADS, 1)
attn_vit_tc_fwd_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmOut, const TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* s_out = smem + NSTAGE * STAGE;            // 8 softmax warps x [32 rows x 128 B] result boxes
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_out + 8 * OUT_BOX);
  uint64_t* qk_full = bars;            // [2] stage
  uint64_t* v_full = bars + 2;         // [2]
  uint64_t* qk_free = bars + 4;        // [2]
  uint64_t* v_free = bars + 6;         // [2]
  uint64_t* s_full = bars + 8;         // [2] query tile
  uint64_t* p_ready = bars + 10;       // [2]
  uint64_t* o_full = bars + 12;        // [2]
  uint64_t* s_free = bars + 14;        // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;   // (warp-uniform for the compiler)
  const int units = p.N * p.heads;
  const int H = p.heads * DH;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQKV);
    tma_prefetch_desc(&tmOut);
  }
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
    if (warp == 0) {
      // ============================== TMA producer (whole warp, converged) ==============================
      int it = 0;
      for (int u = blockIdx.x; u < units; u += gridDim.x, ++it) {
        const int st = it & 1;
        const uint32_t ph = static_cast<uint32_t>(it >> 1) & 1u;
        const int img = u / p.heads, head = u - img * p.heads;
        uint8_t* sq = smem + st * STAGE;
        mbar_wait(&qk_free[st], ph ^ 1u);
        __syncwarp();
        mbar_expect_tx_elect(&qk_full[st], static_cast<uint32_t>((p.nq + p.nk) * BOX));
        for (int t = 0; t < p.nq; ++t) tma_load_3d_elect(&tmQKV, sq + t * BOX, &qk_full[st], head * DH, t * QT, img);
        for (int j = 0; j < p.nk; ++j) tma_load_3d_elect(&tmQKV, sq + (2 + j) * BOX, &qk_full[st], H + head * DH, j * QT, img);
        mbar_wait(&v_free[st], ph ^ 1u);
        __syncwarp();
        mbar_expect_tx_elect(&v_full[st], static_cast<uint32_t>(p.nk * BOX));
        for (int j = 0; j < p.nk; ++j) tma_load_3d_elect(&tmQKV, sq + (4 + j) * BOX, &v_full[st], 2 * H + head * DH, j * QT, img);
      }
    } else if (warp == 1) {
      // ============================== MMA issuer: the WHOLE warp, converged (see umma_*_elect) ==============================
      const uint32_t idesc_s = umma_idesc_bf16(QT, static_cast<uint32_t>(p.Lk));
      const uint32_t idesc_o = umma_idesc_bf16(QT, DH) | (1u << 16);      // B (= V) MN-major
      const int ksteps = p.Lk / 16;
      const int my_units = (units - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
      // Two in-order queues — score MMAs (unit i, tile t) and P V MMAs (unit j, tile t') — merged by readiness: the scores of the
      // next unit are issued as soon as their tile's tensor-memory region has been read out, without waiting for the other
      // tile's P V of the current unit.  Probes are non-blocking and made warp-uniform by a vote.
      int si = 0, st_ = 0, pj = 0, pt = 0;
      while (pj < my_units) {
        bool progressed = false;
        if (si < my_units) {
          const int stg = si & 1;
          if (__all_sync(0xffffffffu, A4R_POLL(&qk_full[stg], static_cast<uint32_t>(si >> 1) & 1u) &&
                                          A4R_POLL(&s_free[st_], (static_cast<uint32_t>(si) & 1u) ^ 1u))) {
            tc_fence_after();
            const uint32_t sq = smem_u32(smem + stg * STAGE);
            const uint64_t adesc = umma_desc_k_sw128(sq + st_ * BOX), bdesc = umma_desc_k_sw128(sq + 2 * BOX);
#pragma unroll
            for (int k = 0; k < DH / 16; ++k)
              umma_bf16_ss_elect(tmem_base + st_ * TILE_COLS, adesc + static_cast<uint64_t>(2 * k), bdesc + static_cast<uint64_t>(2 * k),
                                 idesc_s, k != 0 ? 1u : 0u);
            umma_commit_elect(&s_full[st_]);
#ifdef A4R_ATTN_TIMING
            if (blockIdx.x == 0 && si < 16 && lane == 0) g_attn_trace[si][8][st_] = clock64();
#endif
            if (++st_ == p.nq) {
              umma_commit_elect(&qk_free[stg]);            // Q and K of this stage are dead once the score MMAs retire
              st_ = 0;
              ++si;
            }
            progressed = true;
          }
        }
        if (!progressed && pj < si + (st_ > pt ? 1 : 0)) {   // the P V of (pj, pt) follows the scores of (pj, pt)
          const int stg = pj & 1;
          if (__all_sync(0xffffffffu, A4R_POLL(&v_full[stg], static_cast<uint32_t>(pj >> 1) & 1u) &&
                                          A4R_POLL(&p_ready[pt], static_cast<uint32_t>(pj) & 1u))) {
#ifdef A4R_ATTN_TIMING
            if (blockIdx.x == 0 && pj < 16 && lane == 0) g_attn_trace[pj][9][2 + pt] = clock64();
#endif
            tc_fence_after();
            const uint32_t sq = smem_u32(smem + stg * STAGE);
            const uint64_t vdesc = umma_desc_mn_sw128_1chunk(sq + 4 * BOX);
            const uint32_t ts = tmem_base + pt * TILE_COLS;
            umma_bf16_ts_elect<0>(ts + O_COL, ts, vdesc, idesc_o);
#pragma unroll 4
            for (int k = 1; k < ksteps; ++k) umma_bf16_ts_elect<1>(ts + O_COL, ts + 8 * k, vdesc + static_cast<uint64_t>(128 * k), idesc_o);
            umma_commit_elect(&o_full[pt]);
#ifdef A4R_ATTN_TIMING
            if (blockIdx.x == 0 && pj < 16 && lane == 0) g_attn_trace[pj][9][pt] = clock64();
#endif
            if (++pt == p.nq) {
              umma_commit_elect(&v_free[stg]);
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
    const uint32_t obox = smem_u32(s_out) + (warp - 4) * OUT_BOX;
    const int nblk = (nch + BLK_CH - 1) / BLK_CH;
    const int ltail = p.L & 15;              // valid keys in a partial last chunk (0: the chunk is whole)
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
        // `full` blocks (all BLK_CH chunks valid, no key past L) run as ONE branch-free basic block: with a branch per chunk the
        // compiler schedules 16 exponentials at a time and a warp sits out the MUFU latency 13 times per row (measured: a warp
        // ALONE on its scheduler needed 23 cycles per score against the 8 the MUFU pipe takes).
        // Only block 0 takes a maximum (it sets the shift).  Later blocks exponentiate under the running shift straight away and
        // look at their own SUM: a block sum below 2^20 proves every term is (no overflow anywhere, P exact in bf16's range); a
        // larger one (or inf / NaN) sends the warp through the rare path — raise the shift of the rows that need it, rescale the
        // P columns written so far and the running sum, and exponentiate the block (still in registers) again.
        auto block_max = [&](int blk, const float* sv, bool full) {
          float bm4[BLK_CH];
#pragma unroll
          for (int i = 0; i < BLK_CH; ++i) {
            bm4[i] = -INFINITY;
            if (full || blk * BLK_CH + i < nch) {
#pragma unroll
              for (int e = 0; e < 16; ++e) bm4[i] = fmaxf(bm4[i], sv[i * 16 + e]);
            }
          }
          return fmaxf(fmaxf(bm4[0], bm4[1]), fmaxf(bm4[2], bm4[3]));
        };
        auto process = [&](int blk, float* sv, auto full_tag) {
          constexpr bool FULL = decltype(full_tag)::value;
          if (!FULL && ltail != 0) {                     // the partial last chunk: keys >= L must not count (exp2(-inf) = 0)
#pragma unroll
            for (int i = 0; i < BLK_CH; ++i)
              if (blk * BLK_CH + i == nch - 1) {
#pragma unroll
                for (int e = 0; e < 16; ++e)
                  if (e >= ltail) sv[i * 16 + e] = -INFINITY;
              }
          }
          if (blk == 0) mc = block_max(0, sv, FULL) * p.c;
          auto exponentiate = [&]() {
            // packed fp32 pairs (FFMA2 / FADD2): the shift-and-scale and the row sum cost half an instruction per score each
            float2 cs2[BLK_CH];                           // one partial sum pair per chunk: four independent chains
            const float2 c2 = make_float2(p.c, p.c), m2 = make_float2(-mc, -mc);
#pragma unroll
            for (int i = 0; i < BLK_CH; ++i) {
              const int ch = blk * BLK_CH + i;
              cs2[i] = make_float2(0.0f, 0.0f);
              if (FULL || ch < nch) {
                uint32_t w[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                  const float2 x = __ffma2_rn(make_float2(sv[i * 16 + 2 * e], sv[i * 16 + 2 * e + 1]), c2, m2);
#ifdef A4R_ATTN_EXP_OFF      // timing experiment only (tools/micro/attn_tc_timing.cu)
                  const float2 pp = x;
#else
                  const float2 pp = make_float2(ex2_approx(x.x), ex2_approx(x.y));
#endif
                  cs2[i] = __fadd2_rn(cs2[i], pp);
                  w[e] = pack_bf16x2(pp.x, pp.y);
                }
#ifdef A4R_ATTN_STTM_OFF
                if (w[0] == 0x12345678u)
#endif
                tmem_st_32x32b_x8(tbase + ch * 8, w);   // P chunk ch lands on score columns of chunk ch / 2: already in registers
              }
            }
            const float2 s01 = __fadd2_rn(cs2[0], cs2[1]), s23 = __fadd2_rn(cs2[2], cs2[3]);
            const float2 s4 = __fadd2_rn(s01, s23);
            return s4.x + s4.y;
          };
          float bs = exponentiate();
          if (blk != 0 && __any_sync(0xffffffffu, !(bs < 1048576.0f))) {
            // rare: raise the shift of the rows that need it; rescale their P columns [0, 8 * BLK_CH * blk) and running sums
            const float bmc = block_max(blk, sv, FULL) * p.c;
            const float nmc = !(bs < 1048576.0f) && bmc > mc ? bmc : mc;
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
            bs = exponentiate();
          }
          sum += bs;
        };
        auto process_block = [&](int blk, float* sv) {
          if (blk * BLK_CH + BLK_CH <= nch && (blk * BLK_CH + BLK_CH < nch || ltail == 0)) process(blk, sv, FullBlock<true>{});
          else process(blk, sv, FullBlock<false>{});
        };
        load_block(0, sA);
        tmem_ld_wait();
        TSTAMP(ts2);
        for (int blk = 0; blk < nblk; blk += 2) {
          if (blk + 1 < nblk) load_block(blk + 1, sB);
          process_block(blk, sA);
          tmem_ld_wait();
          if (blk + 1 < nblk) {
            if (blk + 2 < nblk) load_block(blk + 2, sA);
            process_block(blk + 1, sB);
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
        // The warp's 32 rows x 64 columns leave as ONE TMA store from its private 4 KB box (SWIZZLE_128B: the 16-byte row-per-thread
        // writes are conflict-free); rows >= L are clipped by the tensor map.  (32-byte-per-lane global stores touched 32 lines
        // per instruction and kept the warp ~950 cycles.)
        {
          const float inv = 1.0f / sum;
          __syncwarp();
          bulk_wait_read0_elect();                      // this box's previous store (one unit ago) has read its bytes
          __syncwarp();
#pragma unroll
          for (int c8 = 0; c8 < DH / 8; ++c8) {
            uint32_t w[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) w[e] = pack_bf16x2(o[c8 * 8 + 2 * e] * inv, o[c8 * 8 + 2 * e + 1] * inv);
            sts_v4(obox + lane * 128 + ((c8 ^ (lane & 7)) << 4), w[0], w[1], w[2], w[3]);
          }
          fence_proxy_async_smem();
          __syncwarp();
          tma_store_3d_commit_elect(&tmOut, obox, head * DH, t * QT + quad * 32, img);
          if (row < p.L && p.lse != nullptr)
            p.lse[(static_cast<int64_t>(img) * p.L + row) * p.heads + head] = mc * 0.6931471805599453f + __logf(sum);   // mc = shift * scale * log2(e)
        }
#ifdef A4R_ATTN_TIMING
        const long long ts6 = clock64();
        if (blockIdx.x == 0 && lane == 0 && it < 16) {
          long long* tr = g_attn_trace[it][t * 4 + quad];
          tr[0] = ts0; tr[1] = ts1; tr[2] = ts3; tr[3] = ts4; tr[4] = ts5; tr[5] = ts6;
        }
        acc_t[0] += ts1 - ts0; acc_t[1] += ts2 - ts1; acc_t[2] += ts3 - ts2; acc_t[3] += ts4 - ts3; acc_t[4] += ts5 - ts4; acc_t[5] += ts6 - ts5;
#endif
      }
      __syncwarp();
      bulk_wait_read0_elect();                          // the last store still reads this CTA's shared memory
#ifdef A4R_ATTN_TIMING
      if (blockIdx.x == 0 && lane == 0) {
        for (int i = 0; i < 6; ++i) g_attn_timing[t * 4 + quad][i] = acc_t[i];
        g_attn_timing[t * 4 + quad][6] = it;
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
  CUtensorMap tmo;
  if ((rc = make_tmap_tokens(&tmo, a->out, a->N, a->L, a->heads * DH, a->ld_out, 32)) != A4R_OK) return rc;
  const size_t smem = static_cast<size_t>(NSTAGE) * STAGE + 8 * OUT_BOX + 16 * sizeof(uint64_t) + 16 + 1024;
  static_assert(NSTAGE * STAGE + 8 * OUT_BOX + 16 * 8 + 16 + 1024 <= 232448, "shared-memory plan exceeds 227 KB");
  const int64_t units = a->N * a->heads;
  const int grid = static_cast<int>(units < a4r_num_sms() ? units : a4r_num_sms());
  A4R_CUDA_OK(cudaFuncSetAttribute(attn_vit_tc_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  attn_vit_tc_fwd_kernel<<<grid, THREADS, smem, stream>>>(tm, tmo, p);
  A4R_LAUNCH_OK();
  a4r_count_launch(1);
  return A4R_OK;
}
