// gemm_sm100.cu — persistent, warp-specialised tcgen05 + TMA GEMM for sm_100a.
//
//   C[M,N] = epi(alpha * (A[M,K]·B[N,K]ᵀ + A2[M,K2]·B2[N,K2]ᵀ) + bias[N])
//
// One kernel family carries every dense contraction of the TransRec hot path (SURVEY.md §2.5 K2, K4,
// K7, K2-L and the dgrad GEMMs): nn.Linear weights are already [out,in] = K-major, so forward uses
// B = W and the data gradient uses B = Wᵀ (a cached transpose of the frozen weight).
//
// Design (B200-first, not a translation of any library kernel):
//   * CTA tile 128 x BN (BN = 256/128/64), BLOCK_K = 64 bf16 = one 128-byte swizzle row.
//   * warp 0   : TMA producer  (cp.async.bulk.tensor.2d, SWIZZLE_128B, mbarrier complete_tx)
//   * warp 1   : MMA issuer    (one elected lane, tcgen05.mma.cta_group::1.kind::f16, M=128,N=BN,K=16)
//   * warp 2   : TMEM allocator (2 accumulator stages x BN fp32 columns)
//   * warps 4+ : epilogue      (tcgen05.ld 32x32b -> registers -> fused bias/act/residual -> 16-byte stores)
//   * three mbarrier pipelines: smem full/empty (TMA<->MMA), TMEM full/empty (MMA<->epilogue), and a
//     static persistent tile schedule (grid = #SMs, n-tile fastest so an A row-block stays in L2).
//   * the accumulator of tile i+1 is computed while the epilogue drains tile i (double-buffered TMEM).
#include "a4r_common.cuh"

namespace {

constexpr int BM = 128;
constexpr int BK = 64;               // 64 bf16 = 128 B = one swizzle-128B row
constexpr int UMMA_K = 16;           // fixed for 16-bit inputs
// Internal epilogue id (not part of the ABI enum): LINEAR whose dropout mask is applied to v + residual(s) instead of v —
// a4r_gemm_args.dropout_after_residual.  A separate instantiation, so the hot LINEAR kernel keeps its register budget.
constexpr int EPI_LINEAR_DROPSUM = 100;
// Epilogue warps come in groups of 4 (one warp per TMEM lane quadrant); group g owns the 32-column chunks c with
// c % groups == g.  Plain epilogues keep up with the MMAs with 2 groups; the GELU / GELU' epilogues of the 256-wide tile
// are instruction-latency-bound: GELU gets 3 groups (512 threads, 128 registers), GELU' 4 groups (640 threads, 96 registers);
// measured on B200 (M = 161,280, N = 3072, K = 768): GELU' 0.94 -> 0.84 ms with 4 groups, GELU 0.81 -> see profiles/.
template <int BN, int EPI>
struct EpiGroups {
  static constexpr int value = BN != 256 ? 2 : (EPI == A4R_EPI_DGELU ? 4 : (EPI == A4R_EPI_GELU ? 3 : 2));
};

// CG = CTAs per MMA (tcgen05 cta_group): with CG = 2 the two SMs of a TPC own one 256 x BN tile, each CTA staging its own
// 128 rows of A and HALF of the B rows (the pair's MMA reads both halves), which cuts shared-memory and L2->SM operand
// traffic per CTA by a third — the energy that bounds this kernel under the 1 kW cap.
// Epilogue boxes: every epilogue warp owns two 2 KB shared-memory boxes ([32 rows x 32 bf16], SWIZZLE_64B) that TMA fills or
// drains, because a 32-byte global access per lane touches 32 different lines per instruction (32 L1 wavefronts) and the
// K = 768 shapes — one 32-column chunk of a 256 x 256 tile every ~190 cycles per SM — were bound by the load/store unit:
//   * GELU (two OUTPUT tensors per chunk, pre-activation + activation): both leave through the boxes + TMA stores
//     (measured at M = 161,280, N = 3072, K = 768: 630 -> 562 us);
//   * LINEAR / LINEAR_DROPSUM / GELU' / RELU' (one streamed INPUT tensor — residual or pre-activation): the chunk is prefetched
//     two chunks ahead by a TMA load into the boxes instead of into registers; the single output keeps its direct 256-bit
//     stores (a box hand-over adds ~350 cycles of latency per chunk, which two epilogue groups do not hide: measured 192 ->
//     225 us on attention.output when the output went through boxes as well).
#ifndef A4R_GELU_AUX_BOX      // (A/B builds: which of the GELU forward's two outputs leave through the shared-memory boxes)
#define A4R_GELU_AUX_BOX 1
#endif
#ifndef A4R_GELU_C_BOX
#define A4R_GELU_C_BOX 1
#endif
template <bool B>
struct BoxTag {
  static constexpr bool value = B;
};
template <int BN, int EPI>
struct EpiBoxes {
  static constexpr bool kOut = EPI == A4R_EPI_GELU;
  static constexpr bool kIn = EPI == A4R_EPI_LINEAR || EPI == EPI_LINEAR_DROPSUM || EPI == A4R_EPI_DGELU || EPI == A4R_EPI_DRELU;
  static constexpr int kWarps = (kOut || kIn) ? 4 * EpiGroups<BN, EPI>::value : 0;
  static constexpr int kBytes = kWarps * 2 * 2048;
};
template <int BN, int CG, int EPI>
struct Cfg {
  static constexpr int kStageBytesA = BM * BK * 2;
  static constexpr int kStageBytesB = (BN / CG) * BK * 2;
  static constexpr int kStageBytes = kStageBytesA + kStageBytesB;
  static constexpr int kBoxBytes = EpiBoxes<BN, EPI>::kBytes;
  static constexpr int kWantStages = CG == 2 ? 6 : (BN == 256 ? 4 : (BN == 128 ? 6 : 8));
  static constexpr int kFitStages = (232448 - 1024 - 512 - kBoxBytes) / kStageBytes;      // 227 KB per CTA
  static constexpr int kStages = kFitStages < kWantStages ? kFitStages : kWantStages;
  static constexpr int kTmemCols = 2 * BN;  // 128 / 256 / 512: power of two >= 32
  static constexpr int kSmemBytes = kStages * kStageBytes + kBoxBytes + 1024 /*align slack*/ + 512 /*barriers*/;
  static_assert(kStages >= 3, "operand ring too shallow");
};

struct GemmParams {
  void* C;
  void* aux;
  const void* residual;
  const void* residual2;
  const float* bias;
  int64_t ldc, ldaux, ldr, ldr2;
  int M, N;
  int nk1, nk2;  // number of 64-wide k-blocks from (A,B) and from (A2,B2)
  int tiles_m, tiles_n;
  float alpha;
  int epilogue;
  int out_f32;
  // dropout on v = alpha*acc + bias (LINEAR mode only), applied BEFORE the residuals: element (row, col) of the logical
  // [M, N] output uses lane ((row*N + col) & 3) of rng64(seed, offset + (row*N + col) / 4) — the indexing of a4r_dropout
  uint32_t drop_thr16;
  float drop_scale;
  uint64_t drop_seed, drop_offset;
};

// ---- epilogue helpers ---------------------------------------------------------------------------
A4R_DEVICE void tma_store_2d(const CUtensorMap* m, uint32_t smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(reinterpret_cast<uint64_t>(m)),
               "r"(smem_src), "r"(c0), "r"(c1)
               : "memory");
}
A4R_DEVICE void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
A4R_DEVICE void bulk_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
A4R_DEVICE void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// One 32-column chunk of one output row, as 64 bytes of bf16 = 16 x b32 (or 2 x 256-bit / 4 x 128-bit accesses).
struct Row64 {
  uint32_t w[16];
};
template <bool V32>
A4R_DEVICE void load_row64(const __nv_bfloat16* p, Row64& r) {
  if constexpr (V32) {
    uint32_t a[8], b[8];
    ld_nc_v8(p, a);
    ld_nc_v8(p + 16, b);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      r.w[i] = a[i];
      r.w[8 + i] = b[i];
    }
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint4 v = ld_nc_v4(p + 8 * i);
      r.w[4 * i] = v.x; r.w[4 * i + 1] = v.y; r.w[4 * i + 2] = v.z; r.w[4 * i + 3] = v.w;
    }
  }
}
template <bool V32>
A4R_DEVICE void store_row64(__nv_bfloat16* p, const float (&v)[32]) {
  uint32_t w[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) w[i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);
  if constexpr (V32) {
    const uint32_t a[8] = {w[0], w[1], w[2], w[3], w[4], w[5], w[6], w[7]};
    const uint32_t b[8] = {w[8], w[9], w[10], w[11], w[12], w[13], w[14], w[15]};
    st_na_v8(p, a);
    st_na_v8(p + 16, b);
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i) st_na_v4(p + 8 * i, make_uint4(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]));
  }
}

// EPI: epilogue mode (compile time).  V32: every epilogue tensor is 32-byte aligned with ld % 16 == 0, so rows are
// moved with 256-bit accesses; the tail chunk of a ragged N falls back to guarded 128-bit accesses.
#ifdef A4R_GEMM_TIMING     // tools/micro/gemm_epi_timing.cu only
__device__ long long g_gemm_timing[8];
A4R_DEVICE long long clk() { long long t; asm volatile("mov.u64 %0, %%clock64;" : "=l"(t)::"memory"); return t; }
#define GSTAMP(v) const long long v = clk()
#else
#define GSTAMP(v)
#endif

template <int BN, int EPI, bool V32, int CG>
__global__ void __launch_bounds__(128 + 128 * EpiGroups<BN, EPI>::value, 1)
gemm_tn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmB2,
               const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmAux,
               const __grid_constant__ CUtensorMap tmIn, const GemmParams p) {
  constexpr bool kBoxes = EpiBoxes<BN, EPI>::kOut;       // results through boxes + TMA stores
  constexpr bool kInBoxes = EpiBoxes<BN, EPI>::kIn;      // streamed input through TMA loads + boxes
  using C = Cfg<BN, CG, EPI>;
  constexpr int NUM_EPI_GROUPS = EpiGroups<BN, EPI>::value;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* out_boxes = smem + C::kStages * C::kStageBytes;      // (input boxes for the kInBoxes epilogues)
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(out_boxes + C::kBoxBytes);
  uint64_t* empty_bar = full_bar + C::kStages;
  uint64_t* tmem_full_bar = empty_bar + C::kStages;
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);
  uint64_t* in_bar = tmem_empty_bar + 3;                         // [epilogue warp][2]: a streamed-input box has landed

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);   // (warp-uniform for the compiler)
  const int lane = threadIdx.x & 31;
  const int num_tiles = p.tiles_m * p.tiles_n;   // tiles of (128*CG) x BN
  const int nk = p.nk1 + p.nk2;
  const uint32_t cta_rank = CG == 2 ? cluster_ctarank() : 0u;
  const bool leader = cta_rank == 0;
  const int first_tile = CG == 2 ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
  const int tile_step = CG == 2 ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (p.nk2 > 0) {
      tma_prefetch_desc(&tmA2);
      tma_prefetch_desc(&tmB2);
    }
    if constexpr (kBoxes) {
      tma_prefetch_desc(&tmC);
      tma_prefetch_desc(&tmAux);
    }
    if constexpr (kInBoxes) tma_prefetch_desc(&tmIn);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < C::kStages; ++s) {
      mbar_init(&full_bar[s], 1);   // CG = 2: only the leader's copy is used (its expect_tx covers both CTAs' bytes)
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full_bar[s], 1);
      mbar_init(&tmem_empty_bar[s], 4 * NUM_EPI_GROUPS * CG);  // one arrive per epilogue warp of every CTA of the pair
    }
    if constexpr (kInBoxes) {
      for (int s = 0; s < 8 * NUM_EPI_GROUPS; ++s) mbar_init(&in_bar[s], 1);
    }
    mbar_fence_init();
  }
  if (warp == 2) {
    if constexpr (CG == 2) {
      tmem_alloc2(tmem_slot, C::kTmemCols);
      tmem_relinquish2();
    } else {
      tmem_alloc(tmem_slot, C::kTmemCols);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();   // barriers of BOTH CTAs are initialised past this point
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ============================== TMA producer: the WHOLE warp, converged (a4r_common.cuh: *_elect) ==============================
    {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = first_tile; tile < num_tiles; tile += tile_step) {
        const int m0 = (tile / p.tiles_n) * (BM * CG) + static_cast<int>(cta_rank) * BM;
        const int n0 = (tile % p.tiles_n) * BN + static_cast<int>(cta_rank) * (BN / CG) * (CG - 1);
        for (int kb = 0; kb < nk; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          __syncwarp();
          uint8_t* sa = smem + stage * C::kStageBytes;
          uint8_t* sb = sa + C::kStageBytesA;
          const CUtensorMap* ma = kb < p.nk1 ? &tmA : &tmA2;
          const CUtensorMap* mb = kb < p.nk1 ? &tmB : &tmB2;
          const int kc = (kb < p.nk1 ? kb : kb - p.nk1) * BK;
          if constexpr (CG == 2) {
            // both CTAs' bytes are credited to the LEADER's full barrier
            if (leader) mbar_expect_tx_elect(&full_bar[stage], C::kStageBytes * 2);
            const uint32_t bar = mapa_u32(smem_u32(&full_bar[stage]), 0);
            tma_load_2d_2cta_elect(ma, sa, bar, kc, m0);
            tma_load_2d_2cta_elect(mb, sb, bar, kc, n0);
          } else {
            tma_load_2d_elect(ma, sa, &full_bar[stage], kc, m0, C::kStageBytes);
            tma_load_2d_elect_noarm(mb, sb, &full_bar[stage], kc, n0);
          }
          if (++stage == C::kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ============================== MMA issuer (leader CTA only when CG = 2): the WHOLE warp, converged ==============================
    if (leader) {
      constexpr uint32_t idesc = umma_idesc_bf16(BM * CG, BN);
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int tile = first_tile; tile < num_tiles; tile += tile_step) {
        mbar_wait(&tmem_empty_bar[as], aphase ^ 1);
        __syncwarp();
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(as * BN);
        for (int kb = 0; kb < nk; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          __syncwarp();
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * C::kStageBytes);
          const uint64_t adesc = umma_desc_k_sw128(sa);
          const uint64_t bdesc = umma_desc_k_sw128(sa + C::kStageBytesA);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            // advance the start address by k*16 elements (32 B) inside the 128 B swizzle row
            if constexpr (CG == 2)
              umma_bf16_ss_2cta_elect(tmem_d, adesc + static_cast<uint64_t>(k * 2), bdesc + static_cast<uint64_t>(k * 2), idesc,
                                      (kb | k) != 0 ? 1u : 0u);
            else
              umma_bf16_ss_elect(tmem_d, adesc + static_cast<uint64_t>(k * 2), bdesc + static_cast<uint64_t>(k * 2), idesc,
                                 (kb | k) != 0 ? 1u : 0u);
          }
          // frees the smem slot (in both CTAs of the pair) when these MMAs retire
          if constexpr (CG == 2) umma_commit_2cta_elect(&empty_bar[stage], 3); else umma_commit_elect(&empty_bar[stage]);
          if (++stage == C::kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        // accumulator ready for the epilogue warps (of both CTAs)
        if constexpr (CG == 2) umma_commit_2cta_elect(&tmem_full_bar[as], 3); else umma_commit_elect(&tmem_full_bar[as]);
        if (++as == 2) {
          as = 0;
          aphase ^= 1;
        }
      }
    }
  } else if (warp >= 4) {
    // ============================== epilogue ==============================
    // The (at most one) streamed input tensor of the mode — residual for LINEAR, aux for DGELU/DRELU — is
    // prefetched into registers TWO chunks ahead, the first two chunks of a tile before the accumulator barrier,
    // so that its DRAM latency hides behind the MMA of the tile instead of stalling the 8 epilogue warps.
    // The chunk loop is deliberately NOT fully unrolled (two chunk bodies per iteration): a fully unrolled GELU
    // epilogue is > 60 KB of SASS and thrashes the 32 KB instruction cache.
    const int ew = warp - 4;
    const int quad = warp & 3;  // TMEM lane quadrant this warp may read
    const int group = ew >> 2;
    constexpr int CHUNKS = (BN / 32 + NUM_EPI_GROUPS - 1) / NUM_EPI_GROUPS;
    constexpr bool kLinear = (EPI == A4R_EPI_LINEAR) || (EPI == EPI_LINEAR_DROPSUM);
    constexpr bool kHasIn = kLinear || (EPI == A4R_EPI_DGELU) || (EPI == A4R_EPI_DRELU);
    const __nv_bfloat16* in_ptr = nullptr;
    int64_t in_ld = 0;
    if constexpr (kLinear) {
      in_ptr = reinterpret_cast<const __nv_bfloat16*>(p.residual);
      in_ld = p.ldr;
    } else if constexpr (kHasIn) {
      in_ptr = reinterpret_cast<const __nv_bfloat16*>(p.aux);
      in_ld = p.ldaux;
    }
    const bool has_in = kHasIn && in_ptr != nullptr;
    int as = 0;
    uint32_t aphase = 0;
    uint32_t nbox = 0;             // result boxes handed to TMA so far (the warp's two boxes alternate)
    uint32_t n_in[2] = {0u, 0u};   // input boxes consumed per slot (mbarrier phase)
    const bool out_f32 = p.out_f32 != 0;
#ifdef A4R_GEMM_TIMING
    long long tacc[6] = {0, 0, 0, 0, 0, 0};
    const long long gstart = clk();
#endif
    for (int tile = first_tile; tile < num_tiles; tile += tile_step) {
      const int m0 = (tile / p.tiles_n) * (BM * CG) + static_cast<int>(cta_rank) * BM;
      const int n0 = (tile % p.tiles_n) * BN;
      const int row = m0 + quad * 32 + lane;
      const bool row_ok = row < p.M;
      const int64_t r64 = row;
      auto chunk_col = [&](int ci) { return n0 + (group + ci * NUM_EPI_GROUPS) * 32; };
      // the streamed input of chunk ci: one TMA load of the warp's [32 rows x 32 columns] box into slot `slot` (rows past M and
      // columns past N arrive as zeros); issued two chunks ahead, after the slot's previous contents have been consumed
      auto prefetch = [&](int ci, int slot) {
        if constexpr (kInBoxes) {
          if (!has_in || ci >= CHUNKS || (group + ci * NUM_EPI_GROUPS) >= BN / 32 || chunk_col(ci) >= p.N) return;   // warp-uniform
          __syncwarp();
          tma_load_2d_elect(&tmIn, out_boxes + (ew * 2 + slot) * 2048, &in_bar[ew * 2 + slot], chunk_col(ci), m0 + quad * 32, 2048);
        }
      };
      // one 32-column chunk: TMEM -> registers -> fused op -> global
      auto body = [&](int ci, int slot) {
        const int c = group + ci * NUM_EPI_GROUPS;
        if (c >= BN / 32) return;
        const int col0 = n0 + c * 32;
        if (col0 >= p.N) return;  // warp-uniform
        uint32_t acc[32];
        GSTAMP(g0);
        tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) +
                               static_cast<uint32_t>(as * BN + c * 32),
                           acc);
        const bool full = col0 + 32 <= p.N;
        // the bias slice of this chunk is fetched WHILE the TMEM load is in flight (both latencies overlap)
        constexpr bool kBias = kLinear || (EPI == A4R_EPI_GELU) || (EPI == A4R_EPI_RELU);
        float4 bv[kBias ? 8 : 1];
        const bool has_bias = kBias && p.bias != nullptr;
        if (has_bias) {
#pragma unroll
          for (int j = 0; j < (kBias ? 8 : 0); ++j) {
            bv[j] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (full || col0 + 4 * j < p.N) bv[j] = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + 4 * j));
          }
        }
        tmem_ld_wait();
        GSTAMP(g1);
#ifdef A4R_GEMM_TIMING
        tacc[0] += g1 - g0;
#endif
        // (with result boxes rows past M are not skipped: their accumulators are zeros — A is zero-filled there — TMA clips
        // their stores, and the whole warp has to reach the box hand-over below)
        float2 v[16];
        const float2 al = splat2(p.alpha);
        if (has_bias) {
#pragma unroll
          for (int j = 0; j < (kBias ? 8 : 0); ++j) {
            v[2 * j] = __ffma2_rn(al, make_float2(__uint_as_float(acc[4 * j]), __uint_as_float(acc[4 * j + 1])),
                                  make_float2(bv[j].x, bv[j].y));
            v[2 * j + 1] = __ffma2_rn(al, make_float2(__uint_as_float(acc[4 * j + 2]), __uint_as_float(acc[4 * j + 3])),
                                      make_float2(bv[j].z, bv[j].w));
          }
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i)
            v[i] = __fmul2_rn(al, make_float2(__uint_as_float(acc[2 * i]), __uint_as_float(acc[2 * i + 1])));
        }
        // streamed input of this chunk: this lane's row of the prefetched box (64 bytes, 64-byte swizzle)
        Row64 in;
        if constexpr (kInBoxes) {
          if (has_in) {
            mbar_wait(&in_bar[ew * 2 + slot], n_in[slot] & 1u);
            ++n_in[slot];
            const uint32_t box = smem_u32(out_boxes) + static_cast<uint32_t>((ew * 2 + slot) * 2048) + lane * 64;
            const uint32_t sw = static_cast<uint32_t>((lane >> 1) & 3);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint4 t = lds_v4(box + ((static_cast<uint32_t>(j) ^ sw) << 4));
              in.w[4 * j] = t.x; in.w[4 * j + 1] = t.y; in.w[4 * j + 2] = t.z; in.w[4 * j + 3] = t.w;
            }
          }
        }
        if constexpr (!kBoxes) {
          if (!row_ok) return;      // (after the box wait: every lane keeps the slot's phase count)
        }
        // bf16 results leave through the warp's private [32 rows x 32 columns] SWIZZLE_64B box and ONE TMA store: a 32-byte
        // st.global per lane touches 32 different lines per instruction (32 L1 wavefronts), and with two such tensors per chunk
        // (residual / pre-activation + output) the K = 768 shapes were bound by the load/store unit, not by the MMAs.  The two
        // boxes of a warp alternate, so a box is rewritten only after the store issued two hand-overs ago has read it.
        auto store_bf16 = [&](const CUtensorMap* tm, __nv_bfloat16* dst, auto box_tag) {
          constexpr bool kUseBox = kBoxes && decltype(box_tag)::value;
          uint32_t w[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) w[i] = pack_bf16x2(v[i].x, v[i].y);
          if constexpr (kUseBox) {
            const uint32_t box = smem_u32(out_boxes) + static_cast<uint32_t>((ew * 2 + (nbox & 1)) * 2048);
            __syncwarp();
            bulk_wait_read1_elect();
            __syncwarp();
            const uint32_t sw = static_cast<uint32_t>((lane >> 1) & 3);
#pragma unroll
            for (int j = 0; j < 4; ++j)
              sts_v4(box + lane * 64 + ((static_cast<uint32_t>(j) ^ sw) << 4), w[4 * j], w[4 * j + 1], w[4 * j + 2], w[4 * j + 3]);
            fence_proxy_async_smem();
            __syncwarp();
            tma_store_2d_commit_elect(tm, box, col0, m0 + quad * 32);   // columns past N and rows past M are clipped by the tensor map
            ++nbox;
          } else {
            if (!row_ok) return;
            if (full) {
              if constexpr (V32) {
                const uint32_t a8[8] = {w[0], w[1], w[2], w[3], w[4], w[5], w[6], w[7]};
                const uint32_t b8[8] = {w[8], w[9], w[10], w[11], w[12], w[13], w[14], w[15]};
                st_na_v8(dst, a8);
                st_na_v8(dst + 16, b8);
              } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) st_na_v4(dst + 8 * j, make_uint4(w[4 * j], w[4 * j + 1], w[4 * j + 2], w[4 * j + 3]));
              }
            } else {
#pragma unroll
              for (int j = 0; j < 4; ++j)
                if (col0 + 8 * j < p.N) st_na_v4(dst + 8 * j, make_uint4(w[4 * j], w[4 * j + 1], w[4 * j + 2], w[4 * j + 3]));
            }
          }
        };
        auto drop_v = [&]() {
          const uint64_t ctr = p.drop_offset + ((static_cast<uint64_t>(r64) * static_cast<uint64_t>(p.N) + col0) >> 2);
          const uint64_t seed = rng_seed(p.drop_seed);
#pragma unroll
            for (int q = 0; q < 8; ++q) {   // 4 consecutive columns per counter
              const uint64_t r = rng64(seed, ctr + q);
              v[2 * q].x = rng_keep(r, 0, p.drop_thr16) ? v[2 * q].x * p.drop_scale : 0.0f;
              v[2 * q].y = rng_keep(r, 1, p.drop_thr16) ? v[2 * q].y * p.drop_scale : 0.0f;
              v[2 * q + 1].x = rng_keep(r, 2, p.drop_thr16) ? v[2 * q + 1].x * p.drop_scale : 0.0f;
              v[2 * q + 1].y = rng_keep(r, 3, p.drop_thr16) ? v[2 * q + 1].y * p.drop_scale : 0.0f;
            }
        };
        if constexpr (kLinear) {
          if constexpr (EPI == A4R_EPI_LINEAR) {
            if (p.drop_thr16 != 0) drop_v();                      // forward: dropout(x Wᵀ + b) + residual
          }
          if (has_in) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = __fadd2_rn(v[i], unpack_bf16x2(in.w[i]));
          }
          if (p.residual2 != nullptr) {
            const __nv_bfloat16* r2 = reinterpret_cast<const __nv_bfloat16*>(p.residual2) + r64 * p.ldr2 + col0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              if (row_ok && (full || col0 + 8 * j < p.N)) {
                const uint4 t = ld_nc_v4(r2 + 8 * j);
                v[4 * j] = __fadd2_rn(v[4 * j], unpack_bf16x2(t.x));
                v[4 * j + 1] = __fadd2_rn(v[4 * j + 1], unpack_bf16x2(t.y));
                v[4 * j + 2] = __fadd2_rn(v[4 * j + 2], unpack_bf16x2(t.z));
                v[4 * j + 3] = __fadd2_rn(v[4 * j + 3], unpack_bf16x2(t.w));
              }
            }
          }
          // gradient through a dropout whose input gradient is the sum formed above: mask the SUM
          if constexpr (EPI == EPI_LINEAR_DROPSUM) drop_v();
        } else if constexpr (EPI == A4R_EPI_GELU) {
          if (p.aux != nullptr) store_bf16(&tmAux, reinterpret_cast<__nv_bfloat16*>(p.aux) + r64 * p.ldaux + col0, BoxTag<A4R_GELU_AUX_BOX != 0>{});
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = gelu_fast2(v[i]);
        } else if constexpr (EPI == A4R_EPI_RELU) {
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = make_float2(fmaxf(v[i].x, 0.0f), fmaxf(v[i].y, 0.0f));
        } else if constexpr (EPI == A4R_EPI_DGELU) {
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = __fmul2_rn(v[i], gelu_grad_fast2(unpack_bf16x2(in.w[i])));
        } else if constexpr (EPI == A4R_EPI_DRELU) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float2 u = unpack_bf16x2(in.w[i]);
            v[i] = make_float2(u.x > 0.0f ? v[i].x : 0.0f, u.y > 0.0f ? v[i].y : 0.0f);
          }
        }
        if (out_f32) {
          float* cp = reinterpret_cast<float*>(p.C) + r64 * p.ldc + col0;
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (row_ok && (full || col0 + 4 * j < p.N))
              st_na_v4(cp + 4 * j, make_uint4(__float_as_uint(v[2 * j].x), __float_as_uint(v[2 * j].y),
                                              __float_as_uint(v[2 * j + 1].x), __float_as_uint(v[2 * j + 1].y)));
        } else {
          GSTAMP(g2);
          store_bf16(&tmC, reinterpret_cast<__nv_bfloat16*>(p.C) + r64 * p.ldc + col0, BoxTag<A4R_GELU_C_BOX != 0>{});
          GSTAMP(g3);
#ifdef A4R_GEMM_TIMING
          tacc[1] += g2 - g1;
          tacc[2] += g3 - g2;
          tacc[3] += 1;
#endif
        }
      };
      prefetch(0, 0);
      prefetch(1, 1);
      GSTAMP(g4);
      mbar_wait(&tmem_full_bar[as], aphase);
      GSTAMP(g5);
#ifdef A4R_GEMM_TIMING
      tacc[4] += g5 - g4;
      tacc[5] += 1;
#endif
      tc_fence_after();
#pragma unroll 1
      for (int cp = 0; cp < CHUNKS; cp += 2) {
        body(cp, 0);
        prefetch(cp + 2, 0);
        if (cp + 1 < CHUNKS) {
          body(cp + 1, 1);
          prefetch(cp + 3, 1);
        }
      }
      // all of this warp's TMEM reads for this stage are complete (wait::ld in body): release it
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (CG == 2) mbar_arrive_remote(mapa_u32(smem_u32(&tmem_empty_bar[as]), 0));  // the leader's barrier
        else mbar_arrive(&tmem_empty_bar[as]);
      }
      if (++as == 2) {
        as = 0;
        aphase ^= 1;
      }
    }
    if constexpr (kBoxes) {
      __syncwarp();
      bulk_wait0_elect();            // the last result boxes have been written out before the CTA retires
    }
#ifdef A4R_GEMM_TIMING
    if (blockIdx.x == 0 && ew == 0 && lane == 0) {
      for (int i = 0; i < 6; ++i) g_gemm_timing[i] = tacc[i];
      g_gemm_timing[6] = clk() - gstart;
    }
#endif
  }

  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  if (warp == 2) {
    if constexpr (CG == 2) tmem_dealloc2(tmem_base, C::kTmemCols); else tmem_dealloc(tmem_base, C::kTmemCols);
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

// bf16 row-major [rows, cols] with leading dimension ld (elements); box = 64 cols x box_rows, 128B swizzle
int make_tmap(CUtensorMap* m, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
  PFN_encodeTiled fn = get_encode_fn();
  if (fn == nullptr) return a4r_set_error(A4R_ECUDA, "cuTensorMapEncodeTiled entry point not found");
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 2};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(BK), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return a4r_set_error(A4R_ECUDA, "cuTensorMapEncodeTiled failed (%d) rows=%lld cols=%lld ld=%lld", (int)r,
                         (long long)rows, (long long)cols, (long long)ld);
  return A4R_OK;
}

// bf16 row-major [rows, cols] output (leading dimension ld): box = 32 columns x 32 rows, 64-byte swizzle — one epilogue warp's
// chunk; elements outside [rows, cols] are not written
int make_tmap_out(CUtensorMap* m, const void* base, int64_t rows, int64_t cols, int64_t ld) {
  PFN_encodeTiled fn = get_encode_fn();
  if (fn == nullptr) return a4r_set_error(A4R_ECUDA, "cuTensorMapEncodeTiled entry point not found");
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 2};
  cuuint32_t box[2] = {32u, 32u};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return a4r_set_error(A4R_ECUDA, "cuTensorMapEncodeTiled (output box) failed (%d) rows=%lld cols=%lld ld=%lld", (int)r,
                         (long long)rows, (long long)cols, (long long)ld);
  return A4R_OK;
}

template <int BN, int EPI, bool V32, int CG>
int launch_gemm(const a4r_gemm_args* a, cudaStream_t stream) {
  constexpr bool kBoxes = EpiBoxes<BN, EPI>::kOut;
  constexpr bool kInBoxes = EpiBoxes<BN, EPI>::kIn;
  using C = Cfg<BN, CG, EPI>;
  CUtensorMap tmA, tmB, tmA2, tmB2, tmC, tmAux, tmIn;
  int rc;
  {
    // the streamed epilogue input: residual (LINEAR) or aux (GELU' / RELU')
    const void* in_ptr = (EPI == A4R_EPI_LINEAR || EPI == EPI_LINEAR_DROPSUM) ? a->residual : a->aux;
    const int64_t in_ld = (EPI == A4R_EPI_LINEAR || EPI == EPI_LINEAR_DROPSUM) ? a->ldr : a->ldaux;
    if (kInBoxes && in_ptr != nullptr) {
      if ((rc = make_tmap_out(&tmIn, in_ptr, a->M, a->N, in_ld)) != A4R_OK) return rc;
    }
  }
  const bool box_c = kBoxes && !a->out_f32, box_aux = kBoxes && a->aux != nullptr;
  if (box_c && (rc = make_tmap_out(&tmC, a->C, a->M, a->N, a->ldc)) != A4R_OK) return rc;
  if (box_aux && (rc = make_tmap_out(&tmAux, a->aux, a->M, a->N, a->ldaux)) != A4R_OK) return rc;
  if ((rc = make_tmap(&tmA, a->A, a->M, a->K, a->lda, BM)) != A4R_OK) return rc;
  if ((rc = make_tmap(&tmB, a->B, a->N, a->K, a->ldb, BN / CG)) != A4R_OK) return rc;
  if (a->K2 > 0) {
    if ((rc = make_tmap(&tmA2, a->A2, a->M, a->K2, a->lda2, BM)) != A4R_OK) return rc;
    if ((rc = make_tmap(&tmB2, a->B2, a->N, a->K2, a->ldb2, BN / CG)) != A4R_OK) return rc;
  } else {
    tmA2 = tmA;
    tmB2 = tmB;
  }
  if (!box_c) tmC = tmA;        // (unused: stored directly)
  if (!box_aux) tmAux = tmA;    // (unused)
  if (!(kInBoxes && ((EPI == A4R_EPI_LINEAR || EPI == EPI_LINEAR_DROPSUM) ? a->residual : a->aux) != nullptr)) tmIn = tmA;   // (unused)
  GemmParams p;
  p.C = a->C;
  p.aux = a->aux;
  p.residual = a->residual;
  p.residual2 = a->residual2;
  p.bias = a->bias;
  p.ldc = a->ldc;
  p.ldaux = a->ldaux;
  p.ldr = a->ldr;
  p.ldr2 = a->ldr2;
  p.M = static_cast<int>(a->M);
  p.N = static_cast<int>(a->N);
  p.nk1 = static_cast<int>((a->K + BK - 1) / BK);
  p.nk2 = static_cast<int>((a->K2 + BK - 1) / BK);
  p.tiles_m = static_cast<int>((a->M + BM * CG - 1) / (BM * CG));
  p.tiles_n = static_cast<int>((a->N + BN - 1) / BN);
  p.alpha = a->alpha;
  p.epilogue = a->epilogue;
  p.out_f32 = a->out_f32;
  p.drop_thr16 = static_cast<uint32_t>(a->dropout_p * 65536.0f + 0.5f);
  p.drop_scale = 65536.0f / static_cast<float>(65536u - p.drop_thr16);
  p.drop_seed = a->dropout_seed;
  p.drop_offset = a->dropout_offset;

  static bool attr_set = false;
  if (!attr_set) {
    A4R_CUDA_OK(cudaFuncSetAttribute(gemm_tn_kernel<BN, EPI, V32, CG>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     C::kSmemBytes));
    attr_set = true;
  }
  const int64_t tiles = static_cast<int64_t>(p.tiles_m) * p.tiles_n;
  const int units = a4r_num_sms() / CG;   // CTAs (CG = 1) or CTA pairs (CG = 2) resident at once
  const int grid = static_cast<int>(tiles < units ? tiles : units) * CG;
  const int threads = 128 + 128 * EpiGroups<BN, EPI>::value;
  if constexpr (CG == 2) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = C::kSmemBytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    A4R_CUDA_OK(cudaLaunchKernelEx(&cfg, gemm_tn_kernel<BN, EPI, V32, CG>, tmA, tmB, tmA2, tmB2, tmC, tmAux, tmIn, p));
  } else {
    gemm_tn_kernel<BN, EPI, V32, CG><<<grid, threads, C::kSmemBytes, stream>>>(tmA, tmB, tmA2, tmB2, tmC, tmAux, tmIn, p);
  }
  A4R_LAUNCH_OK();
  a4r_count_launch(1);
  return A4R_OK;
}

}  // namespace

extern "C" int a4r_gemm_bf16_tn(const a4r_gemm_args* a, a4r_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  A4R_CHECK_ARG(a != nullptr, "gemm: args is NULL");
  A4R_CHECK_ARG(a->A && a->B && a->C, "gemm: A, B and C must be non-NULL");
  A4R_CHECK_ARG(a->M >= 0 && a->N > 0 && a->K > 0, "gemm: bad M/N/K (%lld,%lld,%lld)", (long long)a->M,
                (long long)a->N, (long long)a->K);
  A4R_CHECK_ARG(a->M < (1ll << 31) && a->N < (1ll << 31) && a->K < (1ll << 31), "gemm: dims must fit int32");
  A4R_CHECK_ARG(a->K % 8 == 0 && a->N % 8 == 0, "gemm: K and N must be multiples of 8 (K=%lld N=%lld)",
                (long long)a->K, (long long)a->N);
  A4R_CHECK_ARG(a->lda % 8 == 0 && a->ldb % 8 == 0 && a->lda >= a->K && a->ldb >= a->K,
                "gemm: lda/ldb must be >= K and multiples of 8");
  A4R_CHECK_ARG(a->ldc >= a->N && a->ldc % (a->out_f32 ? 4 : 8) == 0, "gemm: bad ldc");
  A4R_CHECK_ARG(a4r_aligned16(a->A) && a4r_aligned16(a->B) && a4r_aligned16(a->C), "gemm: A/B/C must be 16B aligned");
  A4R_CHECK_ARG(a->K2 >= 0 && a->K2 % 8 == 0, "gemm: K2 must be a non-negative multiple of 8");
  if (a->K2 > 0) {
    A4R_CHECK_ARG(a->A2 && a->B2 && a4r_aligned16(a->A2) && a4r_aligned16(a->B2), "gemm: A2/B2 missing or unaligned");
    A4R_CHECK_ARG(a->lda2 % 8 == 0 && a->ldb2 % 8 == 0 && a->lda2 >= a->K2 && a->ldb2 >= a->K2, "gemm: bad lda2/ldb2");
  }
  A4R_CHECK_ARG(a->epilogue >= A4R_EPI_LINEAR && a->epilogue <= A4R_EPI_DRELU, "gemm: unknown epilogue %d",
                a->epilogue);
  if (a->epilogue == A4R_EPI_DGELU || a->epilogue == A4R_EPI_DRELU)
    A4R_CHECK_ARG(a->aux != nullptr && a->bias == nullptr, "gemm: DGELU/DRELU need aux and take no bias");
  if (a->aux) A4R_CHECK_ARG(a4r_aligned16(a->aux) && a->ldaux % 8 == 0 && a->ldaux >= a->N, "gemm: bad aux/ldaux");
  if (a->residual)
    A4R_CHECK_ARG(a4r_aligned16(a->residual) && a->ldr % 8 == 0 && a->ldr >= a->N, "gemm: bad residual/ldr");
  if (a->residual2)
    A4R_CHECK_ARG(a4r_aligned16(a->residual2) && a->ldr2 % 8 == 0 && a->ldr2 >= a->N, "gemm: bad residual2/ldr2");
  if (a->bias) A4R_CHECK_ARG(a4r_aligned16(a->bias), "gemm: bias must be 16B aligned");
  A4R_CHECK_ARG(a->dropout_p >= 0.0f && a->dropout_p < 1.0f, "gemm: dropout_p must be in [0,1)");
  A4R_CHECK_ARG(a->dropout_p == 0.0f || a->epilogue == A4R_EPI_LINEAR, "gemm: dropout is a LINEAR-epilogue option");
  int rc = a4r_device_check();
  if (rc != A4R_OK) return rc;
  if (a->M == 0) return A4R_OK;

  int bn = a->block_n;
  // auto: wide outputs run the 256 x 256 CTA-pair tile (cta_group::2: measured +5..15 % over the single-CTA 128 x 256 tile
  // and above cuBLAS on the K >= 2304 shapes), narrow ones a single-CTA 128 x {128, 64} tile
  if (bn == 0) bn = a->N > 128 ? (a->M > 128 ? 512 : 256) : (a->N > 64 ? 128 : 64);
  // block_n = 512 selects the 256-wide tile on a CTA pair (cta_group::2, 256 x 256 per pair)
  const bool pair = bn == 512;
  if (pair) bn = 256;
  if (bn != 64 && bn != 128 && bn != 256)
    return a4r_set_error(A4R_EINVAL, "gemm: block_n must be 0, 64, 128, 256 or 512 (got %d)", a->block_n);
  // 256-bit epilogue accesses need 32-byte aligned rows of every bf16 epilogue tensor
  auto ok32 = [](const void* ptr, int64_t ld) { return ptr == nullptr || ((reinterpret_cast<uintptr_t>(ptr) & 31u) == 0 && ld % 16 == 0); };
  const bool v32 = !a->out_f32 && ok32(a->C, a->ldc) && ok32(a->aux, a->ldaux) && ok32(a->residual, a->ldr);
#define A4R_DISPATCH_BN(EPI, V)                                                                       \
  (pair ? launch_gemm<256, EPI, V, 2>(a, stream)                                                       \
        : (bn == 256 ? launch_gemm<256, EPI, V, 1>(a, stream)                                          \
                     : (bn == 128 ? launch_gemm<128, EPI, V, 1>(a, stream) : launch_gemm<64, EPI, V, 1>(a, stream))))
#define A4R_DISPATCH(EPI) (v32 ? A4R_DISPATCH_BN(EPI, true) : A4R_DISPATCH_BN(EPI, false))
  if (a->epilogue == A4R_EPI_LINEAR && a->dropout_after_residual != 0 && a->dropout_p > 0.0f)
    return A4R_DISPATCH(EPI_LINEAR_DROPSUM);
  switch (a->epilogue) {
    case A4R_EPI_LINEAR: return A4R_DISPATCH(A4R_EPI_LINEAR);
    case A4R_EPI_GELU: return A4R_DISPATCH(A4R_EPI_GELU);
    case A4R_EPI_RELU: return A4R_DISPATCH(A4R_EPI_RELU);
    case A4R_EPI_DGELU: return A4R_DISPATCH(A4R_EPI_DGELU);
    default: return A4R_DISPATCH(A4R_EPI_DRELU);
  }
#undef A4R_DISPATCH
#undef A4R_DISPATCH_BN
}
