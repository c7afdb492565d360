// adapter_ln_sm100.cu — K5: the Houlsby adapter block in ONE pass over the hidden states.
//
//   a   = h + W_u · act(W_d · h + b_d) + b_u            AdapterBlock.forward, Downstream/Text/model/modules.py:131-134
//   out = LayerNorm(a + input)                          BertAdaptedSelfOutput.forward, model.py:292-297          (tail 0)
//   out = a + input                                     VITAdaptedOutput.forward,  Downstream/CV/model/model.py  (tail 1)
//   out = a                                             VITAdaptedSelfOutput.forward                             (tail 2)
//
// Per 128-token tile, one persistent CTA per SM:
//   warp 0    TMA producer: streams h in [128 x 64] k-blocks (+ the matching [r x 64] block of W_d) through an
//             mbarrier ring; W_u ([H x r], 96 KB at H = 768) is loaded once per CTA and stays resident.
//   warp 1    tcgen05 issuer: S1[128 x 64] = h · W_dᵀ (TMEM), then — once the epilogue has written act(S1 + b_d) as a
//             bf16 SWIZZLE_128B tile to shared memory — U = s · W_uᵀ in column chunks of UC <= 192 through two TMEM
//             stages.  The down-projection of the NEXT tile is issued between the up-projection chunks, so its HBM
//             stream overlaps the epilogue of the current tile.
//   warps 4-11 epilogue.  TMEM hands a thread one token ROW, which is the wrong shape for HBM (32 lanes on 32 different
//             rows = 32 half-used sectors per instruction: the first version of this kernel spent 85 % of its time in
//             the LSU).  So every 64-column chunk of U goes TMEM -> registers -> an fp32 staging tile in shared memory
//             (XOR-swizzled, double-buffered, one named barrier per chunk) and is then consumed ROW-DISTRIBUTED: 16 lanes
//             cover the 128 contiguous bytes of one row, a warp instruction touches two full 128-byte lines.  In that
//             layout the threads read h and input straight from global memory (prefetched one chunk ahead in registers),
//             form z = U + b_u + h + input, round it to bf16, write it and accumulate the row statistics; the LayerNorm
//             pass re-reads z (each thread exactly the bytes it wrote: L2 hits, no cross-thread ordering needed) and
//             writes LN(z).  s (the bottleneck activation the backward needs) leaves through the same operand tile.
// HBM traffic per token (H = 768): read h + input, write out = 4,608 B (+ z = 1,536 B and s = 2r B when the backward
// needs them) against ~9.5 KB for the composition GEMM -> GEMM(+2 residuals) -> LayerNorm; the contraction FLOPs
// (2 x 2 x 768 x 64 per token) are ~5 % of what the tensor pipe could do in the time HBM needs, so the bound is HBM.
#include <stdlib.h>

#include "a4r_common.cuh"

namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int RP = 64;                       // adapter rank padded to one k-block (r <= 64)
constexpr int MAX_STAGE = 4;
constexpr int STAGE_A = BM * BK * 2;         // 16 KB of h
constexpr int STAGE_B = RP * BK * 2;         // 8 KB of W_d
constexpr int STAGE_BYTES = STAGE_A + STAGE_B;
constexpr int S_TILE = BM * RP * 2;          // 16 KB
constexpr int CW = 64;                       // epilogue chunk width (columns)
constexpr int UST_BYTES = BM * CW * 4;       // 32 KB: one fp32 staging tile
constexpr int EPI_WARPS = 16;
constexpr int EPI_THREADS = EPI_WARPS * 32;
constexpr int THREADS = 128 + EPI_THREADS;
constexpr int TMEM_COLS = 512;
constexpr int S1_COL = 448;                  // S1 lives in columns [448, 512); U stages at 0 and UC (2 * UC <= 384)
constexpr int PASSES = BM / (EPI_WARPS * 2); // 4 row passes per chunk: a warp covers 2 rows per instruction
constexpr int PASS_ROWS = EPI_WARPS * 2;     // rows between two passes of one thread

struct AdParams {
  const __nv_bfloat16* h;
  const __nv_bfloat16* input;
  int64_t ldh, ldi;
  const float* b_down;
  const float* b_up;
  const float* gamma;
  const float* beta;
  __nv_bfloat16* out;
  __nv_bfloat16* z_out;
  float* mean;
  float* rstd;
  __nv_bfloat16* s_out;
  __nv_bfloat16* u_out;
  int M, H, r, UC, nstage;
  int lds;      // row stride of s_out (elements); lds >= r + 8: columns [r, r + 8) of s_out receive [1, 0, ..., 0]
  int act, tail;
  float eps;
};

A4R_DEVICE void tma_prefetch_l2_2d(const CUtensorMap* m, int c0, int c1, uint64_t pol) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.L2::cache_hint [%0, {%1, %2}], %3;" ::"l"(reinterpret_cast<uint64_t>(m)),
               "r"(c0), "r"(c1), "l"(pol)
               : "memory");
}
A4R_DEVICE void named_bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }

// L2 residency control.  One tile-time of the whole GPU moves ~140 MB through the 126 MB L2, so lines that are re-used one
// phase later (h: TMA read for the down-projection, re-read as a residual; z: written, re-read by the LayerNorm pass; input:
// prefetched a tile ahead) are tagged evict_last when they enter, and evict_first by their final access.
A4R_DEVICE uint64_t l2_policy_evict_last() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
A4R_DEVICE uint64_t l2_policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
A4R_DEVICE uint2 ld_nc_v2(const void* p, uint64_t pol) {   // final read of a residual stream
  uint2 r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.u32 {%0,%1}, [%2], %3;" : "=r"(r.x), "=r"(r.y) : "l"(p), "l"(pol));
  return r;
}
A4R_DEVICE uint2 ld_cg_v2(const void* p, uint64_t pol) {   // the LayerNorm pass re-reads what this thread wrote: L2, never L1
  uint2 r;
  asm volatile("ld.global.cg.L2::cache_hint.v2.u32 {%0,%1}, [%2], %3;" : "=r"(r.x), "=r"(r.y) : "l"(p), "l"(pol) : "memory");
  return r;
}
A4R_DEVICE void st_v2(void* p, uint32_t a, uint32_t b, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.v2.u32 [%0], {%1,%2}, %3;" ::"l"(p), "r"(a), "r"(b), "l"(pol) : "memory");
}
A4R_DEVICE void tma_load_2d_hint(const CUtensorMap* m, void* smem_dst, uint64_t* bar, int c0, int c1, uint64_t pol) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "l"(pol)
      : "memory");
}

__global__ void __launch_bounds__(THREADS, 1)
adapter_ln_kernel(const __grid_constant__ CUtensorMap tmH, const __grid_constant__ CUtensorMap tmWd,
                  const __grid_constant__ CUtensorMap tmWu, const __grid_constant__ CUtensorMap tmI, const AdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* s_wu = smem;                                    // [H][64] bf16, SW128
  uint8_t* s_ring = s_wu + static_cast<size_t>(p.H) * 128;
  uint8_t* s_act = s_ring + p.nstage * STAGE_BYTES;        // [128][64] bf16, SW128: A operand of the up-projection
  uint8_t* s_ust = s_act + S_TILE;                         // 2 x [128][64] f32 staging tiles (16-byte chunks XOR-swizzled)
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(s_ust + 2 * UST_BYTES);
  uint64_t* empty_bar = full_bar + MAX_STAGE;
  uint64_t* wu_bar = empty_bar + MAX_STAGE;
  uint64_t* s1_full = wu_bar + 1;
  uint64_t* s_ready = s1_full + 1;
  uint64_t* u_full = s_ready + 1;    // [2]
  uint64_t* u_empty = u_full + 2;    // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(u_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_tiles = (p.M + BM - 1) / BM;
  const int nkb = p.H / BK;
  const int nch = p.H / p.UC;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmH);
    tma_prefetch_desc(&tmWd);
    tma_prefetch_desc(&tmWu);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.nstage; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(wu_bar, 1);
    mbar_init(s1_full, 1);
    mbar_init(s_ready, EPI_WARPS);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&u_full[s], 1);
      mbar_init(&u_empty[s], EPI_WARPS);
    }
    mbar_fence_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ============================== TMA producer ==============================
    if (lane == 0) {
      mbar_expect_tx(wu_bar, static_cast<uint32_t>(p.H) * 128u);
      for (int c = 0; c < nch; ++c) tma_load_2d(&tmWu, s_wu + static_cast<size_t>(c) * p.UC * 128, wu_bar, 0, c * p.UC);
      int stage = 0;
      uint32_t phase = 0;
      const uint64_t keep = l2_policy_evict_last();
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = s_ring + stage * STAGE_BYTES;
          mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
          tma_load_2d_hint(&tmH, sa, &full_bar[stage], kb * BK, tile * BM, keep);
          tma_load_2d(&tmWd, sa + STAGE_A, &full_bar[stage], kb * BK, 0);
          if (++stage == p.nstage) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ============================== MMA issuer ==============================
    if (lane == 0) {
      const uint32_t idesc_down = umma_idesc_bf16(BM, RP);
      const uint32_t idesc_up = umma_idesc_bf16(BM, static_cast<uint32_t>(p.UC));
      int stage = 0;
      uint32_t phase = 0;
      uint32_t cc = 0;   // running U-chunk counter: TMEM stage = cc & 1
      uint32_t it = 0;
      // One k-block of S1 = h · W_dᵀ.  S1 is free whenever this runs: the epilogue drains it before it signals s_ready of
      // the tile before, and this thread has waited for that signal.
      auto down_kb = [&](int kb) {
        tc_fence_after();
        const uint32_t sa = smem_u32(s_ring + stage * STAGE_BYTES);
        const uint64_t adesc = umma_desc_k_sw128(sa), bdesc = umma_desc_k_sw128(sa + STAGE_A);
#pragma unroll
        for (int k = 0; k < BK / 16; ++k)
          umma_bf16_ss(tmem_base + S1_COL, adesc + static_cast<uint64_t>(k * 2), bdesc + static_cast<uint64_t>(k * 2), idesc_down,
                       (kb | k) != 0 ? 1u : 0u);
        umma_commit(&empty_bar[stage]);
        if (++stage == p.nstage) {
          stage = 0;
          phase ^= 1;
        }
        if (kb == nkb - 1) umma_commit(s1_full);
      };
      mbar_wait(wu_bar, 0);
      if (static_cast<int>(blockIdx.x) < num_tiles) {
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          down_kb(kb);
        }
      }
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const bool has_next = tile + static_cast<int>(gridDim.x) < num_tiles;
        mbar_wait(s_ready, it & 1);
        tc_fence_after();
        const uint64_t adesc = umma_desc_k_sw128(smem_u32(s_act));
        // Two independent streams share this thread: the up-projection chunks of this tile (gated by the epilogue freeing
        // a TMEM stage) and the down-projection k-blocks of the NEXT tile (gated by TMA), so that the next tile's HBM
        // stream overlaps this tile's epilogue.  Poll both; the up-projection has priority (the epilogue waits on it).
        int c = 0, kb = has_next ? 0 : nkb;
        while (c < nch || kb < nkb) {
          if (c < nch) {
            const uint32_t st = cc & 1;
            if (mbar_try_wait(&u_empty[st], ((cc >> 1) & 1) ^ 1)) {
              tc_fence_after();
              const uint64_t bdesc = umma_desc_k_sw128(smem_u32(s_wu + static_cast<size_t>(c) * p.UC * 128));
#pragma unroll
              for (int k = 0; k < RP / 16; ++k)
                umma_bf16_ss(tmem_base + st * p.UC, adesc + static_cast<uint64_t>(k * 2), bdesc + static_cast<uint64_t>(k * 2),
                             idesc_up, k != 0 ? 1u : 0u);
              umma_commit(&u_full[st]);
              ++c;
              ++cc;
              continue;
            }
          }
          if (kb < nkb && mbar_try_wait(&full_bar[stage], phase)) {
            down_kb(kb);
            ++kb;
          }
        }
      }
    }
  } else if (warp >= 4) {
    // ============================== epilogue ==============================
    const int ew = warp - 4;                              // 0..15
    const int quad = warp & 3, cs = ew >> 2;              // TMEM read-out: lane quadrant, 16-column slice of the chunk
    const int rl = quad * 32 + lane;                      // row within the tile = TMEM lane
    const uint32_t lane_addr = static_cast<uint32_t>(quad * 32) << 16;
    const int half = lane >> 4, s16 = lane & 15;          // row-distributed layout: 16 lanes x 4 columns per row
    const int rr0 = ew * 2 + half;                        // this thread's row in pass 0 (pass ps: rr0 + 32 ps)
    const int n64 = p.H / CW;
    const int sub_per_u = p.UC / CW;
    const uint32_t ust = smem_u32(s_ust);
    const uint32_t sact = smem_u32(s_act);
    // staging tile addresses: writer (row rl, 16-byte chunks cs*4 .. cs*4+3), reader (row rr0 + 32 ps, chunk s16)
    const uint32_t wr_base = ust + rl * (CW * 4) + ((cs * 4) & 8) * 16;
    const uint32_t rd_base = ust + rr0 * (CW * 4) + (((s16 & 8) | ((s16 ^ rr0) & 7)) << 4);   // (rr0 + 32 ps) & 7 == rr0 & 7
    __nv_bfloat16* zbuf = (p.tail == 0 && p.z_out != nullptr) ? p.z_out : p.out;
    const bool has_in = p.input != nullptr;
    const int step_h = PASS_ROWS * static_cast<int>(p.ldh), step_i = PASS_ROWS * static_cast<int>(p.ldi);
    const int step_z = PASS_ROWS * p.H;
    const uint64_t keep = l2_policy_evict_last(), drop = l2_policy_evict_first();
    uint32_t cu = 0, it = 0;   // cu: running U-chunk counter (TMEM stage = cu & 1)
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int row0 = tile * BM;
      // rows past M (last tile only): loads and stores are masked
      uint32_t vm = 0;
#pragma unroll
      for (int ps = 0; ps < PASSES; ++ps) vm |= (row0 + rr0 + ps * PASS_ROWS < p.M ? 1u : 0u) << ps;
      const __nv_bfloat16* hp = p.h + static_cast<int64_t>(row0 + rr0) * p.ldh + 4 * s16;
      const __nv_bfloat16* ip = has_in ? p.input + static_cast<int64_t>(row0 + rr0) * p.ldi + 4 * s16 : hp;
      __nv_bfloat16* zp = zbuf + static_cast<int64_t>(row0 + rr0) * p.H + 4 * s16;
      __nv_bfloat16* op = p.out + static_cast<int64_t>(row0 + rr0) * p.H + 4 * s16;
      uint2 hv0[PASSES], iv0[PASSES], hv1[PASSES], iv1[PASSES];
      auto load_res = [&](int c, uint2 (&hv)[PASSES], uint2 (&iv)[PASSES]) {
#pragma unroll
        for (int ps = 0; ps < PASSES; ++ps) {
          const bool ok = (vm >> ps) & 1u;
          hv[ps] = ok ? ld_nc_v2(hp + ps * step_h + c * CW, drop) : make_uint2(0u, 0u);
          iv[ps] = (ok && has_in) ? ld_nc_v2(ip + ps * step_i + c * CW, drop) : make_uint2(0u, 0u);
        }
      };
      load_res(0, hv0, iv0);   // residuals of chunk 0: in flight while phase 1 runs
      // ---- phase 1: s = act(S1 + b_d) -> bf16 operand tile (columns [16 cs, 16 cs + 16)) ----
      mbar_wait(s1_full, it & 1);
      tc_fence_after();
      {
        uint32_t acc[16];
        tmem_ld_32x32b_x16(tmem_base + lane_addr + S1_COL + cs * 16, acc);
        tmem_ld_wait();
        uint32_t w[8], wu[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float pre[2], v[2];
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int c = cs * 16 + 2 * i + e;
            pre[e] = c < p.r ? __uint_as_float(acc[2 * i + e]) + __ldg(p.b_down + c) : 0.0f;
            v[e] = c < p.r ? (p.act == 1 ? gelu_fast(pre[e]) : fmaxf(pre[e], 0.0f)) : 0.0f;
          }
          w[i] = pack_bf16x2(v[0], v[1]);
          wu[i] = pack_bf16x2(pre[0], pre[1]);
        }
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const int c16 = cs * 2 + q;
          sts_v4(sact + rl * 128 + ((c16 ^ (rl & 7)) << 4), w[4 * q], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]);
        }
        if (p.u_out != nullptr && row0 + rl < p.M) {
          // GELU only: the pre-activation feeds GELU' in the backward (2r bytes per token; row-per-thread stores)
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            const int c = cs * 16 + q * 8;
            if (c < p.r)
              st_na_v4(p.u_out + static_cast<int64_t>(row0 + rl) * p.r + c,
                       make_uint4(wu[4 * q], wu[4 * q + 1], wu[4 * q + 2], wu[4 * q + 3]));
          }
        }
      }
      fence_proxy_async_smem();   // generic-proxy writes of s_act -> visible to the tensor core's async-proxy reads
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(s_ready);

      // ---- phase 2: z = U + b_u + h + input, bf16; row statistics ----
      float sum[PASSES];
      float2 sq[PASSES];
#pragma unroll
      for (int ps = 0; ps < PASSES; ++ps) {
        sum[ps] = 0.0f;
        sq[ps] = make_float2(0.0f, 0.0f);
      }

      int sub = 0;               // 64-column sub-chunk inside the current U chunk
      auto chunk = [&](int c, const uint2 (&hv)[PASSES], const uint2 (&iv)[PASSES], uint2 (&hn)[PASSES], uint2 (&in)[PASSES]) {
        const uint32_t st = cu & 1;
        if (sub == 0) {
          mbar_wait(&u_full[st], (cu >> 1) & 1);
          tc_fence_after();
        }
        const uint32_t boff = static_cast<uint32_t>(c & 1) * UST_BYTES;
        {
          uint32_t acc[16];
          tmem_ld_32x32b_x16(tmem_base + lane_addr + st * p.UC + sub * CW + cs * 16, acc);
          tmem_ld_wait();
#pragma unroll
          for (int jj = 0; jj < 4; ++jj)
            sts_v4(wr_base + boff + ((((cs * 4 + jj) ^ rl) & 7) << 4), acc[4 * jj], acc[4 * jj + 1], acc[4 * jj + 2], acc[4 * jj + 3]);
        }
        if (++sub == sub_per_u) {
          sub = 0;
          ++cu;
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&u_empty[st]);
        }
        if (c + 1 < n64) load_res(c + 1, hn, in);   // after the read-out: its 16 accumulator registers are free again
        named_bar_sync(1, EPI_THREADS);
        if (c == 0 && p.s_out != nullptr) {
          // s leaves through the operand tile: 8 lanes per 128-byte row, coalesced
          for (int i = threadIdx.x - 128; i < BM * 8; i += EPI_THREADS) {
            const int r_ = i >> 3, ch = i & 7;
            if (row0 + r_ < p.M && ch * 8 < p.r)
              st_na_v4(p.s_out + static_cast<int64_t>(row0 + r_) * p.lds + ch * 8, lds_v4(sact + r_ * 128 + ((ch ^ (r_ & 7)) << 4)));
            if (row0 + r_ < p.M && ch == 0 && p.lds >= p.r + 8)   // the ones column of [s | 1] (bf16 1.0 = 0x3F80)
              st_na_v4(p.s_out + static_cast<int64_t>(row0 + r_) * p.lds + p.r, make_uint4(0x00003F80u, 0u, 0u, 0u));
          }
        }
        const float4 bu = __ldg(reinterpret_cast<const float4*>(p.b_up + c * CW + 4 * s16));
        const float2 bu0 = make_float2(bu.x, bu.y), bu1 = make_float2(bu.z, bu.w);
#pragma unroll
        for (int ps = 0; ps < PASSES; ++ps) {
          const float4 u = lds_v4f(rd_base + boff + ps * (PASS_ROWS * CW * 4));
          float2 v0 = __fadd2_rn(make_float2(u.x, u.y), bu0), v1 = __fadd2_rn(make_float2(u.z, u.w), bu1);
          v0 = __fadd2_rn(v0, bf16x2_to_f2(hv[ps].x));
          v1 = __fadd2_rn(v1, bf16x2_to_f2(hv[ps].y));
          if (has_in) {
            v0 = __fadd2_rn(v0, bf16x2_to_f2(iv[ps].x));
            v1 = __fadd2_rn(v1, bf16x2_to_f2(iv[ps].y));
          }
          const uint32_t w0 = pack_bf16x2(v0.x, v0.y), w1 = pack_bf16x2(v1.x, v1.y);
          // statistics of the ROUNDED row: LayerNorm's input is the bf16 tensor the backward will read
          const float2 f0 = bf16x2_to_f2(w0), f1 = bf16x2_to_f2(w1);
          const float2 fs = __fadd2_rn(f0, f1);
          sum[ps] += fs.x + fs.y;
          sq[ps] = __ffma2_rn(f0, f0, __ffma2_rn(f1, f1, sq[ps]));
          if ((vm >> ps) & 1u) {
            __nv_bfloat16* dst = zp + ps * step_z + c * CW;
            st_v2(dst, w0, w1, p.tail == 0 ? keep : drop);
          }
        }
      };
      for (int c = 0; c < n64; c += 2) {
        chunk(c, hv0, iv0, hv1, iv1);
        if (c + 1 < n64) chunk(c + 1, hv1, iv1, hv0, iv0);
      }
      if (p.tail != 0) continue;   // out = z: done (uniform across the epilogue warps)

      // ---- phase 3: LayerNorm; the 16 lanes of a row combine their partial statistics ----
      const float inv_h = 1.0f / static_cast<float>(p.H);
      float2 nmu[PASSES], rs2[PASSES];
#pragma unroll
      for (int ps = 0; ps < PASSES; ++ps) {
        float s_ = sum[ps], q_ = sq[ps].x + sq[ps].y;
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) {
          s_ += __shfl_xor_sync(0xffffffffu, s_, o);
          q_ += __shfl_xor_sync(0xffffffffu, q_, o);
        }
        const float mu = s_ * inv_h;
        const float var = fmaxf(q_ * inv_h - mu * mu, 0.0f);
        const float rs = rsqrtf(var + p.eps);
        nmu[ps] = make_float2(-mu, -mu);
        rs2[ps] = make_float2(rs, rs);
        if (s16 == 0 && ((vm >> ps) & 1u)) {
          const int row = row0 + rr0 + ps * PASS_ROWS;
          if (p.mean != nullptr) p.mean[row] = mu;
          if (p.rstd != nullptr) p.rstd[row] = rs;
        }
      }
      // z is re-read by the thread that wrote it (rows past M are not read: their slots were never written)
      auto load_z = [&](int c, uint2 (&zv)[PASSES]) {
#pragma unroll
        for (int ps = 0; ps < PASSES; ++ps)
          zv[ps] = ((vm >> ps) & 1u) ? ld_cg_v2(zp + ps * step_z + c * CW, drop) : make_uint2(0u, 0u);
      };
      auto norm_chunk = [&](int c, const uint2 (&zv)[PASSES], uint2 (&zn)[PASSES]) {
        if (c + 1 < n64) load_z(c + 1, zn);
        const float4 g = __ldg(reinterpret_cast<const float4*>(p.gamma + c * CW + 4 * s16));
        const float4 b = __ldg(reinterpret_cast<const float4*>(p.beta + c * CW + 4 * s16));
        const float2 g0 = make_float2(g.x, g.y), g1 = make_float2(g.z, g.w), b0 = make_float2(b.x, b.y), b1 = make_float2(b.z, b.w);
#pragma unroll
        for (int ps = 0; ps < PASSES; ++ps) {
          const float2 z0 = __fmul2_rn(__fadd2_rn(bf16x2_to_f2(zv[ps].x), nmu[ps]), rs2[ps]);
          const float2 z1 = __fmul2_rn(__fadd2_rn(bf16x2_to_f2(zv[ps].y), nmu[ps]), rs2[ps]);
          const float2 o0 = __ffma2_rn(z0, g0, b0), o1 = __ffma2_rn(z1, g1, b1);
          if ((vm >> ps) & 1u) st_v2(op + ps * step_z + c * CW, pack_bf16x2(o0.x, o0.y), pack_bf16x2(o1.x, o1.y), drop);
        }
      };
      load_z(0, hv0);
      for (int c = 0; c < n64; c += 2) {
        norm_chunk(c, hv0, hv1);
        if (c + 1 < n64) norm_chunk(c + 1, hv1, hv0);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 2) tmem_dealloc(tmem_base, TMEM_COLS);
}

constexpr size_t SMEM_LIMIT = 232448;   // 227 KB opt-in maximum per CTA on sm_100

size_t adapter_smem_fixed(int64_t H) {
  return static_cast<size_t>(H) * 128 + S_TILE + 2 * UST_BYTES + 24 * sizeof(uint64_t) + 16 + 1024;
}

int adapter_stages(int64_t H) {
  const size_t fixed = adapter_smem_fixed(H);
  int n = static_cast<int>((SMEM_LIMIT - fixed) / STAGE_BYTES);
  return n > MAX_STAGE ? MAX_STAGE : n;
}

}  // namespace

// adapter_rows_sm100.cu: the row-per-thread / TMA formulation of the same block
bool a4r_adapter_rows_supported(int64_t H, int64_t r);
int a4r_adapter_rows_launch(const a4r_adapter_args* a, cudaStream_t stream);

extern "C" int a4r_adapter_ln_supported(int64_t H, int64_t r) {
  return (H % 64 == 0 && H >= 64 && H <= 768 && r % 8 == 0 && r >= 8 && r <= 64) ? 1 : 0;
}

extern "C" int a4r_adapter_ln_fwd(const a4r_adapter_args* a, a4r_stream_t stream_) {
  A4R_CHECK_ARG(a != nullptr, "adapter_ln: args is NULL");
  A4R_CHECK_ARG(a->h && a->w_down && a->w_up && a->b_down && a->b_up && a->out, "adapter_ln: NULL pointer");
  A4R_CHECK_ARG(a->M >= 0 && a->M < (1ll << 31), "adapter_ln: bad M");
  A4R_CHECK_ARG(a4r_adapter_ln_supported(a->H, a->r), "adapter_ln: needs H %% 64 == 0, H <= 768, r %% 8 == 0, r <= 64 (H=%lld r=%lld)",
                (long long)a->H, (long long)a->r);
  A4R_CHECK_ARG(a->tail >= 0 && a->tail <= 2 && (a->act == 0 || a->act == 1), "adapter_ln: bad tail/act");
  A4R_CHECK_ARG(a->lds == 0 || (a->lds >= a->r && a->lds % 8 == 0 && a->lds < (1 << 20)), "adapter_ln: bad lds");
  A4R_CHECK_ARG(a->tail != 0 || (a->gamma && a->beta), "adapter_ln: tail 0 (LayerNorm) needs gamma and beta");
  A4R_CHECK_ARG(a->tail == 2 || a->input != nullptr, "adapter_ln: tails 0 and 1 add `input`");
  A4R_CHECK_ARG(a->ldh >= a->H && a->ldh % 8 == 0 && a4r_aligned16(a->h), "adapter_ln: bad h/ldh");
  A4R_CHECK_ARG(a->input == nullptr || (a->ldi >= a->H && a->ldi % 8 == 0 && a4r_aligned16(a->input)), "adapter_ln: bad input/ldi");
  A4R_CHECK_ARG(a4r_aligned16(a->w_down) && a4r_aligned16(a->w_up) && a4r_aligned16(a->out) && a4r_aligned16(a->b_up) &&
                    (a->gamma == nullptr || (a4r_aligned16(a->gamma) && a4r_aligned16(a->beta))),
                "adapter_ln: pointers must be 16B aligned");
  A4R_CHECK_ARG((a->z_out == nullptr || a4r_aligned16(a->z_out)) && (a->s_out == nullptr || a4r_aligned16(a->s_out)) &&
                    (a->u_out == nullptr || a4r_aligned16(a->u_out)),
                "adapter_ln: optional outputs must be 16B aligned");
  int rc = a4r_device_check();
  if (rc != A4R_OK) return rc;
  if (a->M == 0) return A4R_OK;
  A4R_CHECK_ARG(a->impl == 0 || a->impl == 2 || a->impl == 3, "adapter_ln: impl must be 0 (default), 2 (staged) or 3 (rows)");
  A4R_CHECK_ARG(a->impl != 3 || a4r_adapter_rows_supported(a->H, a->r), "adapter_ln: impl 3 does not support H=%lld r=%lld",
                (long long)a->H, (long long)a->r);
  if (a->impl != 2 && a4r_adapter_rows_supported(a->H, a->r)) return a4r_adapter_rows_launch(a, static_cast<cudaStream_t>(stream_));

  AdParams p;
  p.h = static_cast<const __nv_bfloat16*>(a->h);
  p.input = a->tail == 2 ? nullptr : static_cast<const __nv_bfloat16*>(a->input);
  p.ldh = a->ldh;
  p.ldi = a->ldi;
  p.b_down = a->b_down;
  p.b_up = a->b_up;
  p.gamma = a->gamma;
  p.beta = a->beta;
  p.out = static_cast<__nv_bfloat16*>(a->out);
  p.z_out = static_cast<__nv_bfloat16*>(a->z_out);
  p.mean = a->mean;
  p.rstd = a->rstd;
  p.s_out = static_cast<__nv_bfloat16*>(a->s_out);
  p.u_out = static_cast<__nv_bfloat16*>(a->u_out);
  p.M = static_cast<int>(a->M);
  p.H = static_cast<int>(a->H);
  p.r = static_cast<int>(a->r);
  p.lds = a->lds == 0 ? p.r : static_cast<int>(a->lds);
  p.UC = a->H % 192 == 0 ? 192 : (a->H % 128 == 0 ? 128 : 64);
  p.act = a->act;
  p.tail = a->tail;
  p.eps = a->eps;

  CUtensorMap tmH, tmWd, tmWu, tmI;
  if ((rc = a4r_make_tmap_bf16(&tmH, a->h, a->M, a->H, a->ldh, BM)) != A4R_OK) return rc;
  // W_d [r, H]: a 64-row box whose rows >= r are zero-filled; W_u [H, r]: 64-column box whose columns >= r are zero-filled
  if ((rc = a4r_make_tmap_bf16(&tmWd, a->w_down, a->r, a->H, a->H, RP)) != A4R_OK) return rc;
  if ((rc = a4r_make_tmap_bf16(&tmWu, a->w_up, a->H, a->r, a->r, p.UC)) != A4R_OK) return rc;
  if (p.input != nullptr) {
    if ((rc = a4r_make_tmap_bf16(&tmI, p.input, a->M, a->H, a->ldi, BM)) != A4R_OK) return rc;
  } else {
    tmI = tmH;
  }

  p.nstage = adapter_stages(a->H);
  A4R_CHECK_ARG(p.nstage >= 2, "adapter_ln: shared memory does not fit two pipeline stages at H=%lld", (long long)a->H);
  const size_t smem = adapter_smem_fixed(a->H) + static_cast<size_t>(p.nstage) * STAGE_BYTES;
  A4R_CUDA_OK(cudaFuncSetAttribute(adapter_ln_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  const int tiles = (p.M + BM - 1) / BM;
  const int grid = tiles < a4r_num_sms() ? tiles : a4r_num_sms();
  adapter_ln_kernel<<<grid, THREADS, smem, static_cast<cudaStream_t>(stream_)>>>(tmH, tmWd, tmWu, tmI, p);
  A4R_LAUNCH_OK();
  a4r_count_launch(1);
  return A4R_OK;
}
