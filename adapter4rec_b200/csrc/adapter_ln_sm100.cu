// adapter_ln_sm100.cu — K5: the Houlsby adapter block in ONE pass over the hidden states.
//
//   a   = h + W_u · act(W_d · h + b_d) + b_u            AdapterBlock.forward, Downstream/Text/model/modules.py:131-134
//   out = LayerNorm(a + input)                          BertAdaptedSelfOutput.forward, model.py:292-297          (tail 0)
//   out = a + input                                     VITAdaptedOutput.forward,  Downstream/CV/model/model.py  (tail 1)
//   out = a                                             VITAdaptedSelfOutput.forward                             (tail 2)
//
// Per 128-token tile, one persistent CTA per SM:
//   warp 0    TMA producer: streams h in [128 x 64] k-blocks (+ the matching [r x 64] block of W_d) through a 4-stage
//             mbarrier ring; W_u ([H x r], 96 KB at H = 768) is loaded once per CTA and stays resident.
//   warp 1    tcgen05 issuer: S1[128 x 64] = h · W_dᵀ (TMEM), then — once the epilogue has written act(S1 + b_d) as a
//             bf16 SWIZZLE_128B tile to shared memory — U = s · W_uᵀ in column chunks of UC <= 192 through two TMEM stages.
//   warps 4-11 epilogue: each thread owns one token row (a TMEM lane).  Phase 1 turns S1 into the A operand of the second
//             MMA; phase 2 forms z = U + b_u + h + input chunk by chunk, rounds it to bf16, writes it out and accumulates
//             the row sum / sum of squares; after the two column groups of a row have exchanged statistics, phase 3
//             re-reads the row (its own writes: L2-resident) and writes LN(z).
// HBM traffic per token (H = 768): read h + input, write out = 4,608 B (+ z = 1,536 B and s = 2r B when the backward
// needs them) against ~9.5 KB for the composition GEMM -> GEMM(+2 residuals) -> LayerNorm; the contraction FLOPs
// (2 x 2 x 768 x 64 per token) are ~5 % of what the tensor pipe could do in the time HBM needs, so the bound is HBM.
#include "a4r_common.cuh"

namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int RP = 64;                       // adapter rank padded to one k-block (r <= 64)
constexpr int NSTAGE = 4;
constexpr int STAGE_A = BM * BK * 2;         // 16 KB of h
constexpr int STAGE_B = RP * BK * 2;         // 8 KB of W_d
constexpr int STAGE_BYTES = STAGE_A + STAGE_B;
constexpr int S_TILE = BM * RP * 2;          // 16 KB
constexpr int EPI_WARPS = 8;
constexpr int THREADS = 128 + EPI_WARPS * 32;
constexpr int TMEM_COLS = 512;
constexpr int S1_COL = 448;                  // S1 lives in columns [448, 512); U stages at 0 and UC (2 * UC <= 384)

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
  int M, H, r, UC;
  int act, tail;
  float eps;
};

A4R_DEVICE void named_bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }

A4R_DEVICE uint4 ld_v4(const void* p) {  // coherent load: phase 3 re-reads what this thread wrote in phase 2
  uint4 r;
  asm volatile("ld.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
  return r;
}

__global__ void __launch_bounds__(THREADS, 1)
adapter_ln_kernel(const __grid_constant__ CUtensorMap tmH, const __grid_constant__ CUtensorMap tmWd,
                  const __grid_constant__ CUtensorMap tmWu, const AdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* s_wu = smem;                                    // [H][64] bf16, SW128
  uint8_t* s_ring = s_wu + static_cast<size_t>(p.H) * 128;
  uint8_t* s_act = s_ring + NSTAGE * STAGE_BYTES;          // [128][64] bf16, SW128: A operand of the up-projection
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(s_act + S_TILE);
  uint64_t* empty_bar = full_bar + NSTAGE;
  uint64_t* wu_bar = empty_bar + NSTAGE;
  uint64_t* s1_full = wu_bar + 1;
  uint64_t* s_ready = s1_full + 1;
  uint64_t* u_full = s_ready + 1;    // [2]
  uint64_t* u_empty = u_full + 2;    // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(u_empty + 2);
  float2* row_stat = reinterpret_cast<float2*>(tmem_slot + 2);   // [2 groups][128 rows]: (sum, sum of squares)

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
    for (int s = 0; s < NSTAGE; ++s) {
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
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = s_ring + stage * STAGE_BYTES;
          mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
          tma_load_2d(&tmH, sa, &full_bar[stage], kb * BK, tile * BM);
          tma_load_2d(&tmWd, sa + STAGE_A, &full_bar[stage], kb * BK, 0);
          if (++stage == NSTAGE) {
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
      mbar_wait(wu_bar, 0);
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        // S1 is free: the epilogue drained it before signalling s_ready of the previous tile, which this thread waited for
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(s_ring + stage * STAGE_BYTES);
          const uint64_t adesc = umma_desc_k_sw128(sa), bdesc = umma_desc_k_sw128(sa + STAGE_A);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            umma_bf16_ss(tmem_base + S1_COL, adesc + static_cast<uint64_t>(k * 2), bdesc + static_cast<uint64_t>(k * 2), idesc_down,
                         (kb | k) != 0 ? 1u : 0u);
          umma_commit(&empty_bar[stage]);
          if (++stage == NSTAGE) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(s1_full);
        mbar_wait(s_ready, it & 1);
        tc_fence_after();
        const uint64_t adesc = umma_desc_k_sw128(smem_u32(s_act));
        for (int c = 0; c < nch; ++c, ++cc) {
          const uint32_t st = cc & 1;
          mbar_wait(&u_empty[st], ((cc >> 1) & 1) ^ 1);
          tc_fence_after();
          const uint64_t bdesc = umma_desc_k_sw128(smem_u32(s_wu + static_cast<size_t>(c) * p.UC * 128));
#pragma unroll
          for (int k = 0; k < RP / 16; ++k)
            umma_bf16_ss(tmem_base + st * p.UC, adesc + static_cast<uint64_t>(k * 2), bdesc + static_cast<uint64_t>(k * 2), idesc_up,
                         k != 0 ? 1u : 0u);
          umma_commit(&u_full[st]);
        }
      }
    }
  } else if (warp >= 4) {
    // ============================== epilogue ==============================
    const int quad = warp & 3, group = (warp - 4) >> 2;
    const int rl = quad * 32 + lane;                      // row within the tile = TMEM lane
    const uint32_t lane_addr = static_cast<uint32_t>(quad * 32) << 16;
    const int nsub = p.UC / 32;
    uint32_t cc = 0, it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int row = tile * BM + rl;
      const bool row_ok = row < p.M;
      const int64_t r64 = row;
      // ---- phase 1: s = act(S1 + b_d) -> bf16 operand tile (columns [32 group, 32 group + 32)) ----
      mbar_wait(s1_full, it & 1);
      tc_fence_after();
      {
        uint32_t acc[32];
        tmem_ld_32x32b_x32(tmem_base + lane_addr + S1_COL + group * 32, acc);
        tmem_ld_wait();
        float pre[32], v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const int c = group * 32 + i;
          pre[i] = c < p.r ? __uint_as_float(acc[i]) + __ldg(p.b_down + c) : 0.0f;
          v[i] = c < p.r ? (p.act == 1 ? gelu_fast(pre[i]) : fmaxf(pre[i], 0.0f)) : 0.0f;
        }
        uint32_t w[16], wu[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          w[i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);
          wu[i] = pack_bf16x2(pre[2 * i], pre[2 * i + 1]);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int c16 = group * 4 + q;
          *reinterpret_cast<uint4*>(s_act + rl * 128 + ((c16 ^ (rl & 7)) << 4)) =
              make_uint4(w[4 * q], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]);
        }
        if (row_ok) {
          // training: the activation output (and, for GELU, its pre-activation) feed the backward; rows are r wide
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int c = group * 32 + q * 8;
            if (c < p.r) {
              if (p.s_out != nullptr)
                st_na_v4(p.s_out + r64 * p.r + c, make_uint4(w[4 * q], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]));
              if (p.u_out != nullptr)
                st_na_v4(p.u_out + r64 * p.r + c, make_uint4(wu[4 * q], wu[4 * q + 1], wu[4 * q + 2], wu[4 * q + 3]));
            }
          }
        }
      }
      fence_proxy_async_smem();   // generic-proxy writes of s_act -> visible to the tensor core's async-proxy reads
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(s_ready);

      // ---- phase 2: z = U + b_u + h + input, bf16; row statistics ----
      __nv_bfloat16* zbuf = (p.tail == 0 && p.z_out != nullptr) ? p.z_out : p.out;
      float sum = 0.0f, sq = 0.0f;
      for (int c = 0; c < nch; ++c, ++cc) {
        const uint32_t st = cc & 1;
        mbar_wait(&u_full[st], (cc >> 1) & 1);
        tc_fence_after();
        for (int sc = group; sc < nsub; sc += 2) {
          const int col0 = c * p.UC + sc * 32;
          uint4 hv[4], iv[4];
          if (row_ok) {
#pragma unroll
            for (int q = 0; q < 4; ++q) hv[q] = ld_nc_v4(p.h + r64 * p.ldh + col0 + 8 * q);
            if (p.input != nullptr) {
#pragma unroll
              for (int q = 0; q < 4; ++q) iv[q] = ld_nc_v4(p.input + r64 * p.ldi + col0 + 8 * q);
            }
          }
          uint32_t acc[32];
          tmem_ld_32x32b_x32(tmem_base + lane_addr + st * p.UC + sc * 32, acc);
          float4 bv[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) bv[j] = __ldg(reinterpret_cast<const float4*>(p.b_up + col0 + 4 * j));
          tmem_ld_wait();
          if (!row_ok) continue;
          float v[32];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            v[4 * j] = __uint_as_float(acc[4 * j]) + bv[j].x;
            v[4 * j + 1] = __uint_as_float(acc[4 * j + 1]) + bv[j].y;
            v[4 * j + 2] = __uint_as_float(acc[4 * j + 2]) + bv[j].z;
            v[4 * j + 3] = __uint_as_float(acc[4 * j + 3]) + bv[j].w;
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint32_t hw[4] = {hv[q].x, hv[q].y, hv[q].z, hv[q].w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 f = unpack_bf16x2(hw[e]);
              v[8 * q + 2 * e] += f.x;
              v[8 * q + 2 * e + 1] += f.y;
            }
          }
          if (p.input != nullptr) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const uint32_t iw[4] = {iv[q].x, iv[q].y, iv[q].z, iv[q].w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float2 f = unpack_bf16x2(iw[e]);
                v[8 * q + 2 * e] += f.x;
                v[8 * q + 2 * e + 1] += f.y;
              }
            }
          }
          uint32_t w[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            w[i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);
            // statistics of the ROUNDED row: LayerNorm's input is the bf16 tensor the backward will read
            const float2 f = unpack_bf16x2(w[i]);
            sum += f.x + f.y;
            sq = fmaf(f.x, f.x, fmaf(f.y, f.y, sq));
          }
          __nv_bfloat16* dst = zbuf + r64 * p.H + col0;
#pragma unroll
          for (int q = 0; q < 4; ++q)
            *reinterpret_cast<uint4*>(dst + 8 * q) = make_uint4(w[4 * q], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&u_empty[st]);
      }
      if (p.tail != 0) continue;   // out = z: done (uniform across the epilogue warps)

      // ---- phase 3: LayerNorm over the full row ----
      row_stat[group * BM + rl] = make_float2(sum, sq);
      named_bar_sync(1, EPI_WARPS * 32);
      const float2 other = row_stat[(group ^ 1) * BM + rl];
      named_bar_sync(1, EPI_WARPS * 32);   // row_stat may be overwritten by the next tile after this point
      const float inv_h = 1.0f / static_cast<float>(p.H);
      const float mu = (sum + other.x) * inv_h;
      const float var = fmaxf((sq + other.y) * inv_h - mu * mu, 0.0f);
      const float rs = rsqrtf(var + p.eps);
      if (!row_ok) continue;
      if (group == 0) {
        if (p.mean != nullptr) p.mean[row] = mu;
        if (p.rstd != nullptr) p.rstd[row] = rs;
      }
      for (int c = 0; c < nch; ++c) {
        for (int sc = group; sc < nsub; sc += 2) {
          const int col0 = c * p.UC + sc * 32;
          const __nv_bfloat16* src = zbuf + r64 * p.H + col0;
          uint4 zv[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) zv[q] = ld_v4(src + 8 * q);
          __nv_bfloat16* dst = p.out + r64 * p.H + col0;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 g0 = __ldg(reinterpret_cast<const float4*>(p.gamma + col0 + 8 * q));
            const float4 g1 = __ldg(reinterpret_cast<const float4*>(p.gamma + col0 + 8 * q + 4));
            const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.beta + col0 + 8 * q));
            const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.beta + col0 + 8 * q + 4));
            const float2 z0 = unpack_bf16x2(zv[q].x), z1 = unpack_bf16x2(zv[q].y), z2 = unpack_bf16x2(zv[q].z),
                         z3 = unpack_bf16x2(zv[q].w);
            uint4 o;
            o.x = pack_bf16x2(fmaf((z0.x - mu) * rs, g0.x, b0.x), fmaf((z0.y - mu) * rs, g0.y, b0.y));
            o.y = pack_bf16x2(fmaf((z1.x - mu) * rs, g0.z, b0.z), fmaf((z1.y - mu) * rs, g0.w, b0.w));
            o.z = pack_bf16x2(fmaf((z2.x - mu) * rs, g1.x, b1.x), fmaf((z2.y - mu) * rs, g1.y, b1.y));
            o.w = pack_bf16x2(fmaf((z3.x - mu) * rs, g1.z, b1.z), fmaf((z3.y - mu) * rs, g1.w, b1.w));
            st_na_v4(dst + 8 * q, o);
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 2) tmem_dealloc(tmem_base, TMEM_COLS);
}

size_t adapter_smem_bytes(int64_t H) {
  return static_cast<size_t>(H) * 128 + NSTAGE * STAGE_BYTES + S_TILE + 16 * sizeof(uint64_t) + 16 + 2 * BM * sizeof(float2) +
         1024;
}

}  // namespace

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
  p.UC = a->H % 192 == 0 ? 192 : (a->H % 128 == 0 ? 128 : 64);
  p.act = a->act;
  p.tail = a->tail;
  p.eps = a->eps;

  CUtensorMap tmH, tmWd, tmWu;
  if ((rc = a4r_make_tmap_bf16(&tmH, a->h, a->M, a->H, a->ldh, BM)) != A4R_OK) return rc;
  // W_d [r, H]: a 64-row box whose rows >= r are zero-filled; W_u [H, r]: 64-column box whose columns >= r are zero-filled
  if ((rc = a4r_make_tmap_bf16(&tmWd, a->w_down, a->r, a->H, a->H, RP)) != A4R_OK) return rc;
  if ((rc = a4r_make_tmap_bf16(&tmWu, a->w_up, a->H, a->r, a->r, p.UC)) != A4R_OK) return rc;

  const size_t smem = adapter_smem_bytes(a->H);
  A4R_CUDA_OK(cudaFuncSetAttribute(adapter_ln_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  const int tiles = (p.M + BM - 1) / BM;
  const int grid = tiles < a4r_num_sms() ? tiles : a4r_num_sms();
  adapter_ln_kernel<<<grid, THREADS, smem, static_cast<cudaStream_t>(stream_)>>>(tmH, tmWd, tmWu, p);
  A4R_LAUNCH_OK();
  a4r_count_launch(1);
  return A4R_OK;
}
