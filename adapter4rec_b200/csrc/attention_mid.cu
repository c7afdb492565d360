// attention_mid.cu — K3 for MASKED mid-length sequences (32 < L <= 256, head_dim 64) + the dispatch of a4r_attn_mid_fwd / _bwd.
// The unmasked case — ViT-B/16's L = 197 (+ n prompt tokens), i.e. every call the image tree makes — runs on tcgen05
// (attention_tc_sm100.cu, attention_tc_bwd_sm100.cu); this file keeps the mma.sync formulation for sequences with a key mask.
//
// One CTA (4 warps) owns one (image, head): Q, K, V (and dO in the backward) of the whole sequence live in swizzled
// shared memory (<= 32 KB each), so HBM sees q, k, v once and ctx once, exactly like the short-sequence kernel.
// Forward: each warp takes 16-query tiles and runs a flash-style online softmax over 64-key blocks on mma.sync.
// The forward also writes the log-sum-exp of every score row (4 B per token and head).  Backward:
//   prologue: delta_i = <dO_i, O_i> for every query row (O = the forward output), lse and delta to shared memory;
//   phase A, per 16-query tile: recompute S, P = exp(S − lse), dP = dO·Vᵀ, dS = P ⊙ (dP − delta) * scale, dQ = dS·K
//            (written straight to global);
//   phase B, per 16-key tile:   recompute Sᵀ = K·Qᵀ, Pᵀ = exp(Sᵀ*scale − lse), dPᵀ = V·dOᵀ, dSᵀ = Pᵀ ⊙ (dPᵀ − delta) * scale,
//            dK = dSᵀ·Q, dV = Pᵀ·dO.
// Attention is ~4 % of a ViT layer's FLOPs; the tensor-pipe budget of the layer is in gemm_sm100.cu.  There is no mask
// (ViT attends to every token, Downstream/CV/model/encoders.py:31-32 via transformers' ViTSelfAttention); an optional
// key-validity mask with the same additive semantics as the short kernel is supported for completeness.
#include "a4r_common.cuh"
#include "mma_sync.cuh"

namespace {

constexpr int DH = 64;
constexpr int WARPS = 8;
constexpr int LMAX = 256;

A4R_DEVICE uint32_t toff(int row, int chunk) { return static_cast<uint32_t>(row * 128 + (((chunk ^ row) & 7) << 4)); }

struct MidParams {
  const __nv_bfloat16* qkv;
  __nv_bfloat16* out;         // fwd: ctx; bwd: dqkv
  const __nv_bfloat16* dout;  // bwd
  const void* mask;
  float* lse;                 // [N*L, heads] f32: written by fwd, read by bwd
  const __nv_bfloat16* ctx;   // bwd: the forward output O [N*L, ld_out]
  int64_t ld_qkv, ld_out, mask_ld;
  int N, L, Lp, heads;  // Lp = L rounded up to 64
  int mask_dtype;
  float scale, mask_neg;
};

// global [L rows x 64] -> swizzled smem tile with 16-byte cp.async (no register staging: all four tiles of a backward
// item are in flight at once); rows >= L are zero-filled.  Completion: cp_async_wait_all() + __syncthreads().
A4R_DEVICE void load_rows(uint8_t* tile, const __nv_bfloat16* g, int64_t ld, int L, int Lp) {
  const uint32_t base = smem_u32(tile);
  for (int i = threadIdx.x; i < Lp * 8; i += WARPS * 32) {
    const int row = i >> 3, ch = i & 7;
    const bool ok = row < L;
    const __nv_bfloat16* src = g + (ok ? static_cast<int64_t>(row) * ld + ch * 8 : 0);
    const int sz = ok ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(base + toff(row, ch)), "l"(src), "r"(sz) : "memory");
  }
}
A4R_DEVICE void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// A fragment (16 x 16) of a row-major [row][k] tile
A4R_DEVICE void lda(uint32_t (&a)[4], uint32_t base, int m0, int k0, int lane) {
  const int mi = lane >> 3, r = lane & 7;
  ldsm_x4(a, base + toff(m0 + r + ((mi & 1) << 3), (k0 >> 3) + (mi >> 1)));
}
// B fragments for n-tiles n0, n0+8 from a tile stored [n][k]
A4R_DEVICE void ldb_nk(uint32_t (&b)[4], uint32_t base, int n0, int k0, int lane) {
  const int mi = lane >> 3, r = lane & 7;
  ldsm_x4(b, base + toff(n0 + r + ((mi >> 1) << 3), (k0 >> 3) + (mi & 1)));
}
// B fragments for n-tiles n0, n0+8 from a tile stored [k][n]
A4R_DEVICE void ldb_kn(uint32_t (&b)[4], uint32_t base, int n0, int k0, int lane) {
  const int mi = lane >> 3, r = lane & 7;
  ldsm_x4_t(b, base + toff(k0 + r + ((mi & 1) << 3), (n0 >> 3) + (mi >> 1)));
}
A4R_DEVICE void c2a(uint32_t (&a)[4], const float (&c0)[4], const float (&c1)[4]) {
  a[0] = pack_bf16x2(c0[0], c0[1]);
  a[1] = pack_bf16x2(c0[2], c0[3]);
  a[2] = pack_bf16x2(c1[0], c1[1]);
  a[3] = pack_bf16x2(c1[2], c1[3]);
}

A4R_DEVICE bool key_ok(const MidParams& p, int n, int j) {
  if (j >= p.L) return false;
  if (p.mask_dtype == 1) return reinterpret_cast<const int64_t*>(p.mask)[static_cast<int64_t>(n) * p.mask_ld + j] != 0;
  if (p.mask_dtype == 2) return reinterpret_cast<const float*>(p.mask)[static_cast<int64_t>(n) * p.mask_ld + j] != 0.0f;
  return true;
}

// scores of a 16-query tile against one 64-key block: s[nt][4], nt = 0..7 (8 keys each); rows g and g+8 of the tile.
// nv = number of 16-key groups of the block that hold at least one real key (L = 197: the last block has one of four);
// the others are skipped (their scores stay 0 and are masked to -inf / 0 by the callers).
A4R_DEVICE void score_block(float (&s)[8][4], const uint32_t (&qa)[4][4], uint32_t sK, int key0, int lane, int nv) {
#pragma unroll
  for (int nt = 0; nt < 8; ++nt)
#pragma unroll
    for (int e = 0; e < 4; ++e) s[nt][e] = 0.0f;
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
    for (int np = 0; np < 4; ++np) {
      if (np < nv) {
        uint32_t b[4];
        ldb_nk(b, sK, key0 + np * 16, ks * 16, lane);
        const uint32_t b0[2] = {b[0], b[1]}, b1[2] = {b[2], b[3]};
        mma_bf16_16816(s[np * 2], qa[ks], b0);
        mma_bf16_16816(s[np * 2 + 1], qa[ks], b1);
      }
    }
  }
}

// apply scale + additive mask; columns >= L get -inf.  kflags: bit j of word (key0/32 + j/32) = key valid
A4R_DEVICE void mask_block(float (&s)[8][4], const MidParams& p, const uint32_t* kvalid, int key0, int lane) {
  const int t = lane & 3;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt)
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int j = key0 + nt * 8 + 2 * t + (e & 1);
      float v = s[nt][e] * p.scale;
      const bool valid = (kvalid[j >> 5] >> (j & 31)) & 1u;
      if (!valid) v += p.mask_neg;
      if (j >= p.L) v = -INFINITY;
      s[nt][e] = v;
    }
}

__global__ void __launch_bounds__(WARPS * 32, 2) attn_mid_fwd_kernel(const MidParams p) {
  extern __shared__ __align__(128) uint8_t sm_mid[];
  __shared__ uint32_t kvalid[LMAX / 32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int tile_bytes = p.Lp * 128;
  uint8_t *tQ = sm_mid, *tK = sm_mid + tile_bytes, *tV = sm_mid + 2 * tile_bytes;
  const uint32_t sQ = smem_u32(tQ), sK = smem_u32(tK), sV = smem_u32(tV);
  const int64_t Hd = static_cast<int64_t>(p.heads) * DH;
  for (int w = blockIdx.x; w < p.N * p.heads; w += gridDim.x) {
    const int n = w / p.heads, h = w % p.heads;
    const __nv_bfloat16* q = p.qkv + static_cast<int64_t>(n) * p.L * p.ld_qkv + h * DH;
    __syncthreads();  // previous iteration's readers are done
    load_rows(tQ, q, p.ld_qkv, p.L, p.Lp);
    load_rows(tK, q + Hd, p.ld_qkv, p.L, p.Lp);
    load_rows(tV, q + 2 * Hd, p.ld_qkv, p.L, p.Lp);
    if (threadIdx.x < LMAX / 32) {
      uint32_t bits = 0;
      for (int j = 0; j < 32; ++j) bits |= (key_ok(p, n, threadIdx.x * 32 + j) ? 1u : 0u) << j;
      kvalid[threadIdx.x] = bits;
    }
    cp_async_wait_all();
    __syncthreads();
    for (int m0 = warp * 16; m0 < p.Lp && m0 < ((p.L + 15) & ~15); m0 += WARPS * 16) {
      uint32_t qa[4][4];
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) lda(qa[ks], sQ, m0, ks * 16, lane);
      float o[8][4];
#pragma unroll
      for (int nt = 0; nt < 8; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) o[nt][e] = 0.0f;
      float mx[2] = {-INFINITY, -INFINITY}, sum[2] = {0.0f, 0.0f};
      for (int key0 = 0; key0 < p.Lp; key0 += 64) {
        float s[8][4];
        const int nv = min(4, (p.L - key0 + 15) >> 4);
        score_block(s, qa, sK, key0, lane, nv);
        mask_block(s, p, kvalid, key0, lane);
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          float bm = -INFINITY;
#pragma unroll
          for (int nt = 0; nt < 8; ++nt) bm = fmaxf(bm, fmaxf(s[nt][hh * 2], s[nt][hh * 2 + 1]));
          bm = fmaxf(bm, __shfl_xor_sync(0xffffffffu, bm, 1));
          bm = fmaxf(bm, __shfl_xor_sync(0xffffffffu, bm, 2));
          const float nm = fmaxf(mx[hh], bm);
          const float corr = (mx[hh] == -INFINITY) ? 0.0f : __expf(mx[hh] - nm);
          mx[hh] = nm;
          float bs = 0.0f;
#pragma unroll
          for (int nt = 0; nt < 8; ++nt)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const float ex = (nm == -INFINITY || (nt >> 1) >= nv) ? 0.0f : __expf(s[nt][hh * 2 + e] - nm);
              s[nt][hh * 2 + e] = ex;
              bs += ex;
            }
          bs += __shfl_xor_sync(0xffffffffu, bs, 1);
          bs += __shfl_xor_sync(0xffffffffu, bs, 2);
          sum[hh] = sum[hh] * corr + bs;
#pragma unroll
          for (int nt = 0; nt < 8; ++nt) {
            o[nt][hh * 2] *= corr;
            o[nt][hh * 2 + 1] *= corr;
          }
        }
        // O += P_blk · V_blk   (k = 64 keys of this block)
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          if (ks < nv) {
            uint32_t a[4];
            c2a(a, s[2 * ks], s[2 * ks + 1]);
#pragma unroll
            for (int np = 0; np < 4; ++np) {
              uint32_t b[4];
              ldb_kn(b, sV, np * 16, key0 + ks * 16, lane);
              const uint32_t b0[2] = {b[0], b[1]}, b1[2] = {b[2], b[3]};
              mma_bf16_16816(o[np * 2], a, b0);
              mma_bf16_16816(o[np * 2 + 1], a, b1);
            }
          }
        }
      }
      const float inv0 = 1.0f / sum[0], inv1 = 1.0f / sum[1];
      __nv_bfloat16* op = p.out + (static_cast<int64_t>(n) * p.L) * p.ld_out + h * DH;
      const int r0 = m0 + g, r1 = m0 + g + 8;
      if (p.lse != nullptr && t == 0) {
        if (r0 < p.L) p.lse[(static_cast<int64_t>(n) * p.L + r0) * p.heads + h] = mx[0] + __logf(sum[0]);
        if (r1 < p.L) p.lse[(static_cast<int64_t>(n) * p.L + r1) * p.heads + h] = mx[1] + __logf(sum[1]);
      }
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const int col = nt * 8 + 2 * t;
        if (r0 < p.L) *reinterpret_cast<uint32_t*>(op + static_cast<int64_t>(r0) * p.ld_out + col) = pack_bf16x2(o[nt][0] * inv0, o[nt][1] * inv0);
        if (r1 < p.L) *reinterpret_cast<uint32_t*>(op + static_cast<int64_t>(r1) * p.ld_out + col) = pack_bf16x2(o[nt][2] * inv1, o[nt][3] * inv1);
      }
    }
  }
}

__global__ void __launch_bounds__(WARPS * 32) attn_mid_bwd_kernel(const MidParams p) {
  extern __shared__ __align__(128) uint8_t sm_mid[];
  __shared__ uint32_t kvalid[LMAX / 32];
  __shared__ float s_lse[LMAX], s_delta[LMAX];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int tile_bytes = p.Lp * 128;
  uint8_t *tQ = sm_mid, *tK = tQ + tile_bytes, *tV = tK + tile_bytes, *tdO = tV + tile_bytes;
  const uint32_t sQ = smem_u32(tQ), sK = smem_u32(tK), sV = smem_u32(tV), sdO = smem_u32(tdO);
  const int64_t Hd = static_cast<int64_t>(p.heads) * DH;
  const int Lq = (p.L + 15) & ~15;
  for (int w = blockIdx.x; w < p.N * p.heads; w += gridDim.x) {
    const int n = w / p.heads, h = w % p.heads;
    const __nv_bfloat16* q = p.qkv + static_cast<int64_t>(n) * p.L * p.ld_qkv + h * DH;
    __nv_bfloat16* dq = p.out + static_cast<int64_t>(n) * p.L * p.ld_qkv + h * DH;
    __syncthreads();
    load_rows(tQ, q, p.ld_qkv, p.L, p.Lp);
    load_rows(tK, q + Hd, p.ld_qkv, p.L, p.Lp);
    load_rows(tV, q + 2 * Hd, p.ld_qkv, p.L, p.Lp);
    load_rows(tdO, p.dout + static_cast<int64_t>(n) * p.L * p.ld_out + h * DH, p.ld_out, p.L, p.Lp);
    if (threadIdx.x < LMAX / 32) {
      uint32_t bits = 0;
      for (int j = 0; j < 32; ++j) bits |= (key_ok(p, n, threadIdx.x * 32 + j) ? 1u : 0u) << j;
      kvalid[threadIdx.x] = bits;
    }
    cp_async_wait_all();
    __syncthreads();
    // ---------------- prologue: lse (saved by the forward) and delta_i = <dO_i, O_i> into shared memory ----------------
    // 8 lanes per row, one 16-byte chunk each: every load of the prologue is independent and issued up front
    for (int r = threadIdx.x; r < p.Lp; r += WARPS * 32)
      s_lse[r] = r < p.L ? p.lse[(static_cast<int64_t>(n) * p.L + r) * p.heads + h] : 0.0f;
    for (int i = threadIdx.x; i < p.Lp * 8; i += WARPS * 32) {   // Lp * 8 is a multiple of 32: whole warps stay together
      const int r = i >> 3, ch = i & 7;
      float d = 0.0f;
      if (r < p.L) {
        const int64_t grow = static_cast<int64_t>(n) * p.L + r;
        const uint4 ov = ld_nc_v4(p.ctx + grow * p.ld_out + h * DH + ch * 8);
        const uint4 dv4 = *reinterpret_cast<const uint4*>(tdO + toff(r, ch));
        const uint32_t ow[4] = {ov.x, ov.y, ov.z, ov.w}, dw[4] = {dv4.x, dv4.y, dv4.z, dv4.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 a = unpack_bf16x2(ow[e]), b = unpack_bf16x2(dw[e]);
          d += a.x * b.x + a.y * b.y;
        }
      }
      d += __shfl_xor_sync(0xffffffffu, d, 1);
      d += __shfl_xor_sync(0xffffffffu, d, 2);
      d += __shfl_xor_sync(0xffffffffu, d, 4);
      if (ch == 0) s_delta[r] = d;
    }
    __syncthreads();
    // ---------------- phase A: per query tile -> dQ ----------------
    for (int m0 = warp * 16; m0 < Lq; m0 += WARPS * 16) {
      uint32_t qa[4][4], da[4][4];
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        lda(qa[ks], sQ, m0, ks * 16, lane);
        lda(da[ks], sdO, m0, ks * 16, lane);
      }
      const float lse[2] = {s_lse[m0 + g], s_lse[m0 + g + 8]};
      const float delta[2] = {s_delta[m0 + g], s_delta[m0 + g + 8]};
      float acc[8][4];
#pragma unroll
      for (int nt = 0; nt < 8; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[nt][e] = 0.0f;
      for (int key0 = 0; key0 < p.Lp; key0 += 64) {
        float s[8][4], dp[8][4];
        const int nv = min(4, (p.L - key0 + 15) >> 4);
        score_block(s, qa, sK, key0, lane, nv);
        mask_block(s, p, kvalid, key0, lane);
        score_block(dp, da, sV, key0, lane, nv);  // dP = dO·Vᵀ has the same operand structure
#pragma unroll
        for (int nt = 0; nt < 8; ++nt)
          if ((nt >> 1) < nv) {
#pragma unroll
            for (int e = 0; e < 4; ++e)
              s[nt][e] = __expf(s[nt][e] - lse[e >> 1]) * (dp[nt][e] - delta[e >> 1]) * p.scale;
          }
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          if (ks < nv) {
            uint32_t a[4];
            c2a(a, s[2 * ks], s[2 * ks + 1]);
#pragma unroll
            for (int np = 0; np < 4; ++np) {
              uint32_t b[4];
              ldb_kn(b, sK, np * 16, key0 + ks * 16, lane);
              const uint32_t b0[2] = {b[0], b[1]}, b1[2] = {b[2], b[3]};
              mma_bf16_16816(acc[np * 2], a, b0);
              mma_bf16_16816(acc[np * 2 + 1], a, b1);
            }
          }
        }
      }
      const int r0 = m0 + g, r1 = m0 + g + 8;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const int col = nt * 8 + 2 * t;
        if (r0 < p.L) *reinterpret_cast<uint32_t*>(dq + static_cast<int64_t>(r0) * p.ld_qkv + col) = pack_bf16x2(acc[nt][0], acc[nt][1]);
        if (r1 < p.L) *reinterpret_cast<uint32_t*>(dq + static_cast<int64_t>(r1) * p.ld_qkv + col) = pack_bf16x2(acc[nt][2], acc[nt][3]);
      }
    }
    // (phase B only reads shared tiles and s_lse / s_delta, all written before the barrier above)
    // ---------------- phase B: per key tile -> dK, dV ----------------
    for (int j0 = warp * 16; j0 < Lq; j0 += WARPS * 16) {
      uint32_t ka[4][4], va[4][4];
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        lda(ka[ks], sK, j0, ks * 16, lane);
        lda(va[ks], sV, j0, ks * 16, lane);
      }
      // validity of this thread's two key rows (rows of Sᵀ)
      const int jr[2] = {j0 + g, j0 + g + 8};
      bool kv[2];
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) kv[hh] = (kvalid[jr[hh] >> 5] >> (jr[hh] & 31)) & 1u;
      float dk[8][4], dv[8][4];
#pragma unroll
      for (int nt = 0; nt < 8; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) dk[nt][e] = dv[nt][e] = 0.0f;
      for (int i0 = 0; i0 < p.Lp; i0 += 64) {  // 64 queries at a time (columns of Sᵀ)
        float st[8][4], dpt[8][4];
        const int nv = min(4, (p.L - i0 + 15) >> 4);
        score_block(st, ka, sQ, i0, lane, nv);    // Sᵀ = K·Qᵀ
        score_block(dpt, va, sdO, i0, lane, nv);  // dPᵀ = V·dOᵀ
#pragma unroll
        for (int nt = 0; nt < 8; ++nt)
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int i = i0 + nt * 8 + 2 * t + (e & 1);  // query index
            const int hh = e >> 1;                        // key row g / g+8
            float v = st[nt][e] * p.scale;
            if (!kv[hh]) v += p.mask_neg;
            float pr = 0.0f, ds = 0.0f;
            if (jr[hh] < p.L && i < p.L) {
              pr = __expf(v - s_lse[i]);
              ds = pr * (dpt[nt][e] - s_delta[i]) * p.scale;
            }
            st[nt][e] = pr;    // Pᵀ
            dpt[nt][e] = ds;   // dSᵀ
          }
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {  // k = 16 queries
          if (ks >= nv) continue;
          uint32_t pa[4], sa[4];
          c2a(pa, st[2 * ks], st[2 * ks + 1]);
          c2a(sa, dpt[2 * ks], dpt[2 * ks + 1]);
#pragma unroll
          for (int np = 0; np < 4; ++np) {
            uint32_t b[4], c[4];
            ldb_kn(b, sQ, np * 16, i0 + ks * 16, lane);
            ldb_kn(c, sdO, np * 16, i0 + ks * 16, lane);
            const uint32_t b0[2] = {b[0], b[1]}, b1[2] = {b[2], b[3]};
            const uint32_t c0[2] = {c[0], c[1]}, c1[2] = {c[2], c[3]};
            mma_bf16_16816(dk[np * 2], sa, b0);
            mma_bf16_16816(dk[np * 2 + 1], sa, b1);
            mma_bf16_16816(dv[np * 2], pa, c0);
            mma_bf16_16816(dv[np * 2 + 1], pa, c1);
          }
        }
      }
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const int col = nt * 8 + 2 * t;
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          if (jr[hh] < p.L) {
            __nv_bfloat16* row = dq + static_cast<int64_t>(jr[hh]) * p.ld_qkv + col;
            *reinterpret_cast<uint32_t*>(row + Hd) = pack_bf16x2(dk[nt][hh * 2], dk[nt][hh * 2 + 1]);
            *reinterpret_cast<uint32_t*>(row + 2 * Hd) = pack_bf16x2(dv[nt][hh * 2], dv[nt][hh * 2 + 1]);
          }
        }
      }
    }
  }
}

// =====================================================================================================================
// Unmasked fast path (ViT: every token attends to every token).  Same data flow as the kernels above, rebuilt around
// what ncu showed on them: the forward was ISSUE-bound (61 % issue-active, ALU pipe 53 %) on per-element mask / bounds
// logic, the backward LATENCY-bound at 8 warps per SM (255 registers).  Here
//   * shared tiles hold round16(L) rows (208 for L = 197), zero-filled past L; key blocks run over real 16-key groups
//     only (13 instead of 16), the partial last block is a compile-time variant, and "key >= L -> -inf" is applied in
//     that block alone.  Zero rows make every other bounds check unnecessary: a padded key has K = V = 0, a padded
//     query has Q = dO = 0 and lse = delta = 0, so their contributions vanish in the products that consume them;
//   * softmax runs in the exp2 domain with the scale folded in: p = ex2(s * c - m * c), c = scale * log2(e): one FFMA
//     and one MUFU per score;
//   * the backward keeps <= 128 registers (32-key blocks in the dQ phase, 16-query blocks in the dK/dV phase) and
//     4 x 26 KB of tiles, so two CTAs (16 warps) share an SM and one CTA's loads overlap the other's math.
// =====================================================================================================================
constexpr float LOG2E = 1.4426950408889634f;


int check_mid(const a4r_attn_args* a) {
  A4R_CHECK_ARG(a != nullptr && a->qkv && a->out, "attention_mid: NULL pointer");
  A4R_CHECK_ARG(a->L >= 1 && a->L <= LMAX, "attention_mid: L must be in [1,256] (got %lld)", (long long)a->L);
  A4R_CHECK_ARG(a->head_dim == DH, "attention_mid: head_dim must be 64");
  A4R_CHECK_ARG(a->causal == 0, "attention_mid: causal masking is not supported (use the short-sequence kernel)");
  A4R_CHECK_ARG(a->dropout_p == 0.0f, "attention_mid: probability dropout is not implemented (ViT uses p = 0)");
  A4R_CHECK_ARG(a->heads >= 1 && a->N >= 0, "attention_mid: bad heads/N");
  A4R_CHECK_ARG(a->ld_qkv >= 3 * a->heads * DH && a->ld_qkv % 8 == 0 && a->ld_out >= a->heads * DH && a->ld_out % 8 == 0,
                "attention_mid: bad leading dimensions");
  A4R_CHECK_ARG(a4r_aligned16(a->qkv) && a4r_aligned16(a->out), "attention_mid: pointers must be 16B aligned");
  A4R_CHECK_ARG(a->mask_dtype >= 0 && a->mask_dtype <= 2, "attention_mid: mask_dtype must be 0,1,2");
  if (a->mask_dtype != 0) A4R_CHECK_ARG(a->mask != nullptr && a->mask_ld >= a->L, "attention_mid: bad mask");
  return a4r_device_check();
}

MidParams mid_params(const a4r_attn_args* a) {
  MidParams p;
  p.qkv = static_cast<const __nv_bfloat16*>(a->qkv);
  p.out = static_cast<__nv_bfloat16*>(a->out);
  p.dout = static_cast<const __nv_bfloat16*>(a->dout);
  p.mask = a->mask;
  p.lse = a->lse;
  p.ctx = static_cast<const __nv_bfloat16*>(a->ctx);
  p.ld_qkv = a->ld_qkv;
  p.ld_out = a->ld_out;
  p.mask_ld = a->mask_ld;
  p.N = static_cast<int>(a->N);
  p.L = static_cast<int>(a->L);
  p.Lp = (p.L + 63) & ~63;
  p.heads = static_cast<int>(a->heads);
  p.mask_dtype = a->mask_dtype;
  p.scale = a->scale;
  p.mask_neg = a->mask_neg;
  return p;
}

}  // namespace

int a4r_attn_vit_tc_fwd(const a4r_attn_args* a, cudaStream_t stream);   // attention_tc_sm100.cu
int a4r_attn_vit_tc_bwd(const a4r_attn_args* a, cudaStream_t stream);   // attention_tc_bwd_sm100.cu

extern "C" int a4r_attn_mid_fwd(const a4r_attn_args* a, a4r_stream_t stream) {
  int rc = check_mid(a);
  if (rc != A4R_OK) return rc;
  if (a->N == 0) return A4R_OK;
  if (a->mask_dtype == 0) return a4r_attn_vit_tc_fwd(a, static_cast<cudaStream_t>(stream));   // ViT: tcgen05 kernel
  const MidParams p = mid_params(a);
  const int smem = 3 * p.Lp * 128;
  A4R_CUDA_OK(cudaFuncSetAttribute(attn_mid_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * LMAX * 128));
  int64_t blocks = static_cast<int64_t>(p.N) * p.heads;
  const int64_t cap = static_cast<int64_t>(a4r_num_sms()) * 4;
  if (blocks > cap) blocks = cap;
  attn_mid_fwd_kernel<<<static_cast<int>(blocks), WARPS * 32, smem, static_cast<cudaStream_t>(stream)>>>(p);
  A4R_LAUNCH_OK();
  a4r_count_launch(1);
  return A4R_OK;
}

extern "C" int a4r_attn_mid_bwd(const a4r_attn_args* a, a4r_stream_t stream) {
  int rc = check_mid(a);
  if (rc != A4R_OK) return rc;
  A4R_CHECK_ARG(a->dout != nullptr && a4r_aligned16(a->dout), "attention_mid bwd: dout missing or unaligned");
  A4R_CHECK_ARG(a->lse != nullptr && a->ctx != nullptr, "attention_mid bwd: needs the forward's lse and ctx outputs");
  if (a->N == 0) return A4R_OK;
  if (a->mask_dtype == 0) return a4r_attn_vit_tc_bwd(a, static_cast<cudaStream_t>(stream));   // ViT: tcgen05 kernel
  const MidParams p = mid_params(a);
  const int smem = 4 * p.Lp * 128;
  A4R_CUDA_OK(cudaFuncSetAttribute(attn_mid_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * LMAX * 128));
  int64_t blocks = static_cast<int64_t>(p.N) * p.heads;
  const int64_t cap = static_cast<int64_t>(a4r_num_sms()) * 2;
  if (blocks > cap) blocks = cap;
  attn_mid_bwd_kernel<<<static_cast<int>(blocks), WARPS * 32, smem, static_cast<cudaStream_t>(stream)>>>(p);
  A4R_LAUNCH_OK();
  a4r_count_launch(1);
  return A4R_OK;
}
