// inbatch_ce.cu — K9-S: in-batch softmax loss with duplicate-item masking, forward / backward.
//
// Every valid position (b, s) is scored against ALL C = B*(S+1) history items of the batch,
//   logit[v, c] = <prec[b, s], cand[c]> - cand_bias[c],
// a candidate is overwritten with `masked_logit` (-1e4) if its slot is padding (log_mask == 0) or if its item id occurs
// anywhere in user b's own sequence — except the positive itself, column b*(S+1)+s+1 — and the loss is the mean over
// valid positions of  logsumexp_c(logit[v, :]) - logit[v, target].
// The reference (westlake-repl/Adapter4Rec) trains with BCE over one sampled negative (Downstream/Text/model/model.py:62-68)
// and has no such head; the semantics follow the in-batch debiased cross-entropy of the same group's IDvs.MoRec trainer
// ("parity unpinned"); oracle/transrec_oracle.py:inbatch_softmax_loss is the CPU statement.
//
// The [V, C] logit matrix (V = B*S; 440 MB in fp32 for a 512-user batch) is never materialised: all three kernels are
// flash-style — 64 x 64 logit tiles on mma.sync (the contraction length D = 64/128 cannot amortise a TMEM round trip),
// masks evaluated per tile from the item ids, an online log-sum-exp in the forward, P = exp(logit - lse) recomputed in the
// backward.  d_prec is accumulated per ROW tile, d_cand per CANDIDATE tile (a second kernel with the roles swapped), so
// there are no atomics and the result is deterministic.
#include "a4r_common.cuh"
#include "mma_sync.cuh"

namespace {

constexpr int IB_THREADS = 128;  // 4 warps x 16 rows = one 64-row tile
constexpr int IB_TILE = 64;
constexpr int IB_MAX_USERS = 64;  // users touched by one 64-row tile (S = 1 worst case)
constexpr float LOG2E = 1.44269504088896340736f;

A4R_DEVICE uint32_t toff(int row, int chunk) { return static_cast<uint32_t>(row * 128 + (((chunk ^ row) & 7) << 4)); }

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

struct IbParams {
  const __nv_bfloat16* prec;  // [V, D]
  const __nv_bfloat16* cand;  // row c at cand + c * ld_cand
  int64_t ld_cand;
  const int64_t* item_ids;    // [B, S+1] = [C]
  const float* log_mask;      // [B, S]
  const float* cand_bias;     // [C] or NULL
  float* lse;                 // [V]
  const float* count;         // bwd
  const float* grad_out;      // bwd, may be NULL
  __nv_bfloat16* d_prec;      // [V, D]
  __nv_bfloat16* d_cand;      // row c at d_cand + c * ld_dcand
  int64_t ld_dcand;
  int B, S, D, V, C;
  int nu_max;                 // user-id rows staged per 64-row tile
  float masked_logit;
};

// [rows x D] bf16 rows (row r at g + (row0 + r) * ld) -> NKC swizzled [64][64] sub-tiles; rows >= nrows are zero-filled
template <int NKC>
A4R_DEVICE void load_tile(uint8_t* tile, const __nv_bfloat16* g, int64_t ld, int row0, int nrows) {
  const uint32_t base = smem_u32(tile);
  for (int i = threadIdx.x; i < IB_TILE * 8 * NKC; i += IB_THREADS) {
    const int row = (i >> 3) % IB_TILE, ch = i & 7, kc = i / (IB_TILE * 8);
    const bool ok = row0 + row < nrows;
    const __nv_bfloat16* src = g + (ok ? static_cast<int64_t>(row0 + row) * ld + kc * 64 + ch * 8 : 0);
    const int sz = ok ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(base + kc * (IB_TILE * 128) + toff(row, ch)), "l"(src),
                 "r"(sz)
                 : "memory");
  }
}
A4R_DEVICE void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
A4R_DEVICE void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// candidate slot (b', j) is a real item iff j == S (the last item of a sequence is never padding) or log_mask[b', j] != 0
A4R_DEVICE bool cand_valid(const IbParams& p, int c) {
  if (c >= p.C) return false;
  const int b = c / (p.S + 1), j = c - b * (p.S + 1);
  return j == p.S || p.log_mask[static_cast<int64_t>(b) * p.S + j] != 0.0f;
}

// Per-tile candidate metadata used by the row-major kernels (forward, d_prec):
//   rej[u][w]  bit i = candidate (tile*64 + 32 w + i) carries an item id of local user u (duplicate of user u's sequence)
//   valid[w]   bit i = candidate is a real item;  bias[c] = cand_bias
struct CandMeta {
  uint32_t rej[IB_MAX_USERS][2];
  uint32_t valid[2];
  float bias[IB_TILE];
};

// uid: [nu][S+1] ids of the users this row tile touches (shared memory)
A4R_DEVICE void build_cand_meta(CandMeta& m, const IbParams& p, const int64_t* uid, int nu, int c0) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ca = c0 + lane, cb = c0 + 32 + lane;
  const int64_t ida = ca < p.C ? p.item_ids[ca] : -1, idb = cb < p.C ? p.item_ids[cb] : -1;
  for (int u = warp; u < nu; u += IB_THREADS / 32) {
    const int64_t* ids = uid + u * (p.S + 1);
    bool ha = false, hb = false;
    for (int k = 0; k <= p.S; ++k) {
      const int64_t v = ids[k];
      ha |= v == ida;
      hb |= v == idb;
    }
    const uint32_t wa = __ballot_sync(0xffffffffu, ha), wb = __ballot_sync(0xffffffffu, hb);
    if (lane == 0) {
      m.rej[u][0] = wa;
      m.rej[u][1] = wb;
    }
  }
  if (warp == 0) {
    const uint32_t va = __ballot_sync(0xffffffffu, cand_valid(p, ca)), vb = __ballot_sync(0xffffffffu, cand_valid(p, cb));
    if (lane == 0) {
      m.valid[0] = va;
      m.valid[1] = vb;
    }
  }
  if (p.cand_bias != nullptr && threadIdx.x < IB_TILE)
    m.bias[threadIdx.x] = (c0 + threadIdx.x < p.C) ? p.cand_bias[c0 + threadIdx.x] : 0.0f;
}

// logits of a 16-row tile (A fragments qa) against one 64-candidate tile stored [cand][d]
template <int NKC>
A4R_DEVICE void logits_block(float (&s)[8][4], const uint32_t (&qa)[NKC * 4][4], uint32_t sK, int lane) {
#pragma unroll
  for (int nt = 0; nt < 8; ++nt)
#pragma unroll
    for (int e = 0; e < 4; ++e) s[nt][e] = 0.0f;
#pragma unroll
  for (int ks = 0; ks < NKC * 4; ++ks) {
#pragma unroll
    for (int np = 0; np < 4; ++np) {
      uint32_t b[4];
      ldb_nk(b, sK + (ks >> 2) * (IB_TILE * 128), np * 16, (ks & 3) * 16, lane);
      const uint32_t b0[2] = {b[0], b[1]}, b1[2] = {b[2], b[3]};
      mma_bf16_16816(s[np * 2], qa[ks], b0);
      mma_bf16_16816(s[np * 2 + 1], qa[ks], b1);
    }
  }
}

// Row-side bookkeeping of the two fragment rows (g, g + 8) a thread owns in the row-major kernels
struct RowInfo {
  int tgt[2];    // target column
  int ul[2];     // local user index
  bool valid[2]; // log_mask != 0 and row < V
};
A4R_DEVICE RowInfo row_info(const IbParams& p, int v_first, int u0) {
  RowInfo r;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int v = v_first + h * 8;
    const int vc = v < p.V ? v : p.V - 1;
    const int b = vc / p.S, s = vc - b * p.S;
    r.tgt[h] = b * (p.S + 1) + s + 1;
    r.ul[h] = b - u0;
    r.valid[h] = v < p.V && p.log_mask[vc] != 0.0f;
  }
  return r;
}

// applies bias, masks and the column tail to a logits block; `s` becomes the final logits of the reference matrix
A4R_DEVICE void mask_block(float (&s)[8][4], const IbParams& p, const CandMeta& m, const RowInfo& ri, int c0, int lane) {
  const int t = lane & 3;
  const uint32_t r0[2] = {m.rej[ri.ul[0]][0], m.rej[ri.ul[0]][1]}, r1[2] = {m.rej[ri.ul[1]][0], m.rej[ri.ul[1]][1]};
  const uint32_t va[2] = {m.valid[0], m.valid[1]};
  const bool has_bias = p.cand_bias != nullptr;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt)
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int cl = nt * 8 + 2 * t + (e & 1), c = c0 + cl, h = e >> 1;
      float v = s[nt][e];
      if (has_bias) v -= m.bias[cl];
      const bool ok = (va[cl >> 5] >> (cl & 31)) & 1u;
      const bool dup = ((h ? r1[cl >> 5] : r0[cl >> 5]) >> (cl & 31)) & 1u;
      if (!ok || (dup && c != ri.tgt[h])) v = p.masked_logit;
      if (c >= p.C) v = -INFINITY;
      s[nt][e] = v;
    }
}

// users touched by rows [v0, v0 + 64): ids -> shared memory.  returns nu, sets u0
A4R_DEVICE int stage_user_ids(int64_t* uid, const IbParams& p, int v0, int& u0) {
  u0 = v0 / p.S;
  const int vlast = min(v0 + IB_TILE, p.V) - 1;
  const int nu = vlast / p.S - u0 + 1;
  const int n = nu * (p.S + 1);
  const int64_t* src = p.item_ids + static_cast<int64_t>(u0) * (p.S + 1);
  for (int i = threadIdx.x; i < n; i += IB_THREADS) uid[i] = src[i];
  return nu;
}

// ---------------------------------------------------------------------------------------------------------------------
// forward (MODE 0): lse[v], per-CTA loss partial.   d_prec (MODE 1): d_prec[v, :] = w * sum_c (P - onehot)[v, c] cand[c, :]
// ---------------------------------------------------------------------------------------------------------------------
template <int NKC, int MODE>
__global__ void __launch_bounds__(IB_THREADS) ibce_rows_kernel(const IbParams p, float* __restrict__ partial) {
  extern __shared__ __align__(128) uint8_t sm_ib[];
  constexpr int TILE_BYTES = NKC * IB_TILE * 128;
  uint8_t* tQ = sm_ib;
  uint8_t* tK = sm_ib + TILE_BYTES;  // two buffers
  CandMeta* meta = reinterpret_cast<CandMeta*>(sm_ib + 3 * TILE_BYTES);
  int64_t* uid = reinterpret_cast<int64_t*>(sm_ib + 3 * TILE_BYTES + 2 * sizeof(CandMeta));
  __shared__ float row_loss[IB_TILE];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int v0 = blockIdx.x * IB_TILE;
  int u0;
  const int nu = stage_user_ids(uid, p, v0, u0);
  load_tile<NKC>(tQ, p.prec, p.D, v0, p.V);
  load_tile<NKC>(tK, p.cand, p.ld_cand, 0, p.C);
  cp_async_commit();
  __syncthreads();  // uid visible
  build_cand_meta(meta[0], p, uid, nu, 0);
  cp_async_wait<0>();
  __syncthreads();

  uint32_t qa[NKC * 4][4];
#pragma unroll
  for (int ks = 0; ks < NKC * 4; ++ks) lda(qa[ks], smem_u32(tQ) + (ks >> 2) * (IB_TILE * 128), warp * 16, (ks & 3) * 16, lane);
  const RowInfo ri = row_info(p, v0 + warp * 16 + g, u0);

  float mrow[2] = {-INFINITY, -INFINITY}, lrow[2] = {0.0f, 0.0f}, tg[2] = {0.0f, 0.0f};
  float lse_r[2] = {0.0f, 0.0f};
  float acc[NKC * 8][4];
  if (MODE == 1) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int v = v0 + warp * 16 + g + h * 8;
      lse_r[h] = v < p.V ? p.lse[v] : 0.0f;
    }
#pragma unroll
    for (int i = 0; i < NKC * 8; ++i)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[i][e] = 0.0f;
  }

  const int nct = (p.C + IB_TILE - 1) / IB_TILE;
  for (int j = 0; j < nct; ++j) {
    const int b = j & 1;
    if (j + 1 < nct) {
      load_tile<NKC>(tK + (b ^ 1) * TILE_BYTES, p.cand, p.ld_cand, (j + 1) * IB_TILE, p.C);
      cp_async_commit();
      build_cand_meta(meta[b ^ 1], p, uid, nu, (j + 1) * IB_TILE);
    }
    float s[8][4];
    logits_block<NKC>(s, qa, smem_u32(tK + b * TILE_BYTES), lane);
    mask_block(s, p, meta[b], ri, j * IB_TILE, lane);
    if (MODE == 0) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        float mx = -INFINITY;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) mx = fmaxf(mx, fmaxf(s[nt][2 * h], s[nt][2 * h + 1]));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
        const float mnew = fmaxf(mrow[h], mx);
        float sum = 0.0f;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const float x = s[nt][2 * h + e];
            sum += ex2_approx((x - mnew) * LOG2E);
            if (j * IB_TILE + nt * 8 + 2 * t + e == ri.tgt[h]) tg[h] = x;
          }
        lrow[h] = lrow[h] * ex2_approx((mrow[h] - mnew) * LOG2E) + sum;
        mrow[h] = mnew;
      }
    } else {
      // P - onehot, zero where the logit was overwritten by the mask constant (no gradient flows through a constant)
#pragma unroll
      for (int nt = 0; nt < 8; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int h = e >> 1, c = j * IB_TILE + nt * 8 + 2 * t + (e & 1);
          const float x = s[nt][e];
          float pv = 0.0f;
          if (x != p.masked_logit && ri.valid[h]) {
            pv = ex2_approx((x - lse_r[h]) * LOG2E);
            if (c == ri.tgt[h]) pv -= 1.0f;
          }
          s[nt][e] = pv;
        }
      const uint32_t sK = smem_u32(tK + b * TILE_BYTES);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        uint32_t pa[4];
        c2a(pa, s[2 * kk], s[2 * kk + 1]);
#pragma unroll
        for (int dn = 0; dn < NKC * 4; ++dn) {
          uint32_t bb[4];
          ldb_kn(bb, sK + (dn >> 2) * (IB_TILE * 128), (dn & 3) * 16, kk * 16, lane);
          const uint32_t b0[2] = {bb[0], bb[1]}, b1[2] = {bb[2], bb[3]};
          mma_bf16_16816(acc[dn * 2], pa, b0);
          mma_bf16_16816(acc[dn * 2 + 1], pa, b1);
        }
      }
    }
    if (j + 1 < nct) cp_async_wait<0>();
    __syncthreads();  // tile j+1 and its metadata are visible; everyone is done with buffer b
  }

  if (MODE == 0) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      lrow[h] += __shfl_xor_sync(0xffffffffu, lrow[h], 1);
      lrow[h] += __shfl_xor_sync(0xffffffffu, lrow[h], 2);
      tg[h] += __shfl_xor_sync(0xffffffffu, tg[h], 1);
      tg[h] += __shfl_xor_sync(0xffffffffu, tg[h], 2);
      const float lse = mrow[h] + __logf(lrow[h]);
      const int rl = warp * 16 + g + h * 8, v = v0 + rl;
      if (t == 0) {
        if (v < p.V) p.lse[v] = lse;
        row_loss[rl] = ri.valid[h] ? lse - tg[h] : 0.0f;
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      float sum = 0.0f;
      for (int i = 0; i < IB_TILE; ++i) sum += row_loss[i];  // fixed order: deterministic
      partial[blockIdx.x] = sum;
    }
  } else {
    const float w = (p.grad_out != nullptr ? *p.grad_out : 1.0f) / *p.count;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int v = v0 + warp * 16 + g + h * 8;
      if (v < p.V) {
        __nv_bfloat16* dst = p.d_prec + static_cast<int64_t>(v) * p.D;
#pragma unroll
        for (int dn = 0; dn < NKC * 8; ++dn)
          *reinterpret_cast<uint32_t*>(dst + dn * 8 + 2 * t) = pack_bf16x2(acc[dn][2 * h] * w, acc[dn][2 * h + 1] * w);
      }
    }
  }
}

// loss = sum(partials) / count, count = |{log_mask != 0}|; one block, fixed-order tree
__global__ void __launch_bounds__(256) ibce_finalize_kernel(const float* __restrict__ partial, int n, const float* __restrict__ log_mask,
                                                            int V, float* __restrict__ loss, float* __restrict__ count) {
  __shared__ float sl[256], sc[256];
  float a = 0.0f, c = 0.0f;
  for (int i = threadIdx.x; i < n; i += 256) a += partial[i];
  for (int i = threadIdx.x; i < V; i += 256) c += log_mask[i] != 0.0f ? 1.0f : 0.0f;
  sl[threadIdx.x] = a;
  sc[threadIdx.x] = c;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      sl[threadIdx.x] += sl[threadIdx.x + o];
      sc[threadIdx.x] += sc[threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    *count = sc[0];
    *loss = sl[0] / sc[0];  // 0/0 = nan for an all-padding batch, like CrossEntropyLoss over an empty selection
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// d_cand: each CTA owns 64 candidates and walks all row tiles: d_cand[c, :] = w * sum_v (P - onehot)[v, c] prec[v, :]
// ---------------------------------------------------------------------------------------------------------------------
struct RowMeta {
  float lse[IB_TILE];
  int tgt[IB_TILE];
  int ul[IB_TILE];              // local user index, -1 = row contributes nothing (padding position / tail)
  unsigned long long rej[2][IB_TILE];  // [half][cand]: bit u = candidate duplicates an item of local user u
};

template <int NKC>
__global__ void __launch_bounds__(IB_THREADS) ibce_cands_kernel(const IbParams p) {
  extern __shared__ __align__(128) uint8_t sm_ib[];
  constexpr int TILE_BYTES = NKC * IB_TILE * 128;
  uint8_t* tK = sm_ib;
  uint8_t* tQ = sm_ib + TILE_BYTES;  // two buffers
  RowMeta* meta = reinterpret_cast<RowMeta*>(sm_ib + 3 * TILE_BYTES);
  int64_t* uid = reinterpret_cast<int64_t*>(sm_ib + 3 * TILE_BYTES + 2 * sizeof(RowMeta));  // two buffers of nu_max*(S+1)
  const int uid_stride = p.nu_max * (p.S + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int c0 = blockIdx.x * IB_TILE;
  const int my_c = c0 + (threadIdx.x & 63);
  const int64_t my_id = my_c < p.C ? p.item_ids[my_c] : -1;

  // fills meta[b] / uid[b] for row tile i (everything but rej, which needs uid in shared memory first)
  auto stage_rows = [&](int i, int b) {
    int u0;
    stage_user_ids(uid + b * uid_stride, p, i * IB_TILE, u0);
    if (threadIdx.x < IB_TILE) {
      const int v = i * IB_TILE + threadIdx.x;
      RowMeta& m = meta[b];
      if (v < p.V) {
        const int bb = v / p.S, s = v - bb * p.S;
        m.lse[threadIdx.x] = p.lse[v];
        m.tgt[threadIdx.x] = bb * (p.S + 1) + s + 1;
        m.ul[threadIdx.x] = p.log_mask[v] != 0.0f ? bb - u0 : -1;
      } else {
        m.lse[threadIdx.x] = 0.0f;
        m.tgt[threadIdx.x] = -1;
        m.ul[threadIdx.x] = -1;
      }
    }
  };
  auto build_rej = [&](int i, int b) {
    const int v0 = i * IB_TILE;
    const int u0 = v0 / p.S, nu = (min(v0 + IB_TILE, p.V) - 1) / p.S - u0 + 1;
    const int64_t* ids = uid + b * uid_stride;
    unsigned long long bits = 0;
    for (int u = threadIdx.x >> 6; u < nu; u += 2) {
      bool hit = false;
      for (int k = 0; k <= p.S; ++k) hit |= ids[u * (p.S + 1) + k] == my_id;
      bits |= static_cast<unsigned long long>(hit) << u;
    }
    meta[b].rej[threadIdx.x >> 6][threadIdx.x & 63] = bits;
  };

  load_tile<NKC>(tK, p.cand, p.ld_cand, c0, p.C);
  load_tile<NKC>(tQ, p.prec, p.D, 0, p.V);
  cp_async_commit();
  stage_rows(0, 0);
  __syncthreads();
  build_rej(0, 0);
  cp_async_wait<0>();
  __syncthreads();

  uint32_t ka[NKC * 4][4];
#pragma unroll
  for (int ks = 0; ks < NKC * 4; ++ks) lda(ka[ks], smem_u32(tK) + (ks >> 2) * (IB_TILE * 128), warp * 16, (ks & 3) * 16, lane);
  // candidate-side constants of the two fragment rows (g, g + 8)
  bool cok[2];
  float cbias[2];
  int cidx[2];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    cidx[h] = c0 + warp * 16 + g + h * 8;
    cok[h] = cand_valid(p, cidx[h]);
    cbias[h] = (p.cand_bias != nullptr && cidx[h] < p.C) ? p.cand_bias[cidx[h]] : 0.0f;
  }
  float acc[NKC * 8][4];
#pragma unroll
  for (int i = 0; i < NKC * 8; ++i)
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[i][e] = 0.0f;

  const int nrt = (p.V + IB_TILE - 1) / IB_TILE;
  for (int i = 0; i < nrt; ++i) {
    const int b = i & 1;
    if (i + 1 < nrt) {
      load_tile<NKC>(tQ + (b ^ 1) * TILE_BYTES, p.prec, p.D, (i + 1) * IB_TILE, p.V);
      cp_async_commit();
      stage_rows(i + 1, b ^ 1);
    }
    const RowMeta& m = meta[b];
    const uint32_t sQ = smem_u32(tQ + b * TILE_BYTES);
    float s[8][4];
    logits_block<NKC>(s, ka, sQ, lane);  // s[nt][e]: candidate row g + 8 (e >> 1), position column nt * 8 + 2 t + (e & 1)
    unsigned long long rj[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) rj[h] = m.rej[0][warp * 16 + g + h * 8] | m.rej[1][warp * 16 + g + h * 8];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int h = e >> 1, rl = nt * 8 + 2 * t + (e & 1);
        const int ul = m.ul[rl];
        float pv = 0.0f;
        if (ul >= 0 && cidx[h] < p.C) {
          const bool is_t = cidx[h] == m.tgt[rl];
          const bool masked = !cok[h] || (((rj[h] >> ul) & 1ull) && !is_t);
          const float x = s[nt][e] - cbias[h];
          // the x == masked_logit test mirrors the forward / d_prec kernels, which see the constant, not the mask bit
          if (!masked && x != p.masked_logit) {
            pv = ex2_approx((x - m.lse[rl]) * LOG2E);
            if (is_t) pv -= 1.0f;
          }
        }
        s[nt][e] = pv;
      }
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      uint32_t pa[4];
      c2a(pa, s[2 * kk], s[2 * kk + 1]);
#pragma unroll
      for (int dn = 0; dn < NKC * 4; ++dn) {
        uint32_t bb[4];
        ldb_kn(bb, sQ + (dn >> 2) * (IB_TILE * 128), (dn & 3) * 16, kk * 16, lane);
        const uint32_t b0[2] = {bb[0], bb[1]}, b1[2] = {bb[2], bb[3]};
        mma_bf16_16816(acc[dn * 2], pa, b0);
        mma_bf16_16816(acc[dn * 2 + 1], pa, b1);
      }
    }
    __syncthreads();  // uid[b^1] staged; everyone is done with buffer b
    if (i + 1 < nrt) {
      build_rej(i + 1, b ^ 1);
      cp_async_wait<0>();
    }
    __syncthreads();
  }

  const float w = (p.grad_out != nullptr ? *p.grad_out : 1.0f) / *p.count;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    if (cidx[h] < p.C) {
      __nv_bfloat16* dst = p.d_cand + static_cast<int64_t>(cidx[h]) * p.ld_dcand;
#pragma unroll
      for (int dn = 0; dn < NKC * 8; ++dn)
        *reinterpret_cast<uint32_t*>(dst + dn * 8 + 2 * t) = pack_bf16x2(acc[dn][2 * h] * w, acc[dn][2 * h + 1] * w);
    }
  }
}

int nu_max_for(int64_t S) {
  int64_t n = (IB_TILE - 1) / S + 2;
  return static_cast<int>(n > IB_MAX_USERS ? IB_MAX_USERS : n);
}

int check_ib(const a4r_inbatch_ce_args* a, IbParams* p) {
  A4R_CHECK_ARG(a != nullptr, "inbatch_ce: args is NULL");
  A4R_CHECK_ARG(a->prec && a->cand && a->item_ids && a->log_mask && a->lse && a->loss && a->count, "inbatch_ce: NULL pointer");
  A4R_CHECK_ARG(a->B >= 1 && a->S >= 1 && a->S <= 1024, "inbatch_ce: bad B/S");
  A4R_CHECK_ARG(a->D == 64 || a->D == 128, "inbatch_ce: D must be 64 or 128 (got %lld)", static_cast<long long>(a->D));
  A4R_CHECK_ARG(a->B * (a->S + 1) < (1ll << 31), "inbatch_ce: too many candidates");
  A4R_CHECK_ARG(a->ld_cand >= a->D && a->ld_cand % 8 == 0 && a4r_aligned16(a->prec) && a4r_aligned16(a->cand),
                "inbatch_ce: prec/cand must be 16B aligned with ld_cand %% 8 == 0");
  p->prec = static_cast<const __nv_bfloat16*>(a->prec);
  p->cand = static_cast<const __nv_bfloat16*>(a->cand);
  p->ld_cand = a->ld_cand;
  p->item_ids = a->item_ids;
  p->log_mask = a->log_mask;
  p->cand_bias = a->cand_bias;
  p->lse = a->lse;
  p->count = a->count;
  p->grad_out = nullptr;
  p->d_prec = nullptr;
  p->d_cand = nullptr;
  p->ld_dcand = 0;
  p->B = static_cast<int>(a->B);
  p->S = static_cast<int>(a->S);
  p->D = static_cast<int>(a->D);
  p->V = static_cast<int>(a->B * a->S);
  p->C = static_cast<int>(a->B * (a->S + 1));
  p->nu_max = nu_max_for(a->S);
  p->masked_logit = a->masked_logit;
  return a4r_device_check();
}

size_t rows_smem(const IbParams& p) {
  return static_cast<size_t>(3) * (p.D / 64) * IB_TILE * 128 + 2 * sizeof(CandMeta) +
         static_cast<size_t>(p.nu_max) * (p.S + 1) * sizeof(int64_t);
}
size_t cands_smem(const IbParams& p) {
  return static_cast<size_t>(3) * (p.D / 64) * IB_TILE * 128 + 2 * sizeof(RowMeta) +
         static_cast<size_t>(2) * p.nu_max * (p.S + 1) * sizeof(int64_t);
}

template <typename K>
int set_smem(K kernel, size_t bytes) {
  A4R_CHECK_ARG(bytes <= 227 * 1024, "inbatch_ce: S too large for the shared-memory id staging (%zu bytes)", bytes);
  A4R_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes)));
  return A4R_OK;
}

}  // namespace

extern "C" size_t a4r_inbatch_ce_workspace_bytes(int64_t B, int64_t S) {
  return static_cast<size_t>((B * S + IB_TILE - 1) / IB_TILE) * sizeof(float);
}

extern "C" int a4r_inbatch_ce_fwd(const a4r_inbatch_ce_args* a, void* workspace, size_t workspace_bytes, a4r_stream_t stream_) {
  IbParams p;
  int rc = check_ib(a, &p);
  if (rc != A4R_OK) return rc;
  if (workspace == nullptr || workspace_bytes < a4r_inbatch_ce_workspace_bytes(a->B, a->S))
    return a4r_set_error(A4R_EWORKSPACE, "inbatch_ce: workspace too small");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int grid = (p.V + IB_TILE - 1) / IB_TILE;
  const size_t smem = rows_smem(p);
  float* partial = static_cast<float*>(workspace);
  if (p.D == 64) {
    if ((rc = set_smem(ibce_rows_kernel<1, 0>, smem)) != A4R_OK) return rc;
    ibce_rows_kernel<1, 0><<<grid, IB_THREADS, smem, stream>>>(p, partial);
  } else {
    if ((rc = set_smem(ibce_rows_kernel<2, 0>, smem)) != A4R_OK) return rc;
    ibce_rows_kernel<2, 0><<<grid, IB_THREADS, smem, stream>>>(p, partial);
  }
  A4R_LAUNCH_OK();
  ibce_finalize_kernel<<<1, 256, 0, stream>>>(partial, grid, p.log_mask, p.V, a->loss, a->count);
  A4R_LAUNCH_OK();
  a4r_count_launch(2);
  return A4R_OK;
}

extern "C" int a4r_inbatch_ce_bwd(const a4r_inbatch_ce_args* a, const float* grad_out, void* d_prec, void* d_cand,
                                  int64_t ld_dcand, a4r_stream_t stream_) {
  IbParams p;
  int rc = check_ib(a, &p);
  if (rc != A4R_OK) return rc;
  A4R_CHECK_ARG(d_prec && d_cand && a4r_aligned16(d_prec) && a4r_aligned16(d_cand) && ld_dcand >= a->D && ld_dcand % 8 == 0,
                "inbatch_ce bwd: bad outputs");
  p.grad_out = grad_out;
  p.d_prec = static_cast<__nv_bfloat16*>(d_prec);
  p.d_cand = static_cast<__nv_bfloat16*>(d_cand);
  p.ld_dcand = ld_dcand;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int grid_r = (p.V + IB_TILE - 1) / IB_TILE, grid_c = (p.C + IB_TILE - 1) / IB_TILE;
  const size_t smem_r = rows_smem(p), smem_c = cands_smem(p);
  if (p.D == 64) {
    if ((rc = set_smem(ibce_rows_kernel<1, 1>, smem_r)) != A4R_OK) return rc;
    if ((rc = set_smem(ibce_cands_kernel<1>, smem_c)) != A4R_OK) return rc;
    ibce_rows_kernel<1, 1><<<grid_r, IB_THREADS, smem_r, stream>>>(p, nullptr);
    A4R_LAUNCH_OK();
    ibce_cands_kernel<1><<<grid_c, IB_THREADS, smem_c, stream>>>(p);
  } else {
    if ((rc = set_smem(ibce_rows_kernel<2, 1>, smem_r)) != A4R_OK) return rc;
    if ((rc = set_smem(ibce_cands_kernel<2>, smem_c)) != A4R_OK) return rc;
    ibce_rows_kernel<2, 1><<<grid_r, IB_THREADS, smem_r, stream>>>(p, nullptr);
    A4R_LAUNCH_OK();
    ibce_cands_kernel<2><<<grid_c, IB_THREADS, smem_c, stream>>>(p);
  }
  A4R_LAUNCH_OK();
  a4r_count_launch(2);
  return A4R_OK;
}
