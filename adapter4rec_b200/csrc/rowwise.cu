// rowwise.cu — HBM-bound row kernels: K1 (embedding gather + LayerNorm), K6 (LayerNorm fwd/bwd with an
// optional residual), activation gradients and small utility reductions.  All use 128-bit global access,
// fp32 statistics, bf16 activations; a row is owned by G = 8/16/32 lanes so that H = 64 (SASRec) does not
// waste 3/4 of a warp.
#include <type_traits>

#include "a4r_common.cuh"

namespace {

constexpr int ROW_THREADS = 256;

template <int G>
A4R_DEVICE float group_sum(float v) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

A4R_DEVICE void unpack8(const uint4& r, float (&f)[8]) {
  float2 t;
  t = unpack_bf16x2(r.x); f[0] = t.x; f[1] = t.y;
  t = unpack_bf16x2(r.y); f[2] = t.x; f[3] = t.y;
  t = unpack_bf16x2(r.z); f[4] = t.x; f[5] = t.y;
  t = unpack_bf16x2(r.w); f[6] = t.x; f[7] = t.y;
}
A4R_DEVICE uint4 pack8(const float (&f)[8]) {
  uint4 o;
  o.x = pack_bf16x2(f[0], f[1]); o.y = pack_bf16x2(f[2], f[3]);
  o.z = pack_bf16x2(f[4], f[5]); o.w = pack_bf16x2(f[6], f[7]);
  return o;
}

// ------------------------------------------------------------------------------------------------
// LayerNorm forward:  z = x (+ res[row % res_rows]);  y = (z - mean) * rstd * gamma + beta
// ------------------------------------------------------------------------------------------------
template <int G, int CPL>
__global__ void __launch_bounds__(ROW_THREADS) ln_fwd_kernel(const __nv_bfloat16* __restrict__ x,
                                                            const __nv_bfloat16* __restrict__ res, int64_t res_rows,
                                                            const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, float eps,
                                                            __nv_bfloat16* __restrict__ y,
                                                            __nv_bfloat16* __restrict__ z_out,
                                                            float* __restrict__ mean_out, float* __restrict__ rstd_out,
                                                            int64_t M, int H) {
  const int nchunks = H >> 3;
  const int sub = threadIdx.x % G;
  const int64_t rows_per_block = ROW_THREADS / G;
  const float invH = 1.0f / static_cast<float>(H);
  // the loop trip count is uniform across the warp (group shuffles use the full mask): out-of-range row
  // slots recompute row M-1 and skip their stores.
  // gamma / beta live in shared memory (one fill per block) instead of 4 L1 loads per chunk and row
  __shared__ __align__(16) float s_gamma[1024], s_beta[1024];
  for (int i = threadIdx.x; i < H; i += ROW_THREADS) {
    s_gamma[i] = gamma[i];
    s_beta[i] = beta[i];
  }
  __syncthreads();
  // The NEXT row's x chunks are fetched (still packed: CPL x 16 bytes) before the current row is reduced, so every warp
  // keeps two rows of loads in flight: the kernel is latency-bound otherwise (ncu: long_scoreboard 10.8 per issue).
  const int64_t row_step = gridDim.x * rows_per_block;
  uint4 nxt[CPL];
  {
    const int64_t r0 = blockIdx.x * rows_per_block + threadIdx.x / G;
    const int64_t rr = r0 < M ? r0 : M - 1;
#pragma unroll
    for (int c = 0; c < CPL; ++c) {
      const int ch = sub + c * G;
      if (ch < nchunks) nxt[c] = ld_nc_v4(x + rr * H + ch * 8);
    }
  }
  for (int64_t base = blockIdx.x * rows_per_block; base < M; base += row_step) {
    const int64_t row_raw = base + threadIdx.x / G;
    const bool valid = row_raw < M;
    const int64_t row = valid ? row_raw : M - 1;
    float v[CPL][8];
    float s = 0.0f;
#pragma unroll
    for (int c = 0; c < CPL; ++c) {
      const int ch = sub + c * G;
      if (ch < nchunks) unpack8(nxt[c], v[c]);
    }
    if (base + row_step < M) {   // uniform across the block
      const int64_t rn_raw = row_raw + row_step;
      const int64_t rn = rn_raw < M ? rn_raw : M - 1;
#pragma unroll
      for (int c = 0; c < CPL; ++c) {
        const int ch = sub + c * G;
        if (ch < nchunks) nxt[c] = ld_nc_v4(x + rn * H + ch * 8);
      }
    }
#pragma unroll
    for (int c = 0; c < CPL; ++c) {
      const int ch = sub + c * G;
      if (ch < nchunks) {
        if (res != nullptr) {
          float r[8];
          unpack8(ld_nc_v4(res + (row % res_rows) * H + ch * 8), r);
#pragma unroll
          for (int e = 0; e < 8; ++e) v[c][e] += r[e];
        }
        if (z_out != nullptr) {
          const uint4 zz = pack8(v[c]);
          if (valid) st_na_v4(z_out + row * H + ch * 8, zz);
          unpack8(zz, v[c]);  // normalise exactly what the backward will read
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) s += v[c][e];
      }
    }
    const float mean = group_sum<G>(s) * invH;
    float q = 0.0f;
#pragma unroll
    for (int c = 0; c < CPL; ++c) {
      const int ch = sub + c * G;
      if (ch < nchunks) {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float d = v[c][e] - mean;
          q += d * d;
        }
      }
    }
    const float rstd = rsqrtf(group_sum<G>(q) * invH + eps);
#pragma unroll
    for (int c = 0; c < CPL; ++c) {
      const int ch = sub + c * G;
      if (ch < nchunks) {
        const float4 g0 = *reinterpret_cast<const float4*>(s_gamma + ch * 8);
        const float4 g1 = *reinterpret_cast<const float4*>(s_gamma + ch * 8 + 4);
        const float4 b0 = *reinterpret_cast<const float4*>(s_beta + ch * 8);
        const float4 b1 = *reinterpret_cast<const float4*>(s_beta + ch * 8 + 4);
        const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
        const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
        float o[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = (v[c][e] - mean) * rstd * gg[e] + bb[e];
        if (valid) st_na_v4(y + row * H + ch * 8, pack8(o));
      }
    }
    if (sub == 0 && valid) {
      if (mean_out != nullptr) mean_out[row] = mean;
      if (rstd_out != nullptr) rstd_out[row] = rstd;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// LayerNorm backward: dz = rstd * (g*dy - mean(g*dy) - xhat * mean(g*dy*xhat));  optional dgamma/dbeta partials.
// Inputs stay packed (bf16) in registers between the statistics pass and the output pass and gamma lives in shared
// memory, so the frozen-LayerNorm instantiation (WGRAD = false) fits 3 CTAs per SM and keeps enough loads in flight
// to approach HBM bandwidth.
// ------------------------------------------------------------------------------------------------
// SKIP: dz also receives `dskip` (the gradient that reached the LayerNorm's INPUT over a skip connection that branches
// off before it — pre-LN blocks: x1 = x + f(LN(x))), so autograd never runs a separate add over [M, H].
template <int G, int CPL, bool WGRAD, bool SKIP = false>
__global__ void __launch_bounds__(ROW_THREADS, WGRAD ? 1 : (SKIP ? 2 : 3))
ln_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ z,
              const float* __restrict__ mean_in, const float* __restrict__ rstd_in, const float* __restrict__ gamma,
              __nv_bfloat16* __restrict__ dz, float* __restrict__ partial /* [grid,2,H] or NULL */, int64_t M, int H,
              __nv_bfloat16* __restrict__ dz_masked, uint32_t thr16, float dscale, uint64_t seed, uint64_t offset,
              const __nv_bfloat16* __restrict__ dskip = nullptr) {
  __shared__ __align__(16) float sgamma[1024];
  __shared__ float buf[WGRAD ? 8192 : 1];  // (ROW_THREADS/G) * H <= 8192 floats for every supported (G, H)
  const int nchunks = H >> 3;
  const int sub = threadIdx.x % G;
  const int64_t rows_per_block = ROW_THREADS / G;
  const float invH = 1.0f / static_cast<float>(H);
  for (int j = threadIdx.x; j < H; j += ROW_THREADS) sgamma[j] = gamma[j];
  __syncthreads();
  float dg[WGRAD ? CPL : 1][8], db[WGRAD ? CPL : 1][8];
  if constexpr (WGRAD) {
#pragma unroll
    for (int c = 0; c < CPL; ++c)
#pragma unroll
      for (int e = 0; e < 8; ++e) dg[c][e] = db[c][e] = 0.0f;
  }
  // the loop trip count is uniform across the warp (group shuffles use the full mask): out-of-range row
  // slots recompute row M-1 and skip their stores.
  for (int64_t base = blockIdx.x * rows_per_block; base < M; base += gridDim.x * rows_per_block) {
    const int64_t row_raw = base + threadIdx.x / G;
    const bool valid = row_raw < M;
    const int64_t row = valid ? row_raw : M - 1;
    uint4 pdy[CPL], pz[CPL], psk[SKIP ? CPL : 1];
#pragma unroll
    for (int c = 0; c < CPL; ++c) {
      const int ch = sub + c * G;
      if (ch < nchunks) {
        pdy[c] = ld_nc_v4(dy + row * H + ch * 8);
        pz[c] = ld_nc_v4(z + row * H + ch * 8);
        if constexpr (SKIP) psk[c] = ld_nc_v4(dskip + row * H + ch * 8);
      }
    }
    const float mean = mean_in[row], rstd = rstd_in[row];
    float s1 = 0.0f, s2 = 0.0f;
#pragma unroll
    for (int c = 0; c < CPL; ++c) {
      const int ch = sub + c * G;
      if (ch < nchunks) {
        float a[8], b[8];
        unpack8(pdy[c], a);
        unpack8(pz[c], b);
        const float4 g0 = *reinterpret_cast<const float4*>(sgamma + ch * 8);
        const float4 g1 = *reinterpret_cast<const float4*>(sgamma + ch * 8 + 4);
        const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float xh = (b[e] - mean) * rstd;
          const float gy = a[e] * gg[e];
          s1 += gy;
          s2 += gy * xh;
          if constexpr (WGRAD) {
            if (valid) {
              dg[c][e] += a[e] * xh;
              db[c][e] += a[e];
            }
          }
        }
      }
    }
    s1 = group_sum<G>(s1) * invH;
    s2 = group_sum<G>(s2) * invH;
#pragma unroll
    for (int c = 0; c < CPL; ++c) {
      const int ch = sub + c * G;
      if (ch < nchunks) {
        float a[8], b[8], o[8];
        unpack8(pdy[c], a);
        unpack8(pz[c], b);
        const float4 g0 = *reinterpret_cast<const float4*>(sgamma + ch * 8);
        const float4 g1 = *reinterpret_cast<const float4*>(sgamma + ch * 8 + 4);
        const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = rstd * (a[e] * gg[e] - s1 - (b[e] - mean) * rstd * s2);
        if constexpr (SKIP) {
          float sk[8];
          unpack8(psk[c], sk);
#pragma unroll
          for (int e = 0; e < 8; ++e) o[e] += sk[e];
        }
        if (valid) st_na_v4(dz + row * H + ch * 8, pack8(o));
        if (dz_masked != nullptr && valid) {
          // the gradient that flows back through the dropout in front of this LayerNorm's residual add
          const uint64_t ctr = offset + ((static_cast<uint64_t>(row) * H + ch * 8) >> 2);
          const uint64_t sd = rng_seed(seed);
          const uint64_t r0 = rng64(sd, ctr), r1 = rng64(sd, ctr + 1);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            o[e] = rng_keep(r0, e, thr16) ? o[e] * dscale : 0.0f;
            o[4 + e] = rng_keep(r1, e, thr16) ? o[4 + e] * dscale : 0.0f;
          }
          st_na_v4(dz_masked + row * H + ch * 8, pack8(o));
        }
      }
    }
  }
  if constexpr (WGRAD) {
    // block-level reduction of the per-thread column sums over the ROW_THREADS/G row slots (fixed order =>
    // deterministic), one pass for dgamma and one for dbeta through the same shared buffer.
    const int slot = threadIdx.x / G;
    for (int which = 0; which < 2; ++which) {
      __syncthreads();
#pragma unroll
      for (int c = 0; c < CPL; ++c) {
        const int ch = sub + c * G;
        if (ch < nchunks) {
#pragma unroll
          for (int e = 0; e < 8; ++e) buf[slot * H + ch * 8 + e] = which == 0 ? dg[c][e] : db[c][e];
        }
      }
      __syncthreads();
      for (int j = threadIdx.x; j < H; j += ROW_THREADS) {
        float acc = 0.0f;
        for (int r = 0; r < ROW_THREADS / G; ++r) acc += buf[r * H + j];
        partial[(static_cast<int64_t>(blockIdx.x) * 2 + which) * H + j] = acc;
      }
    }
  }
}

// dgamma[j] = sum_b partial[b,0,j], dbeta[j] = sum_b partial[b,1,j]  (fixed order => deterministic)
__global__ void ln_reduce_kernel(const float* __restrict__ partial, float* __restrict__ dgamma,
                                 float* __restrict__ dbeta, int nblocks, int H, int accumulate) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= 2 * H) return;
  const int which = j / H, col = j % H;
  float acc = 0.0f;
  for (int b = 0; b < nblocks; ++b) acc += partial[(static_cast<int64_t>(b) * 2 + which) * H + col];
  float* out = which == 0 ? dgamma : dbeta;
  out[col] = accumulate ? out[col] + acc : acc;
}

// out[j] = sum_b partial[b, j]  (fixed order => deterministic).  32 x 32 threads: thread (x, y) sums rows y, y+32, ...
// of column x (32 independent coalesced load streams per column), then the 32 row-slot sums are added in order.
__global__ void __launch_bounds__(1024) reduce_partials_kernel(const float* __restrict__ partial, float* __restrict__ out,
                                                               int nblocks, int width, int accumulate) {
  __shared__ float red[32][33];
  const int x = threadIdx.x, y = threadIdx.y;
  const int j = blockIdx.x * 32 + x;
  float acc = 0.0f;
  if (j < width)
    for (int b = y; b < nblocks; b += 32) acc += partial[static_cast<int64_t>(b) * width + j];
  red[y][x] = acc;
  __syncthreads();
  if (y == 0 && j < width) {
    float tot = 0.0f;
#pragma unroll
    for (int r = 0; r < 32; ++r) tot += red[r][x];
    out[j] = accumulate ? out[j] + tot : tot;
  }
}

// ------------------------------------------------------------------------------------------------
// K1: token + position + token-type embedding gather, LayerNorm.  One warp per token.
// ------------------------------------------------------------------------------------------------
template <int CPL>
__global__ void __launch_bounds__(ROW_THREADS) embed_ln_kernel(const a4r_embed_args a) {
  const int H = static_cast<int>(a.H);
  const int L = static_cast<int>(a.L);
  const int nchunks = H >> 3;
  const int lane = threadIdx.x & 31;
  const int64_t warps_per_block = ROW_THREADS / 32;
  const int64_t total = a.N * a.L;
  const float invH = 1.0f / static_cast<float>(H);
  const __nv_bfloat16* word = static_cast<const __nv_bfloat16*>(a.word_emb);
  const __nv_bfloat16* pos = static_cast<const __nv_bfloat16*>(a.pos_emb);
  const __nv_bfloat16* type0 = static_cast<const __nv_bfloat16*>(a.type_emb);
  const __nv_bfloat16* prompt = static_cast<const __nv_bfloat16*>(a.prompt);
  __nv_bfloat16* out = static_cast<__nv_bfloat16*>(a.out);
  __nv_bfloat16* z_out = static_cast<__nv_bfloat16*>(a.z_out);
  for (int64_t tok = blockIdx.x * warps_per_block + (threadIdx.x >> 5); tok < total;
       tok += gridDim.x * warps_per_block) {
    const int64_t n = tok / L;
    const int t = static_cast<int>(tok % L);
    const int64_t* ids = a.ids + n * a.ld_ids;
    int64_t id = ids[t];
    int64_t pid = t + a.pos_offset;
    if (a.roberta_pad_id >= 0) {
      // RoBERTa: position = pad + (#non-pad tokens up to and including t) for non-pad tokens, pad otherwise
      const int64_t mine = lane < L ? ids[lane] : a.roberta_pad_id;
      const uint32_t nonpad = __ballot_sync(0xffffffffu, mine != a.roberta_pad_id);
      const uint32_t upto = t >= 31 ? 0xffffffffu : ((2u << t) - 1u);
      pid = (id != a.roberta_pad_id) ? a.roberta_pad_id + __popc(nonpad & upto) : a.roberta_pad_id;
    }
    const bool use_prompt = prompt != nullptr && t < a.n_prompt;
    const __nv_bfloat16* wrow = use_prompt ? prompt + static_cast<int64_t>(t) * H : word + id * H;
    float v[CPL][8];
    float s = 0.0f;
#pragma unroll
    for (int c = 0; c < CPL; ++c) {
      const int ch = lane + c * 32;
      if (ch < nchunks) {
        float w[8], p[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(wrow + ch * 8)), w);
        unpack8(__ldg(reinterpret_cast<const uint4*>(pos + pid * H + ch * 8)), p);
#pragma unroll
        for (int e = 0; e < 8; ++e) v[c][e] = w[e] + p[e];
        if (type0 != nullptr) {
          float ty[8];
          unpack8(__ldg(reinterpret_cast<const uint4*>(type0 + ch * 8)), ty);
#pragma unroll
          for (int e = 0; e < 8; ++e) v[c][e] += ty[e];
        }
        if (z_out != nullptr) {
          const uint4 zz = pack8(v[c]);
          st_na_v4(z_out + tok * H + ch * 8, zz);
          unpack8(zz, v[c]);
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) s += v[c][e];
      }
    }
    const float mean = warp_sum(s) * invH;
    float q = 0.0f;
#pragma unroll
    for (int c = 0; c < CPL; ++c) {
      const int ch = lane + c * 32;
      if (ch < nchunks) {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float d = v[c][e] - mean;
          q += d * d;
        }
      }
    }
    const float rstd = rsqrtf(warp_sum(q) * invH + a.eps);
#pragma unroll
    for (int c = 0; c < CPL; ++c) {
      const int ch = lane + c * 32;
      if (ch < nchunks) {
        const float4 g0 = __ldg(reinterpret_cast<const float4*>(a.gamma + ch * 8));
        const float4 g1 = __ldg(reinterpret_cast<const float4*>(a.gamma + ch * 8 + 4));
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(a.beta + ch * 8));
        const float4 b1 = __ldg(reinterpret_cast<const float4*>(a.beta + ch * 8 + 4));
        const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
        const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
        float o[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = (v[c][e] - mean) * rstd * gg[e] + bb[e];
        st_na_v4(out + tok * H + ch * 8, pack8(o));
      }
    }
    if (lane == 0) {
      if (a.mean_out != nullptr) a.mean_out[tok] = mean;
      if (a.rstd_out != nullptr) a.rstd_out[tok] = rstd;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// elementwise activation gradient: out = dy * act'(u)   (u = saved pre-activation for GELU / tanh-GELU, output for
// ReLU / LeakyReLU), and the stand-alone forward for the activations that have no GEMM-epilogue mode
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float gelu_tanh(float x) {
  const float t = tanhf(0.7978845608028654f * (x + 0.044715f * x * x * x));
  return 0.5f * x * (1.0f + t);
}
__device__ __forceinline__ float gelu_tanh_grad(float x) {
  const float t = tanhf(0.7978845608028654f * (x + 0.044715f * x * x * x));
  return 0.5f * (1.0f + t) + 0.5f * x * (1.0f - t * t) * 0.7978845608028654f * (1.0f + 3.0f * 0.044715f * x * x);
}
__device__ __forceinline__ float act_grad(int kind, float dy, float u) {
  switch (kind) {
    case 0: return dy * gelu_erf_grad(u);
    case 1: return u > 0.0f ? dy : 0.0f;
    case 2: return u > 0.0f ? dy : 0.01f * dy;          // nn.LeakyReLU() default slope; sign(out) == sign(in)
    default: return dy * gelu_tanh_grad(u);
  }
}
__device__ __forceinline__ float act_value(int kind, float u) {
  switch (kind) {
    case 0: return gelu_erf(u);
    case 1: return fmaxf(u, 0.0f);
    case 2: return u > 0.0f ? u : 0.01f * u;
    default: return gelu_tanh(u);
  }
}

__global__ void act_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ u,
                               __nv_bfloat16* __restrict__ out, int64_t nvec, int kind) {
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < nvec;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float a[8], b[8], o[8];
    unpack8(ld_nc_v4(dy + i * 8), a);
    unpack8(ld_nc_v4(u + i * 8), b);
#pragma unroll
    for (int e = 0; e < 8; ++e) o[e] = act_grad(kind, a[e], b[e]);
    st_na_v4(out + i * 8, pack8(o));
  }
}

__global__ void act_fwd_kernel(const __nv_bfloat16* __restrict__ u, __nv_bfloat16* __restrict__ out, int64_t nvec, int kind) {
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < nvec;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float b[8], o[8];
    unpack8(ld_nc_v4(u + i * 8), b);
#pragma unroll
    for (int e = 0; e < 8; ++e) o[e] = act_value(kind, b[e]);
    st_na_v4(out + i * 8, pack8(o));
  }
}

// embedding-table gradient: dst[idx[r], :] += src[r, :] (fp32 accumulation of bf16 rows); rows with idx < 0 or
// idx == skip_idx (nn.Embedding's padding_idx: its row never receives a gradient) are skipped.  One warp per source row,
// 128-bit loads, red.global.add.v4.f32 (sm_90+): 2 vector atomics per 8 elements.  The accumulation ORDER across rows
// that hit the same table row is unspecified, as it is in torch's own embedding backward on CUDA.
__global__ void scatter_add_rows_kernel(const __nv_bfloat16* __restrict__ src, int64_t ld, const int64_t* __restrict__ idx,
                                        float* __restrict__ dst, int64_t R, int H, int64_t V, int64_t skip_idx) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  for (int64_t r = warp; r < R; r += nwarps) {
    const int64_t t = idx[r];
    if (t < 0 || t >= V || t == skip_idx) continue;
    float* row = dst + t * H;
    for (int c = lane; c < H / 8; c += 32) {
      float v[8];
      unpack8(ld_nc_v4(src + r * ld + c * 8), v);
      float* o = row + c * 8;
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]) : "memory");
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o + 4), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]) : "memory");
    }
  }
}

// dropout (+ residual): out = x * mask * scale (+ res); element i uses lane (i & 3) of rng64(seed, offset + i / 4)
__global__ void dropout_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ res,
                               __nv_bfloat16* __restrict__ out, int64_t nvec, uint32_t thr16, float scale, uint64_t seed,
                               uint64_t offset) {
  seed = rng_seed(seed);
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < nvec;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float a[8], o[8];
    unpack8(ld_nc_v4(x + i * 8), a);
    const uint64_t r0 = rng64(seed, offset + static_cast<uint64_t>(i) * 2), r1 = rng64(seed, offset + static_cast<uint64_t>(i) * 2 + 1);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      o[e] = rng_keep(r0, e, thr16) ? a[e] * scale : 0.0f;
      o[4 + e] = rng_keep(r1, e, thr16) ? a[4 + e] * scale : 0.0f;
    }
    if (res != nullptr) {
      float r[8];
      unpack8(ld_nc_v4(res + i * 8), r);
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] += r[e];
    }
    st_na_v4(out + i * 8, pack8(o));
  }
}

// column sums of a bf16 [M, ld] matrix over `width` columns: out[j] (+)= sum_m x[m, j].  Two-stage.
__global__ void colsum_partial_kernel(const __nv_bfloat16* __restrict__ x, int64_t ld, int64_t M, int width,
                                      float* __restrict__ partial) {
  // block handles a strided set of rows; thread j handles 8-column chunk(s)
  const int nchunks = width >> 3;
  for (int ch = threadIdx.x; ch < nchunks; ch += blockDim.x) {
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int64_t r = blockIdx.x; r < M; r += gridDim.x) {
      float a[8];
      unpack8(ld_nc_v4(x + r * ld + ch * 8), a);
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] += a[e];
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) partial[static_cast<int64_t>(blockIdx.x) * width + ch * 8 + e] = acc[e];
  }
}

template <typename F>
int dispatch_ln(int H, F&& f) {
  // (G lanes per row, chunks per lane)
  if (H <= 64) return f(std::integral_constant<int, 8>{}, std::integral_constant<int, 1>{});
  if (H <= 128) return f(std::integral_constant<int, 16>{}, std::integral_constant<int, 1>{});
  if (H <= 256) return f(std::integral_constant<int, 32>{}, std::integral_constant<int, 1>{});
  if (H <= 512) return f(std::integral_constant<int, 32>{}, std::integral_constant<int, 2>{});
  if (H <= 768) return f(std::integral_constant<int, 32>{}, std::integral_constant<int, 3>{});
  return f(std::integral_constant<int, 32>{}, std::integral_constant<int, 4>{});
}

constexpr int LN_BWD_MAX_BLOCKS = 592;  // 4 x 148

}  // namespace

extern "C" int a4r_layernorm_fwd(const void* x, const void* res, int64_t res_rows, const float* gamma,
                                 const float* beta, float eps, void* y, void* z_out, float* mean, float* rstd,
                                 int64_t M, int64_t H, a4r_stream_t stream_) {
  A4R_CHECK_ARG(x && gamma && beta && y, "layernorm_fwd: NULL pointer");
  A4R_CHECK_ARG(H >= 8 && H <= 1024 && H % 8 == 0, "layernorm: H must be a multiple of 8 in [8,1024] (got %lld)",
                (long long)H);
  A4R_CHECK_ARG(M >= 0, "layernorm: bad M");
  A4R_CHECK_ARG(a4r_aligned16(x) && a4r_aligned16(y) && a4r_aligned16(gamma) && a4r_aligned16(beta) &&
                    a4r_aligned16(res) && a4r_aligned16(z_out),
                "layernorm: pointers must be 16B aligned");
  if (res != nullptr) A4R_CHECK_ARG(res_rows > 0, "layernorm: res_rows must be > 0 when res is given");
  int rc = a4r_device_check();
  if (rc != A4R_OK) return rc;
  if (M == 0) return A4R_OK;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  return dispatch_ln(static_cast<int>(H), [&](auto G, auto CPL) -> int {
    const int64_t rows_per_block = ROW_THREADS / G.value;
    int64_t blocks = (M + rows_per_block - 1) / rows_per_block;
    const int64_t cap = static_cast<int64_t>(a4r_num_sms()) * 16;
    if (blocks > cap) blocks = cap;
    ln_fwd_kernel<G.value, CPL.value><<<static_cast<int>(blocks), ROW_THREADS, 0, stream>>>(
        static_cast<const __nv_bfloat16*>(x), static_cast<const __nv_bfloat16*>(res), res_rows, gamma, beta, eps,
        static_cast<__nv_bfloat16*>(y), static_cast<__nv_bfloat16*>(z_out), mean, rstd, M, static_cast<int>(H));
    A4R_LAUNCH_OK();
    a4r_count_launch(1);
    return A4R_OK;
  });
}

extern "C" size_t a4r_layernorm_bwd_workspace_bytes(int64_t H) {
  return static_cast<size_t>(LN_BWD_MAX_BLOCKS) * 2 * static_cast<size_t>(H) * sizeof(float);
}

static int layernorm_bwd_impl(const void* dy, const void* z, const float* mean, const float* rstd,
                              const float* gamma, void* dz, float* dgamma, float* dbeta, int32_t accumulate,
                              void* workspace, size_t workspace_bytes, int64_t M, int64_t H, void* dz_masked,
                              float dropout_p, uint64_t dropout_seed, uint64_t dropout_offset, const void* dskip,
                              a4r_stream_t stream_) {
  A4R_CHECK_ARG(dskip == nullptr || (a4r_aligned16(dskip) && dz_masked == nullptr && dgamma == nullptr && dbeta == nullptr),
                "layernorm_bwd_add: dskip must be 16B aligned and excludes dz_masked / dgamma / dbeta");
  A4R_CHECK_ARG(dropout_p >= 0.0f && dropout_p < 1.0f && a4r_aligned16(dz_masked), "layernorm_bwd: bad dropout args");
  const uint32_t thr16 = static_cast<uint32_t>(dropout_p * 65536.0f + 0.5f);
  const float dscale = 65536.0f / static_cast<float>(65536u - thr16);
  A4R_CHECK_ARG(dy && z && mean && rstd && gamma && dz, "layernorm_bwd: NULL pointer");
  A4R_CHECK_ARG(H >= 8 && H <= 1024 && H % 8 == 0, "layernorm: H must be a multiple of 8 in [8,1024]");
  A4R_CHECK_ARG(a4r_aligned16(dy) && a4r_aligned16(z) && a4r_aligned16(dz) && a4r_aligned16(gamma),
                "layernorm_bwd: pointers must be 16B aligned");
  const bool want_wgrad = dgamma != nullptr || dbeta != nullptr;
  if (want_wgrad) {
    A4R_CHECK_ARG(dgamma && dbeta, "layernorm_bwd: dgamma and dbeta must be given together");
    if (workspace == nullptr || workspace_bytes < a4r_layernorm_bwd_workspace_bytes(H))
      return a4r_set_error(A4R_EWORKSPACE, "layernorm_bwd: workspace too small (%zu < %zu)", workspace_bytes,
                           a4r_layernorm_bwd_workspace_bytes(H));
  }
  int rc = a4r_device_check();
  if (rc != A4R_OK) return rc;
  if (M == 0) return A4R_OK;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  return dispatch_ln(static_cast<int>(H), [&](auto G, auto CPL) -> int {
    const int64_t rows_per_block = ROW_THREADS / G.value;
    int64_t blocks = (M + rows_per_block - 1) / rows_per_block;
    float* partial = want_wgrad ? static_cast<float*>(workspace) : nullptr;
    if (want_wgrad) {
      if (blocks > LN_BWD_MAX_BLOCKS) blocks = LN_BWD_MAX_BLOCKS;
      ln_bwd_kernel<G.value, CPL.value, true><<<static_cast<int>(blocks), ROW_THREADS, 0, stream>>>(
          static_cast<const __nv_bfloat16*>(dy), static_cast<const __nv_bfloat16*>(z), mean, rstd, gamma,
          static_cast<__nv_bfloat16*>(dz), partial, M, static_cast<int>(H), static_cast<__nv_bfloat16*>(dz_masked), thr16,
          dscale, dropout_seed, dropout_offset);
    } else {
      const int64_t cap = static_cast<int64_t>(a4r_num_sms()) * 12;
      if (blocks > cap) blocks = cap;
      if (dskip != nullptr)
        ln_bwd_kernel<G.value, CPL.value, false, true><<<static_cast<int>(blocks), ROW_THREADS, 0, stream>>>(
            static_cast<const __nv_bfloat16*>(dy), static_cast<const __nv_bfloat16*>(z), mean, rstd, gamma,
            static_cast<__nv_bfloat16*>(dz), partial, M, static_cast<int>(H), nullptr, thr16, dscale, dropout_seed,
            dropout_offset, static_cast<const __nv_bfloat16*>(dskip));
      else
      ln_bwd_kernel<G.value, CPL.value, false><<<static_cast<int>(blocks), ROW_THREADS, 0, stream>>>(
          static_cast<const __nv_bfloat16*>(dy), static_cast<const __nv_bfloat16*>(z), mean, rstd, gamma,
          static_cast<__nv_bfloat16*>(dz), partial, M, static_cast<int>(H), static_cast<__nv_bfloat16*>(dz_masked), thr16,
          dscale, dropout_seed, dropout_offset);
    }
    A4R_LAUNCH_OK();
    a4r_count_launch(1);
    if (want_wgrad) {
      ln_reduce_kernel<<<(2 * static_cast<int>(H) + 255) / 256, 256, 0, stream>>>(
          partial, dgamma, dbeta, static_cast<int>(blocks), static_cast<int>(H), accumulate);
      A4R_LAUNCH_OK();
      a4r_count_launch(1);
    }
    return A4R_OK;
  });
}

extern "C" int a4r_layernorm_bwd(const void* dy, const void* z, const float* mean, const float* rstd,
                                 const float* gamma, void* dz, float* dgamma, float* dbeta, int32_t accumulate,
                                 void* workspace, size_t workspace_bytes, int64_t M, int64_t H, void* dz_masked,
                                 float dropout_p, uint64_t dropout_seed, uint64_t dropout_offset,
                                 a4r_stream_t stream_) {
  return layernorm_bwd_impl(dy, z, mean, rstd, gamma, dz, dgamma, dbeta, accumulate, workspace, workspace_bytes, M, H,
                            dz_masked, dropout_p, dropout_seed, dropout_offset, nullptr, stream_);
}

extern "C" int a4r_layernorm_bwd_add(const void* dy, const void* z, const float* mean, const float* rstd,
                                     const float* gamma, const void* dskip, void* dz, int64_t M, int64_t H,
                                     a4r_stream_t stream_) {
  A4R_CHECK_ARG(dskip != nullptr, "layernorm_bwd_add: dskip is NULL");
  return layernorm_bwd_impl(dy, z, mean, rstd, gamma, dz, nullptr, nullptr, 0, nullptr, 0, M, H, nullptr, 0.0f, 0, 0, dskip,
                            stream_);
}

extern "C" int a4r_embed_ln_fwd(const a4r_embed_args* a, a4r_stream_t stream_) {
  A4R_CHECK_ARG(a != nullptr, "embed_ln: args is NULL");
  A4R_CHECK_ARG(a->ids && a->word_emb && a->pos_emb && a->gamma && a->beta && a->out, "embed_ln: NULL pointer");
  A4R_CHECK_ARG(a->H >= 8 && a->H <= 1024 && a->H % 8 == 0, "embed_ln: H must be a multiple of 8 in [8,1024]");
  A4R_CHECK_ARG(a->N >= 0 && a->L >= 1 && a->ld_ids >= a->L, "embed_ln: bad N/L/ld_ids");
  A4R_CHECK_ARG(a->roberta_pad_id < 0 || a->L <= 32, "embed_ln: RoBERTa position rule supports L <= 32");
  A4R_CHECK_ARG(a->n_prompt >= 0 && a->n_prompt <= a->L, "embed_ln: bad n_prompt");
  A4R_CHECK_ARG(a4r_aligned16(a->word_emb) && a4r_aligned16(a->pos_emb) && a4r_aligned16(a->type_emb) &&
                    a4r_aligned16(a->prompt) && a4r_aligned16(a->out) && a4r_aligned16(a->z_out) &&
                    a4r_aligned16(a->gamma) && a4r_aligned16(a->beta),
                "embed_ln: pointers must be 16B aligned");
  int rc = a4r_device_check();
  if (rc != A4R_OK) return rc;
  if (a->N == 0) return A4R_OK;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int64_t total = a->N * a->L;
  int64_t blocks = (total + 7) / 8;
  const int64_t cap = static_cast<int64_t>(a4r_num_sms()) * 16;
  if (blocks > cap) blocks = cap;
  const int cpl = static_cast<int>((a->H / 8 + 31) / 32);
  switch (cpl) {
    case 1: embed_ln_kernel<1><<<static_cast<int>(blocks), ROW_THREADS, 0, stream>>>(*a); break;
    case 2: embed_ln_kernel<2><<<static_cast<int>(blocks), ROW_THREADS, 0, stream>>>(*a); break;
    case 3: embed_ln_kernel<3><<<static_cast<int>(blocks), ROW_THREADS, 0, stream>>>(*a); break;
    default: embed_ln_kernel<4><<<static_cast<int>(blocks), ROW_THREADS, 0, stream>>>(*a); break;
  }
  A4R_LAUNCH_OK();
  a4r_count_launch(1);
  return A4R_OK;
}

extern "C" int a4r_act_bwd(const void* dy, const void* u, void* out, int64_t n, int32_t kind, a4r_stream_t stream_) {
  A4R_CHECK_ARG(dy && u && out, "act_bwd: NULL pointer");
  A4R_CHECK_ARG(n >= 0 && n % 8 == 0, "act_bwd: n must be a multiple of 8");
  A4R_CHECK_ARG(kind >= 0 && kind <= 3, "act_bwd: kind must be 0 (gelu) 1 (relu) 2 (leaky_relu) 3 (gelu_new)");
  A4R_CHECK_ARG(a4r_aligned16(dy) && a4r_aligned16(u) && a4r_aligned16(out), "act_bwd: pointers must be 16B aligned");
  int rc = a4r_device_check();
  if (rc != A4R_OK) return rc;
  if (n == 0) return A4R_OK;
  const int64_t nvec = n / 8;
  int64_t blocks = (nvec + 255) / 256;
  const int64_t cap = static_cast<int64_t>(a4r_num_sms()) * 16;
  if (blocks > cap) blocks = cap;
  act_bwd_kernel<<<static_cast<int>(blocks), 256, 0, static_cast<cudaStream_t>(stream_)>>>(
      static_cast<const __nv_bfloat16*>(dy), static_cast<const __nv_bfloat16*>(u), static_cast<__nv_bfloat16*>(out),
      nvec, kind);
  A4R_LAUNCH_OK();
  a4r_count_launch(1);
  return A4R_OK;
}

// ---- weight caches: fp32 master -> bf16 copy and / or bf16 transpose in ONE pass --------------------------------------------
namespace {
// 32 x 32 tiles through shared memory (+1 padding): both outputs are written with contiguous rows
__global__ void __launch_bounds__(256) cast_transpose_kernel(const float* __restrict__ src, int64_t ld_src, __nv_bfloat16* __restrict__ dst,
                                                             __nv_bfloat16* __restrict__ dst_t, int64_t rows, int64_t cols) {
  __shared__ float tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;      // 32 x 8 threads
  const int64_t tiles_c = (cols + 31) / 32, tiles = ((rows + 31) / 32) * tiles_c;
  for (int64_t t = blockIdx.x; t < tiles; t += gridDim.x) {
    const int64_t r0 = (t / tiles_c) * 32, c0 = (t % tiles_c) * 32;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int64_t r = r0 + ty + 8 * i, c = c0 + tx;
      const float v = (r < rows && c < cols) ? src[r * ld_src + c] : 0.0f;
      tile[ty + 8 * i][tx] = v;
      if (dst != nullptr && r < rows && c < cols) dst[r * cols + c] = __float2bfloat16(v);
    }
    __syncthreads();
    if (dst_t != nullptr) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int64_t c = c0 + ty + 8 * i, r = r0 + tx;         // row of the transpose = column of the source
        if (c < cols && r < rows) dst_t[c * rows + r] = __float2bfloat16(tile[tx][ty + 8 * i]);
      }
    }
    __syncthreads();
  }
}
}  // namespace

extern "C" int a4r_cast_transpose_f32_bf16(const float* src, int64_t ld_src, void* dst, void* dst_t, int64_t rows, int64_t cols,
                                           a4r_stream_t stream_) {
  A4R_CHECK_ARG(rows >= 0 && cols >= 0 && ld_src >= cols, "cast_transpose: bad shape");
  A4R_CHECK_ARG(dst != nullptr || dst_t != nullptr, "cast_transpose: no output requested");
  if (rows == 0 || cols == 0) return A4R_OK;
  A4R_CHECK_ARG(src != nullptr, "cast_transpose: src is NULL");
  int rc = a4r_device_check();
  if (rc != A4R_OK) return rc;
  const int64_t tiles = ((rows + 31) / 32) * ((cols + 31) / 32);
  const int64_t cap = static_cast<int64_t>(a4r_num_sms()) * 8;
  cast_transpose_kernel<<<static_cast<int>(tiles < cap ? tiles : cap), 256, 0, static_cast<cudaStream_t>(stream_)>>>(
      src, ld_src, static_cast<__nv_bfloat16*>(dst), static_cast<__nv_bfloat16*>(dst_t), rows, cols);
  A4R_LAUNCH_OK();
  a4r_count_launch(1);
  return A4R_OK;
}

extern "C" int a4r_act_fwd(const void* u, void* out, int64_t n, int32_t kind, a4r_stream_t stream_) {
  A4R_CHECK_ARG(u && out, "act_fwd: NULL pointer");
  A4R_CHECK_ARG(n >= 0 && n % 8 == 0, "act_fwd: n must be a multiple of 8");
  A4R_CHECK_ARG(kind >= 0 && kind <= 3, "act_fwd: kind must be 0 (gelu) 1 (relu) 2 (leaky_relu) 3 (gelu_new)");
  A4R_CHECK_ARG(a4r_aligned16(u) && a4r_aligned16(out), "act_fwd: pointers must be 16B aligned");
  int rc = a4r_device_check();
  if (rc != A4R_OK) return rc;
  if (n == 0) return A4R_OK;
  const int64_t nvec = n / 8;
  int64_t blocks = (nvec + 255) / 256;
  const int64_t cap = static_cast<int64_t>(a4r_num_sms()) * 16;
  if (blocks > cap) blocks = cap;
  act_fwd_kernel<<<static_cast<int>(blocks), 256, 0, static_cast<cudaStream_t>(stream_)>>>(
      static_cast<const __nv_bfloat16*>(u), static_cast<__nv_bfloat16*>(out), nvec, kind);
  A4R_LAUNCH_OK();
  a4r_count_launch(1);
  return A4R_OK;
}

constexpr int COLSUM_BLOCKS = 592;

extern "C" size_t a4r_colsum_workspace_bytes(int64_t width) {
  return static_cast<size_t>(COLSUM_BLOCKS) * static_cast<size_t>(width) * sizeof(float);
}

extern "C" int a4r_colsum(const void* x, int64_t ld, int64_t M, int64_t width, float* out, int32_t accumulate,
                          void* workspace, size_t workspace_bytes, a4r_stream_t stream_) {
  A4R_CHECK_ARG(x && out, "colsum: NULL pointer");
  A4R_CHECK_ARG(width >= 8 && width % 8 == 0 && ld >= width && ld % 8 == 0, "colsum: bad width/ld");
  A4R_CHECK_ARG(a4r_aligned16(x), "colsum: x must be 16B aligned");
  if (workspace == nullptr || workspace_bytes < a4r_colsum_workspace_bytes(width))
    return a4r_set_error(A4R_EWORKSPACE, "colsum: workspace too small");
  int rc = a4r_device_check();
  if (rc != A4R_OK) return rc;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int blocks = static_cast<int>(M < COLSUM_BLOCKS ? (M > 0 ? M : 1) : COLSUM_BLOCKS);
  int threads = static_cast<int>(width / 8);
  threads = threads < 32 ? 32 : (threads > 256 ? 256 : ((threads + 31) / 32) * 32);
  colsum_partial_kernel<<<blocks, threads, 0, stream>>>(static_cast<const __nv_bfloat16*>(x), ld, M,
                                                        static_cast<int>(width), static_cast<float*>(workspace));
  A4R_LAUNCH_OK();
  reduce_partials_kernel<<<(static_cast<int>(width) + 31) / 32, dim3(32, 32), 0, stream>>>(
      static_cast<const float*>(workspace), out, blocks, static_cast<int>(width), accumulate);
  A4R_LAUNCH_OK();
  a4r_count_launch(2);
  return A4R_OK;
}

extern "C" int a4r_dropout(const void* x, const void* res, void* out, int64_t n, float p, uint64_t seed, uint64_t offset,
                           a4r_stream_t stream_) {
  A4R_CHECK_ARG(x && out, "dropout: NULL pointer");
  A4R_CHECK_ARG(n >= 0 && n % 8 == 0, "dropout: n must be a multiple of 8");
  A4R_CHECK_ARG(p >= 0.0f && p < 1.0f, "dropout: p must be in [0, 1)");
  A4R_CHECK_ARG(a4r_aligned16(x) && a4r_aligned16(res) && a4r_aligned16(out), "dropout: pointers must be 16B aligned");
  int rc = a4r_device_check();
  if (rc != A4R_OK) return rc;
  if (n == 0) return A4R_OK;
  const uint32_t thr16 = static_cast<uint32_t>(p * 65536.0f + 0.5f);
  const float scale = 65536.0f / static_cast<float>(65536u - thr16);
  const int64_t nvec = n / 8;
  int64_t blocks = (nvec + 255) / 256;
  const int64_t cap = static_cast<int64_t>(a4r_num_sms()) * 16;
  if (blocks > cap) blocks = cap;
  dropout_kernel<<<static_cast<int>(blocks), 256, 0, static_cast<cudaStream_t>(stream_)>>>(
      static_cast<const __nv_bfloat16*>(x), static_cast<const __nv_bfloat16*>(res), static_cast<__nv_bfloat16*>(out), nvec,
      thr16, scale, seed, offset);
  A4R_LAUNCH_OK();
  a4r_count_launch(1);
  return A4R_OK;
}

extern "C" int a4r_scatter_add_rows(const void* src, int64_t ld, const int64_t* idx, float* dst, int64_t R, int64_t H,
                                    int64_t V, int64_t skip_idx, a4r_stream_t stream_) {
  A4R_CHECK_ARG(R >= 0 && H >= 8 && H % 8 == 0 && V > 0 && ld >= H && ld % 8 == 0, "scatter_add_rows: bad sizes");
  if (R == 0) return A4R_OK;
  A4R_CHECK_ARG(src && idx && dst, "scatter_add_rows: NULL pointer");
  A4R_CHECK_ARG(a4r_aligned16(src) && a4r_aligned16(dst), "scatter_add_rows: src and dst must be 16B aligned");
  int rc = a4r_device_check();
  if (rc != A4R_OK) return rc;
  int64_t blocks = (R + 7) / 8;
  const int64_t cap = static_cast<int64_t>(a4r_num_sms()) * 8;
  if (blocks > cap) blocks = cap;
  scatter_add_rows_kernel<<<static_cast<int>(blocks), 256, 0, static_cast<cudaStream_t>(stream_)>>>(
      static_cast<const __nv_bfloat16*>(src), ld, idx, dst, R, static_cast<int>(H), V, skip_idx);
  A4R_LAUNCH_OK();
  a4r_count_launch(1);
  return A4R_OK;
}
