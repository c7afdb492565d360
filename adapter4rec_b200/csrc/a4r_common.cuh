// adapter4rec_b200 — shared device/host helpers for the sm_100a kernels.
// Everything here is internal to libadapter4rec_sm100.so; the public surface is include/adapter4rec.h.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/adapter4rec.h"

// ---------------------------------------------------------------------------------------------
// host-side error plumbing (thread-local last-error string, negative return codes)
// ---------------------------------------------------------------------------------------------
int a4r_set_error(int code, const char* fmt, ...);

#define A4R_CHECK_ARG(cond, ...)                                   \
  do {                                                              \
    if (!(cond)) return a4r_set_error(A4R_EINVAL, __VA_ARGS__);     \
  } while (0)

#define A4R_CUDA_OK(expr)                                                                  \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess)                                                                  \
      return a4r_set_error(A4R_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                           __FILE__, __LINE__);                                             \
  } while (0)

#define A4R_LAUNCH_OK()                                                                   \
  do {                                                                                     \
    cudaError_t _e = cudaGetLastError();                                                   \
    if (_e != cudaSuccess)                                                                 \
      return a4r_set_error(A4R_ECUDA, "kernel launch failed: %s (%s:%d)",                  \
                           cudaGetErrorString(_e), __FILE__, __LINE__);                    \
  } while (0)

int a4r_num_sms();  // cached per process (current device)
void a4r_count_launch(int n);

// TMA descriptor of a bf16 row-major [rows, cols] matrix (leading dimension ld, elements): box = 64 columns x box_rows,
// SWIZZLE_128B, out-of-bounds elements read as zero.  Defined in api.cu.
int a4r_make_tmap_bf16(CUtensorMap* m, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows);

static inline bool a4r_aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ---------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------
#ifdef __CUDACC__

#define A4R_DEVICE __device__ __forceinline__

A4R_DEVICE uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

A4R_DEVICE float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
A4R_DEVICE float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

A4R_DEVICE float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }
A4R_DEVICE float gelu_erf_grad(float x) {
  const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
  const float pdf = 0.39894228040143267794f * __expf(-0.5f * x * x);
  return cdf + x * pdf;
}

// erf-GELU for the GEMM / adapter epilogues, which are bound by the FMA pipe (measured: the packed FFMA2 form issues at half the
// rate of FFMA, so an epilogue pays ~2 pipe cycles per packed operation and the degree-8 erfc polynomial of round 1 — 14 packed
// operations per pair — made the FFN1 epilogue slower than the MMAs of its tile).  Phi(x) is approximated by
//     Phi(x) ~= 1/2 + 1/2 tanh(x (c0 + c1 x^2 + c2 x^4)),   x^2 clamped at 64,
// a three-coefficient minimax fit against the exact erf form (max |x dPhi| = 2.5e-5 over the real line, fitted offline) evaluated
// with ONE MUFU.TANH (relative error 2^-11, i.e. |d gelu| <= 2.4e-4 |x|: an eighth of a bf16 ulp of the result) and 6 packed
// operations per pair.  The density term of GELU' keeps the exact exponential (one more MUFU.EX2).
A4R_DEVICE float ex2_approx(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
A4R_DEVICE float tanh_approx(float x) {
  float r;
  asm("tanh.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
constexpr float kGeluC0 = 0.797507884f, kGeluC1 = 0.0370056460f, kGeluC2 = -3.51516783e-4f;
A4R_DEVICE float gelu_cdf(float x) {
  const float x2 = fminf(x * x, 64.0f);
  const float u = x * fmaf(fmaf(x2, kGeluC2, kGeluC1), x2, kGeluC0);
  return fmaf(0.5f, tanh_approx(u), 0.5f);
}
A4R_DEVICE float gelu_fast(float x) { return x * gelu_cdf(x); }
A4R_DEVICE float gelu_grad_fast(float x) {
  const float e = ex2_approx(-0.72134752044448170368f * x * x);          // exp(-x^2 / 2)
  return fmaf(x * 0.39894228040143267794f, e, gelu_cdf(x));
}
// Two elements at a time on the packed fp32x2 pipe of sm_100 (FFMA2 / FMUL2): half the issue slots of the scalar form.
A4R_DEVICE float2 splat2(float c) { return make_float2(c, c); }
A4R_DEVICE float2 gelu_cdf2(float2 x) {
  float2 x2 = __fmul2_rn(x, x);
  x2 = make_float2(fminf(x2.x, 64.0f), fminf(x2.y, 64.0f));
  float2 q = __ffma2_rn(x2, splat2(kGeluC2), splat2(kGeluC1));
  q = __ffma2_rn(q, x2, splat2(kGeluC0));
  const float2 u = __fmul2_rn(q, x);
  return __ffma2_rn(make_float2(tanh_approx(u.x), tanh_approx(u.y)), splat2(0.5f), splat2(0.5f));
}
A4R_DEVICE float2 gelu_fast2(float2 x) { return __fmul2_rn(x, gelu_cdf2(x)); }
A4R_DEVICE float2 gelu_grad_fast2(float2 x) {
  const float2 t = __fmul2_rn(__fmul2_rn(x, splat2(-0.72134752044448170368f)), x);
  const float2 e = make_float2(ex2_approx(t.x), ex2_approx(t.y));
  return __ffma2_rn(__fmul2_rn(x, splat2(0.39894228040143267794f)), e, gelu_cdf2(x));
}

// 256-bit global access (sm_100: LDG.256 / STG.256): one full 32-byte sector per thread per instruction
A4R_DEVICE void ld_nc_v8(const void* p, uint32_t (&x)[8]) {
  asm volatile("ld.global.nc.L1::no_allocate.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(x[0]), "=r"(x[1]), "=r"(x[2]), "=r"(x[3]), "=r"(x[4]), "=r"(x[5]), "=r"(x[6]), "=r"(x[7])
               : "l"(p));
}
A4R_DEVICE void st_na_v8(void* p, const uint32_t (&x)[8]) {
  asm volatile("st.global.L1::no_allocate.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(x[0]), "r"(x[1]),
               "r"(x[2]), "r"(x[3]), "r"(x[4]), "r"(x[5]), "r"(x[6]), "r"(x[7])
               : "memory");
}

A4R_DEVICE uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
A4R_DEVICE float2 unpack_bf16x2(uint32_t v) {
  __nv_bfloat162 h = *reinterpret_cast<__nv_bfloat162*>(&v);
  return __bfloat1622float2(h);
}

// 128-bit streaming global access (read-once / write-once tensors)
A4R_DEVICE uint4 ld_nc_v4(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
A4R_DEVICE void st_na_v4(void* p, const uint4& v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y),
               "r"(v.z), "r"(v.w)
               : "memory");
}

// ---- mbarrier ---------------------------------------------------------------------------------
A4R_DEVICE void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
A4R_DEVICE void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
A4R_DEVICE void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
A4R_DEVICE void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
A4R_DEVICE void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
A4R_DEVICE bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
A4R_DEVICE void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}

// ---- TMA ----------------------------------------------------------------------------------------
A4R_DEVICE void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
// 2-D tiled load: coordinates are (c0 = innermost/K element index, c1 = row index)
A4R_DEVICE void tma_load_2d(const CUtensorMap* m, void* smem_dst, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

// ---- tcgen05 / TMEM -----------------------------------------------------------------------------
A4R_DEVICE void tmem_alloc(uint32_t* smem_slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)),
               "r"(ncols)
               : "memory");
}
A4R_DEVICE void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
A4R_DEVICE void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
A4R_DEVICE void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
A4R_DEVICE void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem]; bf16 x bf16 -> fp32, issued by ONE thread.
A4R_DEVICE void umma_bf16_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier once all previously issued tcgen05.mma of this thread have completed
A4R_DEVICE void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
A4R_DEVICE void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// 32 lanes x 32 consecutive fp32 columns: thread i of the warp receives row (lane base + i).
A4R_DEVICE void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}

// 32 lanes x 16 consecutive fp32 columns
A4R_DEVICE void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

// explicit shared-space 128-bit access through a 32-bit shared address (keeps ptxas from emitting generic LD/ST)
A4R_DEVICE void sts_v4(uint32_t saddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(saddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
A4R_DEVICE float4 lds_v4f(uint32_t saddr) {
  float4 r;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(saddr) : "memory");
  return r;
}
A4R_DEVICE uint4 lds_v4(uint32_t saddr) {
  uint4 r;
  asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(saddr) : "memory");
  return r;
}
// bf16x2 -> two fp32 with two ALU ops (shift / mask): the epilogues that stream residuals are instruction-bound
A4R_DEVICE float2 bf16x2_to_f2(uint32_t v) { return make_float2(__uint_as_float(v << 16), __uint_as_float(v & 0xFFFF0000u)); }

// ---- CTA-pair (cta_group::2) variants: two SMs of a TPC cooperate on one 256-row tile ----------------------------------
A4R_DEVICE uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
A4R_DEVICE void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same shared-memory offset in CTA `rank` of the cluster
A4R_DEVICE uint32_t mapa_u32(uint32_t smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
A4R_DEVICE void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// Remote arrive WITHOUT cluster-scope release.  The .release.cluster form compiles to MEMBAR.ALL.GPU + ERRBAR, i.e. it
// waits until every global store of the warp is visible device-wide — the epilogue warps would drain their output
// stores before they may hand the TMEM stage back.  Handing a TMEM stage back orders only tcgen05 accesses, which
// tcgen05.fence::before_thread_sync / after_thread_sync around the arrive / wait already do.
A4R_DEVICE void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
A4R_DEVICE void tmem_alloc2(uint32_t* smem_slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)), "r"(ncols)
               : "memory");
}
A4R_DEVICE void tmem_relinquish2() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
A4R_DEVICE void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// issued by ONE thread of the LEADER CTA: D (both CTAs' TMEM, 128 lanes each) (+)= A (128 rows from each CTA) * B (N/2 rows
// from each CTA)
A4R_DEVICE void umma_bf16_ss_2cta(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the mbarrier at the same shared offset in every CTA of `cta_mask` once the issued MMAs have completed
A4R_DEVICE void umma_commit_2cta(uint64_t* bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(cta_mask)
               : "memory");
}
// TMA load into LOCAL shared memory whose completion bytes are credited to the mbarrier at `bar_cluster_addr`
// (the leader CTA's full barrier)
A4R_DEVICE void tma_load_2d_2cta(const CUtensorMap* m, void* smem_dst, uint32_t bar_cluster_addr, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}

// Shared-memory matrix descriptor, K-major operand, 128-byte swizzle (what a TMA box of 64 bf16 x rows
// with CU_TENSOR_MAP_SWIZZLE_128B produces): 8-row groups are 1024 B apart (SBO), LBO unused.
A4R_DEVICE uint64_t umma_desc_k_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);  // start address, bits [0,14)
  d |= static_cast<uint64_t>(0) << 16;                      // leading byte offset (ignored for SW128 K-major)
  d |= static_cast<uint64_t>(1024 >> 4) << 32;              // stride byte offset, bits [32,46)
  d |= static_cast<uint64_t>(1) << 46;                      // descriptor version (Blackwell)
  d |= static_cast<uint64_t>(2) << 61;                      // layout type: SWIZZLE_128B
  return d;
}
// kind::f16 instruction descriptor: bf16 A/B (both K-major), fp32 accumulate, M x N tile.
A4R_DEVICE constexpr uint32_t umma_idesc_bf16(uint32_t m, uint32_t n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((n >> 3) << 17) | ((m >> 4) << 24);
}

// ---- counter-based dropout RNG ------------------------------------------------------------------
// splitmix64 of (seed, counter): 64 random bits = four 16-bit lanes, i.e. one call decides 4 consecutive elements.
// An element is KEPT iff its 16-bit lane >= thr16 = round(p * 65536); kept values are scaled by 65536 / (65536 - thr16)
// (the exact reciprocal of the realised keep probability, so the mask is unbiased).  Stateless: the backward pass
// regenerates the mask from the same (seed, counter) instead of storing it.
A4R_DEVICE uint64_t rng64(uint64_t seed, uint64_t counter) {
  uint64_t z = counter * 0x9E3779B97F4A7C15ull + seed;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
// A seed argument with bit 63 set (A4R_SEED_INDIRECT, adapter4rec.h) is the device address of the 64-bit seed: a step
// recorded in a CUDA graph bakes its kernel arguments, so the seed that must change between replays lives in device
// memory.  Resolved once per thread, next to the draws, only on paths that have dropout switched on.
A4R_DEVICE uint64_t rng_seed(uint64_t s) {
  if (static_cast<int64_t>(s) < 0)
    s = __ldg(reinterpret_cast<const unsigned long long*>(s & 0x7FFFFFFFFFFFFFFFull));
  return s;
}
A4R_DEVICE bool rng_keep(uint64_t bits, int lane4, uint32_t thr16) {
  return ((static_cast<uint32_t>(bits >> (16 * lane4))) & 0xFFFFu) >= thr16;
}

// Non-blocking probe of an mbarrier phase.  (mbarrier.try_wait may SUSPEND the thread for a system-dependent time when the phase is
// not complete: in a loop that polls two queues, a probe of the queue that cannot advance then hides the other queue's event.)
A4R_DEVICE bool mbar_test_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// ---- tcgen05 issue from WARP-UNIFORM code ----------------------------------------------------------------------------------
// `if (lane == 0) { tcgen05.mma ... }` puts the instruction in a divergent region: ptxas then wraps EVERY UTCHMMA in an
// ELECT / BRA.U.ANY loop and rebuilds its uniform-register operands inside it — about ten dependent uniform-datapath
// instructions, measured at 77-87 cycles per MMA on the issuing thread (clock64 traces of the attention forward: 13 P V
// instructions of N = 64 took 1.0-1.1 k cycles to ISSUE, 32 cycles each to execute).  Executed by all 32 lanes of a converged warp
// with the election inside the PTX, the same instructions compile to back-to-back UTCHMMAs.  The callers keep every operand
// warp-uniform (loop counters, kernel parameters, shared-memory addresses) and make poll results uniform with a vote.
A4R_DEVICE void umma_bf16_ss_elect(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// expect_tx + one / two 2-D TMA loads on the same mbarrier, from a converged warp (one elected lane arms and issues)
A4R_DEVICE void tma_load_2d_elect(const CUtensorMap* m, void* smem_dst, uint64_t* bar, int c0, int c1, uint32_t expect_bytes) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "@q mbarrier.arrive.expect_tx.shared::cta.b64 _, [%2], %5;\n\t"
      "@q cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n\t}"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(expect_bytes)
      : "memory");
}
A4R_DEVICE void tma_load_2d_elect_noarm(const CUtensorMap* m, void* smem_dst, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "@q cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n\t}"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
// the same store issued by a converged warp with the election inside the PTX (no ELECT / BRA.U.ANY loop around the UTMASTG); the
// elected lane of a full, converged warp is always the same one, so its bulk groups can be waited on the same way
A4R_DEVICE void tma_store_2d_commit_elect(const CUtensorMap* m, uint32_t smem_src, int c0, int c1) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "@q cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];\n\t"
      "@q cp.async.bulk.commit_group;\n\t}"
      ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_src), "r"(c0), "r"(c1)
      : "memory");
}
A4R_DEVICE void bulk_wait_read1_elect() {
  asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t@q cp.async.bulk.wait_group.read 1;\n\t}" ::: "memory");
}
A4R_DEVICE void bulk_wait0_elect() {
  asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t@q cp.async.bulk.wait_group 0;\n\t}" ::: "memory");
}
A4R_DEVICE void mbar_expect_tx_elect(uint64_t* bar, uint32_t bytes) {
  asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t@q mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n\t}"
               ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
A4R_DEVICE void umma_bf16_ss_2cta_elect(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@q tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
A4R_DEVICE void umma_commit_2cta_elect(uint64_t* bar, uint16_t cta_mask) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "@q tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n\t}"
      ::"r"(smem_u32(bar)), "h"(cta_mask)
      : "memory");
}
A4R_DEVICE void tma_load_2d_2cta_elect(const CUtensorMap* m, void* smem_dst, uint32_t bar_cluster_addr, int c0, int c1) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "@q cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n\t}"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
// L2 eviction-priority policies for streams whose reuse distance is known (createpolicy; 64-bit opaque operand of .L2::cache_hint)
A4R_DEVICE uint64_t l2_evict_last_policy() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
A4R_DEVICE uint64_t l2_evict_normal_policy() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
A4R_DEVICE uint64_t l2_evict_first_policy() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
A4R_DEVICE void tma_load_2d_elect_hint(const CUtensorMap* m, void* smem_dst, uint64_t* bar, int c0, int c1, uint32_t expect_bytes, uint64_t pol) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "@q mbarrier.arrive.expect_tx.shared::cta.b64 _, [%2], %5;\n\t"
      "@q cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %6;\n\t}"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(expect_bytes), "l"(pol)
      : "memory");
}
A4R_DEVICE void tma_load_2d_elect_noarm_hint(const CUtensorMap* m, void* smem_dst, uint64_t* bar, int c0, int c1, uint64_t pol) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "@q cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;\n\t}"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "l"(pol)
      : "memory");
}
A4R_DEVICE void tma_store_2d_commit_elect_hint(const CUtensorMap* m, uint32_t smem_src, int c0, int c1, uint64_t pol) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "@q cp.async.bulk.tensor.2d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%2, %3}], [%1], %4;\n\t"
      "@q cp.async.bulk.commit_group;\n\t}"
      ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_src), "r"(c0), "r"(c1), "l"(pol)
      : "memory");
}
A4R_DEVICE void umma_commit_elect(uint64_t* bar) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(smem_u32(bar))
      : "memory");
}
A4R_DEVICE bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

#endif  // __CUDACC__
