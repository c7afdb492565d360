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
constexpr int NUM_EPI_GROUPS = 2;    // 2 x 4 epilogue warps; group g owns 32-column chunks c with c % 2 == g
constexpr int NUM_THREADS = 128 + 128 * NUM_EPI_GROUPS;

template <int BN>
struct Cfg {
  static constexpr int kStageBytesA = BM * BK * 2;
  static constexpr int kStageBytesB = BN * BK * 2;
  static constexpr int kStageBytes = kStageBytesA + kStageBytesB;
  static constexpr int kStages = BN == 256 ? 4 : (BN == 128 ? 6 : 8);
  static constexpr int kTmemCols = 2 * BN;  // 128 / 256 / 512: power of two >= 32
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
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
};

template <int BN>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmB2,
               const GemmParams p) {
  using C = Cfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + C::kStages * C::kStageBytes);
  uint64_t* empty_bar = full_bar + C::kStages;
  uint64_t* tmem_full_bar = empty_bar + C::kStages;
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_tiles = p.tiles_m * p.tiles_n;
  const int nk = p.nk1 + p.nk2;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (p.nk2 > 0) {
      tma_prefetch_desc(&tmA2);
      tma_prefetch_desc(&tmB2);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < C::kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full_bar[s], 1);
      mbar_init(&tmem_empty_bar[s], 4 * NUM_EPI_GROUPS);  // one arrive per epilogue warp
    }
    mbar_fence_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, C::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ============================== TMA producer ==============================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = (tile / p.tiles_n) * BM;
        const int n0 = (tile % p.tiles_n) * BN;
        for (int kb = 0; kb < nk; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * C::kStageBytes;
          uint8_t* sb = sa + C::kStageBytesA;
          mbar_expect_tx(&full_bar[stage], C::kStageBytes);
          if (kb < p.nk1) {
            tma_load_2d(&tmA, sa, &full_bar[stage], kb * BK, m0);
            tma_load_2d(&tmB, sb, &full_bar[stage], kb * BK, n0);
          } else {
            tma_load_2d(&tmA2, sa, &full_bar[stage], (kb - p.nk1) * BK, m0);
            tma_load_2d(&tmB2, sb, &full_bar[stage], (kb - p.nk1) * BK, n0);
          }
          if (++stage == C::kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ============================== MMA issuer ==============================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&tmem_empty_bar[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(as * BN);
        for (int kb = 0; kb < nk; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * C::kStageBytes);
          const uint64_t adesc = umma_desc_k_sw128(sa);
          const uint64_t bdesc = umma_desc_k_sw128(sa + C::kStageBytesA);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            // advance the start address by k*16 elements (32 B) inside the 128 B swizzle row
            umma_bf16_ss(tmem_d, adesc + static_cast<uint64_t>(k * 2), bdesc + static_cast<uint64_t>(k * 2), idesc,
                         (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);  // frees the smem slot when these MMAs retire
          if (++stage == C::kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&tmem_full_bar[as]);  // accumulator ready for the epilogue
        if (++as == 2) {
          as = 0;
          aphase ^= 1;
        }
      }
    }
  } else if (warp >= 4) {
    // ============================== epilogue ==============================
    const int ew = warp - 4;
    const int quad = warp & 3;  // TMEM lane quadrant this warp may read
    const int group = ew >> 2;
    int as = 0;
    uint32_t aphase = 0;
    const bool out_f32 = p.out_f32 != 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m0 = (tile / p.tiles_n) * BM;
      const int n0 = (tile % p.tiles_n) * BN;
      const int row = m0 + quad * 32 + lane;
      const bool row_ok = row < p.M;
      mbar_wait(&tmem_full_bar[as], aphase);
      tc_fence_after();
#pragma unroll 1
      for (int c = group; c < BN / 32; c += NUM_EPI_GROUPS) {
        const int col0 = n0 + c * 32;
        if (col0 >= p.N) break;  // warp-uniform
        uint32_t acc[32];
        tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) +
                               static_cast<uint32_t>(as * BN + c * 32),
                           acc);
        tmem_ld_wait();
        if (row_ok) {
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            const int col = col0 + j;
            if (col < p.N) {
              float v[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) v[e] = p.alpha * __uint_as_float(acc[j + e]);
              if (p.bias != nullptr) {
                const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + col));
                const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + col + 4));
                v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
                v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
              }
              const int64_t r64 = row;
              switch (p.epilogue) {
                case A4R_EPI_LINEAR: {
                  if (p.residual != nullptr) {
                    const uint4 r = ld_nc_v4(reinterpret_cast<const __nv_bfloat16*>(p.residual) + r64 * p.ldr + col);
                    float2 f;
                    f = unpack_bf16x2(r.x); v[0] += f.x; v[1] += f.y;
                    f = unpack_bf16x2(r.y); v[2] += f.x; v[3] += f.y;
                    f = unpack_bf16x2(r.z); v[4] += f.x; v[5] += f.y;
                    f = unpack_bf16x2(r.w); v[6] += f.x; v[7] += f.y;
                  }
                  if (p.residual2 != nullptr) {
                    const uint4 r = ld_nc_v4(reinterpret_cast<const __nv_bfloat16*>(p.residual2) + r64 * p.ldr2 + col);
                    float2 f;
                    f = unpack_bf16x2(r.x); v[0] += f.x; v[1] += f.y;
                    f = unpack_bf16x2(r.y); v[2] += f.x; v[3] += f.y;
                    f = unpack_bf16x2(r.z); v[4] += f.x; v[5] += f.y;
                    f = unpack_bf16x2(r.w); v[6] += f.x; v[7] += f.y;
                  }
                } break;
                case A4R_EPI_GELU: {
                  if (p.aux != nullptr) {
                    uint4 o;
                    o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]);
                    o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]);
                    st_na_v4(reinterpret_cast<__nv_bfloat16*>(p.aux) + r64 * p.ldaux + col, o);
                  }
#pragma unroll
                  for (int e = 0; e < 8; ++e) v[e] = gelu_erf(v[e]);
                } break;
                case A4R_EPI_RELU: {
#pragma unroll
                  for (int e = 0; e < 8; ++e) v[e] = fmaxf(v[e], 0.0f);
                } break;
                case A4R_EPI_DGELU: {
                  const uint4 r = ld_nc_v4(reinterpret_cast<const __nv_bfloat16*>(p.aux) + r64 * p.ldaux + col);
                  float u[8];
                  float2 f;
                  f = unpack_bf16x2(r.x); u[0] = f.x; u[1] = f.y;
                  f = unpack_bf16x2(r.y); u[2] = f.x; u[3] = f.y;
                  f = unpack_bf16x2(r.z); u[4] = f.x; u[5] = f.y;
                  f = unpack_bf16x2(r.w); u[6] = f.x; u[7] = f.y;
#pragma unroll
                  for (int e = 0; e < 8; ++e) v[e] *= gelu_erf_grad(u[e]);
                } break;
                case A4R_EPI_DRELU: {
                  const uint4 r = ld_nc_v4(reinterpret_cast<const __nv_bfloat16*>(p.aux) + r64 * p.ldaux + col);
                  float u[8];
                  float2 f;
                  f = unpack_bf16x2(r.x); u[0] = f.x; u[1] = f.y;
                  f = unpack_bf16x2(r.y); u[2] = f.x; u[3] = f.y;
                  f = unpack_bf16x2(r.z); u[4] = f.x; u[5] = f.y;
                  f = unpack_bf16x2(r.w); u[6] = f.x; u[7] = f.y;
#pragma unroll
                  for (int e = 0; e < 8; ++e) v[e] = u[e] > 0.0f ? v[e] : 0.0f;
                } break;
                default: break;
              }
              if (out_f32) {
                float* cp = reinterpret_cast<float*>(p.C) + r64 * p.ldc + col;
                st_na_v4(cp, make_uint4(__float_as_uint(v[0]), __float_as_uint(v[1]), __float_as_uint(v[2]),
                                        __float_as_uint(v[3])));
                st_na_v4(cp + 4, make_uint4(__float_as_uint(v[4]), __float_as_uint(v[5]), __float_as_uint(v[6]),
                                            __float_as_uint(v[7])));
              } else {
                uint4 o;
                o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]);
                o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]);
                st_na_v4(reinterpret_cast<__nv_bfloat16*>(p.C) + r64 * p.ldc + col, o);
              }
            }
          }
        }
      }
      // all of this warp's TMEM reads for this stage are complete (wait::ld above): release it
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty_bar[as]);
      if (++as == 2) {
        as = 0;
        aphase ^= 1;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 2) tmem_dealloc(tmem_base, C::kTmemCols);
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

template <int BN>
int launch_gemm(const a4r_gemm_args* a, cudaStream_t stream) {
  using C = Cfg<BN>;
  CUtensorMap tmA, tmB, tmA2, tmB2;
  int rc;
  if ((rc = make_tmap(&tmA, a->A, a->M, a->K, a->lda, BM)) != A4R_OK) return rc;
  if ((rc = make_tmap(&tmB, a->B, a->N, a->K, a->ldb, BN)) != A4R_OK) return rc;
  if (a->K2 > 0) {
    if ((rc = make_tmap(&tmA2, a->A2, a->M, a->K2, a->lda2, BM)) != A4R_OK) return rc;
    if ((rc = make_tmap(&tmB2, a->B2, a->N, a->K2, a->ldb2, BN)) != A4R_OK) return rc;
  } else {
    tmA2 = tmA;
    tmB2 = tmB;
  }
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
  p.tiles_m = static_cast<int>((a->M + BM - 1) / BM);
  p.tiles_n = static_cast<int>((a->N + BN - 1) / BN);
  p.alpha = a->alpha;
  p.epilogue = a->epilogue;
  p.out_f32 = a->out_f32;

  static bool attr_set = false;
  if (!attr_set) {
    A4R_CUDA_OK(cudaFuncSetAttribute(gemm_tn_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes));
    attr_set = true;
  }
  const int64_t tiles = static_cast<int64_t>(p.tiles_m) * p.tiles_n;
  const int grid = static_cast<int>(tiles < a4r_num_sms() ? tiles : a4r_num_sms());
  gemm_tn_kernel<BN><<<grid, NUM_THREADS, C::kSmemBytes, stream>>>(tmA, tmB, tmA2, tmB2, p);
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
    A4R_CHECK_ARG(a->aux != nullptr, "gemm: DGELU/DRELU need aux");
  if (a->aux) A4R_CHECK_ARG(a4r_aligned16(a->aux) && a->ldaux % 8 == 0 && a->ldaux >= a->N, "gemm: bad aux/ldaux");
  if (a->residual)
    A4R_CHECK_ARG(a4r_aligned16(a->residual) && a->ldr % 8 == 0 && a->ldr >= a->N, "gemm: bad residual/ldr");
  if (a->residual2)
    A4R_CHECK_ARG(a4r_aligned16(a->residual2) && a->ldr2 % 8 == 0 && a->ldr2 >= a->N, "gemm: bad residual2/ldr2");
  if (a->bias) A4R_CHECK_ARG(a4r_aligned16(a->bias), "gemm: bias must be 16B aligned");
  int rc = a4r_device_check();
  if (rc != A4R_OK) return rc;
  if (a->M == 0) return A4R_OK;

  int bn = a->block_n;
  if (bn == 0) bn = a->N > 128 ? 256 : (a->N > 64 ? 128 : 64);
  switch (bn) {
    case 256: return launch_gemm<256>(a, stream);
    case 128: return launch_gemm<128>(a, stream);
    case 64: return launch_gemm<64>(a, stream);
    default: return a4r_set_error(A4R_EINVAL, "gemm: block_n must be 0, 64, 128 or 256 (got %d)", bn);
  }
}
