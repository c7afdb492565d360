// api.cu — error plumbing, device check and process-wide counters of libadapter4rec_sm100.so
#include <stdarg.h>
#include <atomic>

#include "a4r_common.cuh"

namespace {
thread_local char g_err[512] = {0};
std::atomic<int64_t> g_launches{0};
}  // namespace

int a4r_set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

void a4r_count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int a4r_num_sms() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
  }
  return sms;
}

namespace {
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
PFN_encodeTiled encode_fn() {
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
}  // namespace

int a4r_make_tmap_bf16(CUtensorMap* m, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
  PFN_encodeTiled fn = encode_fn();
  if (fn == nullptr) return a4r_set_error(A4R_ECUDA, "cuTensorMapEncodeTiled entry point not found");
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 2};
  cuuint32_t box[2] = {64u, static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return a4r_set_error(A4R_ECUDA, "cuTensorMapEncodeTiled failed (%d) rows=%lld cols=%lld ld=%lld", (int)r,
                         (long long)rows, (long long)cols, (long long)ld);
  return A4R_OK;
}

extern "C" int a4r_version(void) { return 100; /* 0.1.0 */ }

extern "C" const char* a4r_last_error_string(void) { return g_err; }

extern "C" int64_t a4r_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

extern "C" int a4r_device_check(void) {
  static int cached_dev = -1;
  static int cached_rc = A4R_OK;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return a4r_set_error(A4R_ECUDA, "cudaGetDevice failed: %s", cudaGetErrorString(e));
  if (dev == cached_dev) {
    if (cached_rc != A4R_OK) a4r_set_error(cached_rc, "device %d is not compute capability 10.x (sm_100a only)", dev);
    return cached_rc;
  }
  int major = 0;
  e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (e != cudaSuccess) return a4r_set_error(A4R_ECUDA, "cudaDeviceGetAttribute failed: %s", cudaGetErrorString(e));
  cached_dev = dev;
  cached_rc = (major == 10) ? A4R_OK : A4R_EARCH;
  if (cached_rc != A4R_OK)
    return a4r_set_error(A4R_EARCH, "device %d has compute capability %d.x; this library is sm_100a only", dev, major);
  return A4R_OK;
}
