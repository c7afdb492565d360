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
