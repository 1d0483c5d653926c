// Error handling, arch gate and launch accounting of the C ABI.
#include "common.cuh"
#include <atomic>
#include <string.h>

namespace mpa {
static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int check_arch() {
  static thread_local int cached_dev = -1;
  static thread_local int cached_rc = MPA_ERR_ARCH;
  int dev = -1;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    set_error("no CUDA device: %s (libmpa has no CPU fallback)", cudaGetErrorString(e));
    return MPA_ERR_ARCH;
  }
  if (dev == cached_dev) return cached_rc;
  int major = 0, minor = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  cached_dev = dev;
  if (major != 10) {
    set_error("device %d is sm_%d%d; libmpa is built for sm_100a only and has no fallback path", dev, major, minor);
    cached_rc = MPA_ERR_ARCH;
  } else {
    cached_rc = MPA_OK;
  }
  return cached_rc;
}
}  // namespace mpa

extern "C" {
int mpa_version(void) { return 100; }
const char* mpa_last_error(void) { return mpa::g_err; }
int mpa_device_check(void) { return mpa::check_arch(); }
long long mpa_launch_count(void) { return mpa::g_launches.load(); }
}
