// Library plumbing: version, thread-local error string, device check, launch counter.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace xtag {

static thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int num_sms() {
  static int cached = 0;
  if (cached == 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
      cached = n;
    else
      return 148;
  }
  return cached;
}

}  // namespace xtag

extern "C" {

int xtag_version(void) { return XTAG_ABI_VERSION; }

const char* xtag_last_error(void) { return xtag::g_err; }

uint64_t xtag_launch_count(void) { return xtag::g_launches.load(std::memory_order_relaxed); }

int xtag_device_check(void) {
  int dev = 0, major = 0;
  XTAG_CUDA(cudaGetDevice(&dev));
  XTAG_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  XTAG_REQUIRE(major == 10, XTAG_ERR_CUDA,
               "libxtag_b200 is built for sm_100a only; device %d has compute capability %d.x", dev, major);
  return XTAG_OK;
}

}  // extern "C"
