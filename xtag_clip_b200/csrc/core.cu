// Library plumbing: version, thread-local error string, device check, launch counter.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <vector>

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
  static std::atomic<int> cached[64];   // per device (zero-initialised)
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  int n = cached[dev & 63].load(std::memory_order_relaxed);
  if (n == 0) {
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) return 148;
    cached[dev & 63].store(n, std::memory_order_relaxed);
  }
  return n;
}

// runtime tuning bits of the tcgen05 kernels (see EpiParams::tune in clip_tc.cu); initialised once from the
// environment variable XTAG_TC_TUNE, overridable with xtag_set_tune()
static std::atomic<int> g_tune{-1};
// defaults: bit 11 = single-pass K4 backward (1.4-2x faster than the two-kernel path); n-slab of 32 tiles (8192 columns
// = 16 MB of the B operand stay L2-resident while the m groups stream past: -1.8 % step time in the sustained power
// state, profiles/r2_sustained_ab.json); CTA-pair kernels on (bit 24 clear)
static constexpr int kDefaultTune = 0x200800;
int tc_tune() {
  int t = g_tune.load(std::memory_order_relaxed);
  if (t < 0) {
    const char* e = getenv("XTAG_TC_TUNE");
    t = e ? (int)strtol(e, nullptr, 0) & 0x7fffffff : kDefaultTune;
    g_tune.store(t, std::memory_order_relaxed);
  }
  return t;
}

// spin budget of the bounded device-side waits (tc_ptx.cuh); generation 0 = the environment / default value
static std::atomic<long long> g_spin_ms{-1};
static std::atomic<int> g_spin_gen{0};
unsigned long long spin_timeout_ns() {
  long long ms = g_spin_ms.load(std::memory_order_relaxed);
  if (ms < 0) {
    const char* e = getenv("XTAG_SPIN_TIMEOUT_MS");
    ms = e ? strtoll(e, nullptr, 0) : 1800LL * 1000LL;
    if (ms <= 0) ms = 1800LL * 1000LL;
    g_spin_ms.store(ms, std::memory_order_relaxed);
  }
  return (unsigned long long)ms * 1000000ull;
}
int spin_timeout_gen() { return g_spin_gen.load(std::memory_order_relaxed); }

struct ProfRec {
  int tag;
  double work;
  cudaEvent_t e0, e1;
};
static std::mutex g_prof_mu;
static bool g_prof_on = false;
static std::vector<ProfRec> g_prof;

ProfScope::ProfScope(int tag, double work, cudaStream_t s) : idx(-1), st(s) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  if (!g_prof_on) return;
  ProfRec r;
  r.tag = tag;
  r.work = work;
  if (cudaEventCreate(&r.e0) != cudaSuccess || cudaEventCreate(&r.e1) != cudaSuccess) return;
  cudaEventRecord(r.e0, s);
  g_prof.push_back(r);
  idx = (int)g_prof.size() - 1;
}
ProfScope::~ProfScope() {
  if (idx < 0) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  if (idx < (int)g_prof.size()) cudaEventRecord(g_prof[idx].e1, st);
}

}  // namespace xtag

extern "C" {

int xtag_prof_enable(int on) {
  std::lock_guard<std::mutex> lk(xtag::g_prof_mu);
  for (auto& r : xtag::g_prof) {
    cudaEventDestroy(r.e0);
    cudaEventDestroy(r.e1);
  }
  xtag::g_prof.clear();
  xtag::g_prof_on = on != 0;
  return XTAG_OK;
}

int xtag_prof_read(int* tags, float* ms, double* work, int cap) {
  std::lock_guard<std::mutex> lk(xtag::g_prof_mu);
  int n = 0;
  for (auto& r : xtag::g_prof) {
    if (n >= cap) break;
    if (cudaEventSynchronize(r.e1) != cudaSuccess) break;
    float t = 0.f;
    if (cudaEventElapsedTime(&t, r.e0, r.e1) != cudaSuccess) break;
    tags[n] = r.tag;
    ms[n] = t;
    work[n] = r.work;
    ++n;
  }
  return n;
}


int xtag_set_tune(int bits) {
  const int old = xtag::tc_tune();
  xtag::g_tune.store(bits & 0x7fffffff, std::memory_order_relaxed);
  return old;
}

int xtag_get_tune(void) { return xtag::tc_tune(); }

int xtag_set_spin_timeout_ms(long long ms) {
  XTAG_REQUIRE(ms > 0, XTAG_ERR_INVALID, "xtag_set_spin_timeout_ms: the budget must be positive (got %lld)", ms);
  xtag::g_spin_ms.store(ms, std::memory_order_relaxed);
  xtag::g_spin_gen.fetch_add(1, std::memory_order_relaxed);
  return XTAG_OK;
}

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
