// Shared helpers for libxtag_b200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>

#include "../../include/xtag_b200.h"

namespace xtag {

// ---- error plumbing ---------------------------------------------------------------------------
void set_error(const char* fmt, ...);
extern std::atomic<uint64_t> g_launches;
inline void count_launch(int n = 1) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

#define XTAG_REQUIRE(cond, code, ...)                \
  do {                                               \
    if (!(cond)) {                                   \
      ::xtag::set_error(__VA_ARGS__);                \
      return (code);                                 \
    }                                                \
  } while (0)

#define XTAG_CUDA(expr)                                                              \
  do {                                                                               \
    cudaError_t _e = (expr);                                                         \
    if (_e != cudaSuccess) {                                                         \
      ::xtag::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),      \
                        __FILE__, __LINE__);                                         \
      return XTAG_ERR_CUDA;                                                          \
    }                                                                                \
  } while (0)

#define XTAG_CHECK_LAUNCH()                                                          \
  do {                                                                               \
    cudaError_t _e = cudaGetLastError();                                             \
    if (_e != cudaSuccess) {                                                         \
      ::xtag::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e),  \
                        __FILE__, __LINE__);                                         \
      return XTAG_ERR_CUDA;                                                          \
    }                                                                                \
    ::xtag::count_launch();                                                          \
  } while (0)

int num_sms();
int tc_tune();   // runtime tuning bits of the tcgen05 kernels (xtag_set_tune / XTAG_TC_TUNE)

// opt-in per-launch timing (CUDA events on the launching stream), see xtag_prof_enable in the header
struct ProfScope {
  int idx;
  cudaStream_t st;
  ProfScope(int tag, double work, cudaStream_t s);
  ~ProfScope();
};

// ---- small device helpers ---------------------------------------------------------------------
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// block reductions over <= 1024 threads; `red` is >= 32 floats of shared memory.
__device__ __forceinline__ float block_sum(float v, float* red) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  float t = (threadIdx.x < nw) ? red[threadIdx.x] : 0.f;
  if (wid == 0) {
    t = warp_sum(t);
    if (lane == 0) red[0] = t;
  }
  __syncthreads();
  return red[0];
}

// 2^x with the hardware approximation (MUFU.EX2); inputs are pre-multiplied by log2(e).
__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float fast_log2(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

}  // namespace xtag
