// K6: the "dense output" block of the tag head's BERT layers, fused:
//     y = LayerNorm( dropout(x) + resid ) * gamma + beta
// (reference src/open_clip/tagging_heads/bert.py:281-292 BertSelfOutput and :359-370 BertOutput: dense -> dropout ->
//  LayerNorm(hidden + input); the dense GEMM -- bias included -- stays a library call and hands x over in bf16.)
//
// Stock PyTorch under bf16 autocast runs this as dropout (2 kernels) + add + fp32 LayerNorm + casts forward and
// masked-scale + LayerNorm input-grad + the gamma/beta reduction (386 us per call at [45056, 768]) + adds backward:
// ~0.8 ms per LayerNorm, four of them per tag-head step.  Here: ONE pass forward (read x, resid; write z = the
// pre-normalisation sum, y) and ONE pass backward (read dy, z; write dx, d resid; per-CTA partial sums of d gamma /
// d beta, finished by a small reduction), both HBM-bound.
//
// One warp per row; H = 256 * NV (NV <= 4), lane l owns the 8-element vectors at columns 256 v + 8 l.  Statistics in
// fp32 over the unrounded sum; z and y are stored as bf16.  The dropout keep-mask is Philox4x32-7 keyed by (seed,
// offset, element index / 4), regenerated in the backward.  resid may have fewer rows than x (resid_rows): row r reads
// resid[r % resid_rows] -- layer 0 adds the same 44 label embeddings to every sample (model.py:342).
#include "common.cuh"
#include "philox.cuh"

namespace xtag {

__device__ __forceinline__ void ld8_bf16(const __nv_bfloat16* p, float (&v)[8]) {
  const uint4 r = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __bfloat1622float2(h[i]);
    v[2 * i] = f.x; v[2 * i + 1] = f.y;
  }
}
__device__ __forceinline__ void st8_bf16(__nv_bfloat16* p, const float (&v)[8]) {
  uint4 r;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = r;
}
__device__ __forceinline__ void ld8_f32(const float* p, float (&v)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
template <typename T> __device__ __forceinline__ void ld8(const T* p, float (&v)[8]);
template <> __device__ __forceinline__ void ld8<float>(const float* p, float (&v)[8]) { ld8_f32(p, v); }
template <> __device__ __forceinline__ void ld8<__nv_bfloat16>(const __nv_bfloat16* p, float (&v)[8]) { ld8_bf16(p, v); }

// keep bits of the 8 elements starting at flat index e0 (a multiple of 8): two Philox blocks of 4 elements
__device__ __forceinline__ uint32_t keep8(uint64_t seed, uint64_t offset, uint64_t e0, uint32_t thr) {
  uint32_t r[4], bits = 0;
  philox4x32(seed, e0 >> 2, offset, r);
#pragma unroll
  for (int k = 0; k < 4; ++k) bits |= ((r[k] >> 8) >= thr ? 1u : 0u) << k;
  philox4x32(seed, (e0 >> 2) + 1, offset, r);
#pragma unroll
  for (int k = 0; k < 4; ++k) bits |= ((r[k] >> 8) >= thr ? 1u : 0u) << (4 + k);
  return bits;
}

template <int NV, typename TR>
__global__ void __launch_bounds__(256, 2) ln_res_fwd_kernel(const __nv_bfloat16* __restrict__ x, const TR* __restrict__ resid,
                                                         int resid_rows, const float* __restrict__ gamma,
                                                         const float* __restrict__ beta, __nv_bfloat16* __restrict__ z,
                                                         __nv_bfloat16* __restrict__ y, float* __restrict__ mean,
                                                         float* __restrict__ rstd, int rows, float eps, float p_drop,
                                                         uint64_t seed, uint64_t offset) {
  constexpr int H = 256 * NV;
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  const uint32_t thr = philox_drop_threshold(p_drop);
  const float ks = p_drop > 0.f ? 1.f / (1.f - p_drop) : 1.f;
  float g[NV][8], bt[NV][8];
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    ld8_f32(gamma + v * 256 + lane * 8, g[v]);
    ld8_f32(beta + v * 256 + lane * 8, bt[v]);
  }
  for (int r = blockIdx.x * wpb + (threadIdx.x >> 5); r < rows; r += gridDim.x * wpb) {
    float a[NV][8];
    float s = 0.f;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const int c = v * 256 + lane * 8;
      float xv[8], rv[8];
      ld8_bf16(x + (size_t)r * H + c, xv);
      ld8<TR>(resid + (size_t)(r % resid_rows) * H + c, rv);
      uint32_t keep = 0xffu;
      if (p_drop > 0.f) keep = keep8(seed, offset, (uint64_t)r * H + c, thr);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        a[v][i] = (((keep >> i) & 1u) ? xv[i] * ks : 0.f) + rv[i];
        s += a[v][i];
      }
    }
    s = warp_sum(s);
    const float mu = s * (1.f / H);
    float q = 0.f;
#pragma unroll
    for (int v = 0; v < NV; ++v)
#pragma unroll
      for (int i = 0; i < 8; ++i) q = fmaf(a[v][i] - mu, a[v][i] - mu, q);
    q = warp_sum(q);
    const float rs = rsqrtf(q * (1.f / H) + eps);
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const int c = v * 256 + lane * 8;
      float o[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = fmaf((a[v][i] - mu) * rs, g[v][i], bt[v][i]);
      st8_bf16(z + (size_t)r * H + c, a[v]);
      st8_bf16(y + (size_t)r * H + c, o);
    }
    if (lane == 0) {
      mean[r] = mu;
      rstd[r] = rs;
    }
  }
}

// dz = rstd * ( dy*gamma - mean_H(dy*gamma) - xhat * mean_H(dy*gamma*xhat) ),  xhat = (z - mean) * rstd
// dx = dz * keep / (1 - p);  d resid = dz;  d gamma = sum_rows dy * xhat;  d beta = sum_rows dy
// part: [gridDim.x][2][H] fp32 partial sums of (d gamma, d beta), one slab per CTA.
template <int NV, typename TG>
__global__ void __launch_bounds__(256, 2) ln_res_bwd_kernel(const TG* __restrict__ dy, const __nv_bfloat16* __restrict__ z,
                                                         const float* __restrict__ mean, const float* __restrict__ rstd,
                                                         const float* __restrict__ gamma, __nv_bfloat16* __restrict__ dx,
                                                         __nv_bfloat16* __restrict__ dres, float* __restrict__ part,
                                                         int rows, float p_drop, uint64_t seed, uint64_t offset) {
  constexpr int H = 256 * NV;
  __shared__ float red[8][256 * NV];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const uint32_t thr = philox_drop_threshold(p_drop);
  const float ks = p_drop > 0.f ? 1.f / (1.f - p_drop) : 1.f;
  float g[NV][8], dg[NV][8], db[NV][8];
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    ld8_f32(gamma + v * 256 + lane * 8, g[v]);
#pragma unroll
    for (int i = 0; i < 8; ++i) dg[v][i] = db[v][i] = 0.f;
  }
  for (int r = blockIdx.x * wpb + warp; r < rows; r += gridDim.x * wpb) {
    const float mu = mean[r], rs = rstd[r];
    float d[NV][8], xh[NV][8];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const int c = v * 256 + lane * 8;
      float zv[8];
      ld8<TG>(dy + (size_t)r * H + c, d[v]);
      ld8_bf16(z + (size_t)r * H + c, zv);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        xh[v][i] = (zv[i] - mu) * rs;
        dg[v][i] = fmaf(d[v][i], xh[v][i], dg[v][i]);
        db[v][i] += d[v][i];
        d[v][i] *= g[v][i];                       // dy * gamma
        s1 += d[v][i];
        s2 = fmaf(d[v][i], xh[v][i], s2);
      }
    }
    s1 = warp_sum(s1) * (1.f / H);
    s2 = warp_sum(s2) * (1.f / H);
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const int c = v * 256 + lane * 8;
      float dz[8], dxv[8];
      uint32_t keep = 0xffu;
      if (p_drop > 0.f) keep = keep8(seed, offset, (uint64_t)r * H + c, thr);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        dz[i] = rs * (d[v][i] - s1 - xh[v][i] * s2);
        dxv[i] = ((keep >> i) & 1u) ? dz[i] * ks : 0.f;
      }
      st8_bf16(dres + (size_t)r * H + c, dz);
      st8_bf16(dx + (size_t)r * H + c, dxv);
    }
  }
  // CTA partial of d gamma / d beta: 8 warps -> shared memory -> one slab per CTA (d gamma first, then d beta)
#pragma unroll
  for (int which = 0; which < 2; ++which) {
#pragma unroll
    for (int v = 0; v < NV; ++v)
#pragma unroll
      for (int i = 0; i < 8; ++i) red[warp][v * 256 + lane * 8 + i] = which ? db[v][i] : dg[v][i];
    __syncthreads();
    for (int col = threadIdx.x; col < H; col += blockDim.x) {
      float t = 0.f;
      for (int w = 0; w < wpb; ++w) t += red[w][col];
      part[((size_t)blockIdx.x * 2 + which) * H + col] = t;
    }
    __syncthreads();
  }
}

// out[which][col] = sum_p part[p][which][col]   (fixed order: deterministic)
__global__ void __launch_bounds__(256) ln_param_reduce_kernel(const float* __restrict__ part, int P, int H2,
                                                              float* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= H2) return;
  float t[4] = {0.f, 0.f, 0.f, 0.f};                 // four independent chains: the loads overlap
  int p = 0;
  for (; p + 4 <= P; p += 4) {
#pragma unroll
    for (int u = 0; u < 4; ++u) t[u] += part[(size_t)(p + u) * H2 + c];
  }
  for (; p < P; ++p) t[0] += part[(size_t)p * H2 + c];
  out[c] = (t[0] + t[1]) + (t[2] + t[3]);
}

static int ln_grid(int rows) {
  int blocks = (rows + 7) / 8;
  const int cap = num_sms() * 2;                     // 2 CTAs of 8 warps per SM (launch bounds), one wave
  return blocks > cap ? cap : (blocks < 1 ? 1 : blocks);
}

}  // namespace xtag

using namespace xtag;

extern "C" size_t xtag_ln_res_bwd_ws_bytes(int rows, int H) {
  if (rows <= 0 || H <= 0) return 0;
  return (size_t)ln_grid(rows) * 2 * (size_t)H * sizeof(float) + 256;
}

extern "C" int xtag_ln_res_fwd(const void* x, const void* resid, int resid_dtype, int resid_rows, const float* gamma,
                               const float* beta, void* z, void* y, float* mean, float* rstd, int rows, int H, float eps,
                               float dropout_p, uint64_t seed, uint64_t offset, void* stream) {
  XTAG_REQUIRE(x && resid && gamma && beta && z && y && mean && rstd && rows > 0 && resid_rows > 0, XTAG_ERR_INVALID,
               "ln_res_fwd: bad arguments");
  XTAG_REQUIRE(H % 256 == 0 && H >= 256 && H <= 1024, XTAG_ERR_UNSUPPORTED,
               "ln_res_fwd: hidden size %d not in {256, 512, 768, 1024}", H);
  XTAG_REQUIRE(resid_dtype == XTAG_F32 || resid_dtype == XTAG_BF16, XTAG_ERR_INVALID, "ln_res_fwd: bad resid dtype");
  XTAG_REQUIRE(dropout_p >= 0.f && dropout_p < 1.f, XTAG_ERR_INVALID, "ln_res_fwd: dropout_p must be in [0, 1)");
  int rc = xtag_device_check();
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = ln_grid(rows);
#define XTAG_LN_FWD(NV, TR)                                                                                            \
  ln_res_fwd_kernel<NV, TR><<<grid, 256, 0, st>>>((const __nv_bfloat16*)x, (const TR*)resid, resid_rows, gamma, beta,  \
                                                  (__nv_bfloat16*)z, (__nv_bfloat16*)y, mean, rstd, rows, eps,         \
                                                  dropout_p, seed, offset)
  const int nv = H / 256;
  if (resid_dtype == XTAG_BF16) {
    if (nv == 1) XTAG_LN_FWD(1, __nv_bfloat16); else if (nv == 2) XTAG_LN_FWD(2, __nv_bfloat16);
    else if (nv == 3) XTAG_LN_FWD(3, __nv_bfloat16); else XTAG_LN_FWD(4, __nv_bfloat16);
  } else {
    if (nv == 1) XTAG_LN_FWD(1, float); else if (nv == 2) XTAG_LN_FWD(2, float);
    else if (nv == 3) XTAG_LN_FWD(3, float); else XTAG_LN_FWD(4, float);
  }
#undef XTAG_LN_FWD
  XTAG_CHECK_LAUNCH();
  return XTAG_OK;
}

extern "C" int xtag_ln_res_bwd(const void* dy, int dy_dtype, const void* z, const float* mean, const float* rstd,
                               const float* gamma, void* dx, void* dresid, float* dgamma, float* dbeta, int rows, int H,
                               float dropout_p, uint64_t seed, uint64_t offset, void* ws, size_t ws_bytes,
                               void* stream) {
  XTAG_REQUIRE(dy && z && mean && rstd && gamma && dx && dresid && dgamma && dbeta && rows > 0, XTAG_ERR_INVALID,
               "ln_res_bwd: bad arguments");
  XTAG_REQUIRE(H % 256 == 0 && H >= 256 && H <= 1024, XTAG_ERR_UNSUPPORTED,
               "ln_res_bwd: hidden size %d not in {256, 512, 768, 1024}", H);
  XTAG_REQUIRE(dy_dtype == XTAG_F32 || dy_dtype == XTAG_BF16, XTAG_ERR_INVALID, "ln_res_bwd: bad dy dtype");
  XTAG_REQUIRE(ws && ws_bytes >= xtag_ln_res_bwd_ws_bytes(rows, H), XTAG_ERR_WORKSPACE, "ln_res_bwd: workspace %zu < %zu",
               ws_bytes, xtag_ln_res_bwd_ws_bytes(rows, H));
  XTAG_REQUIRE(dgamma + H == dbeta, XTAG_ERR_INVALID, "ln_res_bwd: dgamma and dbeta must be one [2, H] buffer");
  int rc = xtag_device_check();
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = ln_grid(rows);
  float* part = (float*)ws;
#define XTAG_LN_BWD(NV, TG)                                                                                            \
  ln_res_bwd_kernel<NV, TG><<<grid, 256, 0, st>>>((const TG*)dy, (const __nv_bfloat16*)z, mean, rstd, gamma,           \
                                                  (__nv_bfloat16*)dx, (__nv_bfloat16*)dresid, part, rows, dropout_p,   \
                                                  seed, offset)
  const int nv = H / 256;
  if (dy_dtype == XTAG_BF16) {
    if (nv == 1) XTAG_LN_BWD(1, __nv_bfloat16); else if (nv == 2) XTAG_LN_BWD(2, __nv_bfloat16);
    else if (nv == 3) XTAG_LN_BWD(3, __nv_bfloat16); else XTAG_LN_BWD(4, __nv_bfloat16);
  } else {
    if (nv == 1) XTAG_LN_BWD(1, float); else if (nv == 2) XTAG_LN_BWD(2, float);
    else if (nv == 3) XTAG_LN_BWD(3, float); else XTAG_LN_BWD(4, float);
  }
#undef XTAG_LN_BWD
  XTAG_CHECK_LAUNCH();
  ln_param_reduce_kernel<<<(2 * H + 63) / 64, 64, 0, st>>>(part, grid, 2 * H, dgamma);
  XTAG_CHECK_LAUNCH();
  return XTAG_OK;
}
