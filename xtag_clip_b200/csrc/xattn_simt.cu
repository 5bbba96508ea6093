// K4, exact fp32 path: tag-head cross-attention core for any dtype / any shape.
// Reference: BertSelfAttention.forward cross branch, src/open_clip/tagging_heads/bert.py:219-274
//   scores = q k^T / sqrt(dh) (+0 mask) ; P = softmax(scores) ; P = dropout(P) ; ctx = P v
// One CTA per (sample, head).  fp32 math throughout; the bf16 production kernel is xattn_mma.cu.
#include "common.cuh"
#include "philox.cuh"

namespace xtag {

template <typename T>
__global__ void __launch_bounds__(256) xattn_fwd_simt_kernel(const T* __restrict__ q, const T* __restrict__ k,
                                                             const T* __restrict__ v, T* __restrict__ o,
                                                             float* __restrict__ lse, int Lq, int Lk, int heads, int dh,
                                                             int ldq, int ldk, int ldv, float sm_scale, float p_drop,
                                                             uint64_t seed, uint64_t offset) {
  extern __shared__ float sm[];
  float* Qs = sm;                       // [Lq][dh+1]
  float* Ss = sm + Lq * (dh + 1);       // [Lq][Lk]
  const int bh = blockIdx.x, b = bh / heads, h = bh % heads;
  const int tid = threadIdx.x, nt = blockDim.x;
  const T* qb = q + (size_t)b * Lq * ldq + h * dh;
  const T* kb = k + (size_t)b * Lk * ldk + h * dh;
  const T* vb = v + (size_t)b * Lk * ldv + h * dh;
  for (int i = tid; i < Lq * dh; i += nt) {
    const int r = i / dh, c = i % dh;
    Qs[r * (dh + 1) + c] = to_f32(qb[(size_t)r * ldq + c]);
  }
  __syncthreads();
  // scores (log2 domain): thread per (n, q) with n fastest so a warp shares q rows from smem (broadcast)
  const float sl2 = sm_scale * kLog2e;
  for (int i = tid; i < Lq * Lk; i += nt) {
    const int n = i % Lk, r = i / Lk;
    const T* kr = kb + (size_t)n * ldk;
    float acc = 0.f;
    for (int c = 0; c < dh; ++c) acc = fmaf(Qs[r * (dh + 1) + c], to_f32(kr[c]), acc);
    Ss[r * Lk + n] = acc * sl2;
  }
  __syncthreads();
  // softmax: one warp per row
  const int lane = tid & 31, wid = tid >> 5, nw = nt >> 5;
  const float keep_scale = p_drop > 0.f ? 1.f / (1.f - p_drop) : 1.f;
  for (int r = wid; r < Lq; r += nw) {
    float mx = -INFINITY;
    for (int n = lane; n < Lk; n += 32) mx = fmaxf(mx, Ss[r * Lk + n]);
    mx = warp_max(mx);
    float l = 0.f;
    for (int n = lane; n < Lk; n += 32) l += exp2f(Ss[r * Lk + n] - mx);
    l = warp_sum(l);
    const float inv = 1.f / l;
    for (int n = lane; n < Lk; n += 32) {
      float p = exp2f(Ss[r * Lk + n] - mx) * inv;
      if (p_drop > 0.f) {
        p = philox_keep(seed, offset, (uint64_t)bh * Lq + r, n, (Lk + 3) >> 2, philox_drop_threshold(p_drop))
                ? p * keep_scale : 0.f;
      }
      Ss[r * Lk + n] = p;
    }
    if (lane == 0) lse[(size_t)bh * Lq + r] = (mx + log2f(l)) * kLn2;
  }
  __syncthreads();
  // ctx: thread per (r, c), c fastest -> coalesced v reads and o writes
  T* ob = o + (size_t)b * Lq * (heads * dh) + h * dh;
  for (int i = tid; i < Lq * dh; i += nt) {
    const int r = i / dh, c = i % dh;
    float acc = 0.f;
    for (int n = 0; n < Lk; ++n) acc = fmaf(Ss[r * Lk + n], to_f32(vb[(size_t)n * ldv + c]), acc);
    ob[(size_t)r * (heads * dh) + c] = from_f32<T>(acc);
  }
}

template <typename T>
__global__ void __launch_bounds__(256) xattn_bwd_simt_kernel(const T* __restrict__ q, const T* __restrict__ k,
                                                             const T* __restrict__ v, const T* __restrict__ o,
                                                             const T* __restrict__ d_o, const float* __restrict__ lse,
                                                             T* __restrict__ dq, T* __restrict__ dk, T* __restrict__ dv,
                                                             int Lq, int Lk, int heads, int dh, int ldq, int ldk, int ldv,
                                                             float sm_scale, float p_drop, uint64_t seed, uint64_t offset) {
  extern __shared__ float sm[];
  const int dp1 = dh + 1;
  float* Qs = sm;                      // [Lq][dh+1]
  float* dOs = Qs + Lq * dp1;          // [Lq][dh+1]
  float* Ps = dOs + Lq * dp1;          // [Lq][Lk]  dropped probabilities (P * mask / (1-p))
  float* dSs = Ps + Lq * Lk;           // [Lq][Lk]
  float* delta = dSs + Lq * Lk;        // [Lq]
  const int bh = blockIdx.x, b = bh / heads, h = bh % heads;
  const int tid = threadIdx.x, nt = blockDim.x, HD = heads * dh;
  const T* qb = q + (size_t)b * Lq * ldq + h * dh;
  const T* kb = k + (size_t)b * Lk * ldk + h * dh;
  const T* vb = v + (size_t)b * Lk * ldv + h * dh;
  const T* ob = o + (size_t)b * Lq * HD + h * dh;
  const T* dob = d_o + (size_t)b * Lq * HD + h * dh;
  for (int i = tid; i < Lq * dh; i += nt) {
    const int r = i / dh, c = i % dh;
    Qs[r * dp1 + c] = to_f32(qb[(size_t)r * ldq + c]);
    dOs[r * dp1 + c] = to_f32(dob[(size_t)r * HD + c]);
  }
  __syncthreads();
  const int lane = tid & 31, wid = tid >> 5, nw = nt >> 5;
  for (int r = wid; r < Lq; r += nw) {   // delta_r = sum_c dO[r,c] * O[r,c]
    float s = 0.f;
    for (int c = lane; c < dh; c += 32) s = fmaf(dOs[r * dp1 + c], to_f32(ob[(size_t)r * HD + c]), s);
    s = warp_sum(s);
    if (lane == 0) delta[r] = s;
  }
  __syncthreads();
  const float sl2 = sm_scale * kLog2e;
  const float keep_scale = p_drop > 0.f ? 1.f / (1.f - p_drop) : 1.f;
  for (int i = tid; i < Lq * Lk; i += nt) {
    const int n = i % Lk, r = i / Lk;
    const T* kr = kb + (size_t)n * ldk;
    const T* vr = vb + (size_t)n * ldv;
    float s = 0.f, dp = 0.f;
    for (int c = 0; c < dh; ++c) {
      s = fmaf(Qs[r * dp1 + c], to_f32(kr[c]), s);
      dp = fmaf(dOs[r * dp1 + c], to_f32(vr[c]), dp);
    }
    const float p = exp2f(s * sl2 - lse[(size_t)bh * Lq + r] * kLog2e);
    float m = 1.f;
    if (p_drop > 0.f) {
      m = philox_keep(seed, offset, (uint64_t)bh * Lq + r, n, (Lk + 3) >> 2, philox_drop_threshold(p_drop))
              ? keep_scale : 0.f;
    }
    Ps[r * Lk + n] = p * m;
    dSs[r * Lk + n] = p * (dp * m - delta[r]) * sm_scale;   // d(score before scale) folded with sm_scale
  }
  __syncthreads();
  T* dqb = dq + (size_t)b * Lq * HD + h * dh;
  for (int i = tid; i < Lq * dh; i += nt) {
    const int r = i / dh, c = i % dh;
    float acc = 0.f;
    for (int n = 0; n < Lk; ++n) acc = fmaf(dSs[r * Lk + n], to_f32(kb[(size_t)n * ldk + c]), acc);
    dqb[(size_t)r * HD + c] = from_f32<T>(acc);
  }
  T* dkb = dk + (size_t)b * Lk * HD + h * dh;
  T* dvb = dv + (size_t)b * Lk * HD + h * dh;
  for (int i = tid; i < Lk * dh; i += nt) {
    const int n = i / dh, c = i % dh;
    float ak = 0.f, av = 0.f;
    for (int r = 0; r < Lq; ++r) {
      ak = fmaf(dSs[r * Lk + n], Qs[r * dp1 + c], ak);
      av = fmaf(Ps[r * Lk + n], dOs[r * dp1 + c], av);
    }
    dkb[(size_t)n * HD + c] = from_f32<T>(ak);
    dvb[(size_t)n * HD + c] = from_f32<T>(av);
  }
}

size_t xattn_simt_fwd_smem(int Lq, int Lk, int dh) { return ((size_t)Lq * (dh + 1) + (size_t)Lq * Lk) * 4; }
size_t xattn_simt_bwd_smem(int Lq, int Lk, int dh) {
  return (2 * (size_t)Lq * (dh + 1) + 2 * (size_t)Lq * Lk + Lq) * 4;
}

template <typename T>
int xattn_simt_fwd(const void* q, const void* k, const void* v, void* o, float* lse, int b, int Lq, int Lk, int heads,
                   int dh, int ldq, int ldk, int ldv, float sm_scale, float p_drop, uint64_t seed, uint64_t offset,
                   cudaStream_t st) {
  const size_t smem = xattn_simt_fwd_smem(Lq, Lk, dh);
  XTAG_REQUIRE(smem <= 227 * 1024, XTAG_ERR_UNSUPPORTED, "xattn_fwd(simt): Lq=%d Lk=%d dh=%d needs %zu B smem", Lq, Lk, dh, smem);
  XTAG_CUDA(cudaFuncSetAttribute(xattn_fwd_simt_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  xattn_fwd_simt_kernel<T><<<b * heads, 256, smem, st>>>((const T*)q, (const T*)k, (const T*)v, (T*)o, lse, Lq, Lk, heads, dh,
                                                         ldq, ldk, ldv, sm_scale, p_drop, seed, offset);
  XTAG_CHECK_LAUNCH();
  return XTAG_OK;
}

template <typename T>
int xattn_simt_bwd(const void* q, const void* k, const void* v, const void* o, const void* d_o, const float* lse,
                   void* dq, void* dk, void* dv, int b, int Lq, int Lk, int heads, int dh, int ldq, int ldk, int ldv,
                   float sm_scale, float p_drop, uint64_t seed, uint64_t offset, cudaStream_t st) {
  const size_t smem = xattn_simt_bwd_smem(Lq, Lk, dh);
  XTAG_REQUIRE(smem <= 227 * 1024, XTAG_ERR_UNSUPPORTED, "xattn_bwd(simt): Lq=%d Lk=%d dh=%d needs %zu B smem", Lq, Lk, dh, smem);
  XTAG_CUDA(cudaFuncSetAttribute(xattn_bwd_simt_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  xattn_bwd_simt_kernel<T><<<b * heads, 256, smem, st>>>((const T*)q, (const T*)k, (const T*)v, (const T*)o, (const T*)d_o, lse,
                                                         (T*)dq, (T*)dk, (T*)dv, Lq, Lk, heads, dh, ldq, ldk, ldv, sm_scale,
                                                         p_drop, seed, offset);
  XTAG_CHECK_LAUNCH();
  return XTAG_OK;
}

template int xattn_simt_fwd<float>(const void*, const void*, const void*, void*, float*, int, int, int, int, int, int, int,
                                   int, float, float, uint64_t, uint64_t, cudaStream_t);
template int xattn_simt_fwd<__nv_bfloat16>(const void*, const void*, const void*, void*, float*, int, int, int, int, int,
                                           int, int, int, float, float, uint64_t, uint64_t, cudaStream_t);
template int xattn_simt_bwd<float>(const void*, const void*, const void*, const void*, const void*, const float*, void*,
                                   void*, void*, int, int, int, int, int, int, int, int, float, float, uint64_t, uint64_t,
                                   cudaStream_t);
template int xattn_simt_bwd<__nv_bfloat16>(const void*, const void*, const void*, const void*, const void*, const float*,
                                           void*, void*, void*, int, int, int, int, int, int, int, int, float, float,
                                           uint64_t, uint64_t, cudaStream_t);

}  // namespace xtag
