// K5: AsymmetricLoss forward + d/dx, and the per-category top-1 tag indices of prepare_control_words.
// Reference: src/open_clip/tagging_heads/asymmetric_loss.py:16-50 and src/open_clip/model.py:354-374.
// Elementwise and tiny ([b,44]): one 1024-thread CTA, fp64 block reduction, deterministic.
#include "common.cuh"

namespace xtag {

template <typename T>
__global__ void __launch_bounds__(1024) asl_kernel(const T* __restrict__ x, const float* __restrict__ y, int n,
                                                   float gamma_neg, float gamma_pos, float clip, float eps,
                                                   float* __restrict__ loss_out, float* __restrict__ dx) {
  __shared__ double red[32];
  double local = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float xv = to_f32(x[i]);
    const float yv = y[i];
    const float p = 1.f / (1.f + expf(-xv));                  // xs_pos
    float pn = 1.f - p;                                        // xs_neg
    const bool clipped_hi = (clip > 0.f) && (pn + clip > 1.f);
    if (clip > 0.f) pn = fminf(pn + clip, 1.f);
    const float lp = logf(fmaxf(p, eps)), ln = logf(fmaxf(pn, eps));
    const float loss = yv * lp + (1.f - yv) * ln;
    float w = 1.f;
    if (gamma_neg > 0.f || gamma_pos > 0.f) {
      const float pt = p * yv + pn * (1.f - yv);
      const float gam = gamma_pos * yv + gamma_neg * (1.f - yv);
      w = (gam == 0.f) ? 1.f : powf(1.f - pt, gam);           // no gradient through w (asymmetric_loss.py:41-48)
    }
    local += (double)(loss * w);
    if (dx) {
      // d/dx [ y log(clamp(p,eps)) + (1-y) log(clamp(pn,eps)) ] * w ; clamp/min pass zero gradient when active
      const float dp = p * (1.f - p);
      const float g_pos = (p >= eps) ? dp / p : 0.f;
      const float g_neg = (!clipped_hi && pn >= eps) ? -dp / pn : 0.f;
      dx[i] = -(yv * g_pos + (1.f - yv) * g_neg) * w;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x < 32) {
    double t = red[threadIdx.x];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (threadIdx.x == 0) loss_out[0] = (float)(-t);
  }
}

// per sample, per category (sizes 3,4,3,4,4,4 over 22 tags): argmax_j sigmoid(l[j]) + sigmoid(l[22+j])
template <typename T>
__global__ void __launch_bounds__(256) tag_top1_kernel(const T* __restrict__ x, int rows, int32_t* __restrict__ idx6) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * 6) return;
  const int r = i / 6, cat = i % 6;
  const int sizes[6] = {3, 4, 3, 4, 4, 4};
  int pos = 0;
  for (int c = 0; c < cat; ++c) pos += sizes[c];
  const T* xr = x + (size_t)r * 44;
  float best = -1.f;
  int arg = pos;
  for (int j = pos; j < pos + sizes[cat]; ++j) {
    const float s = 1.f / (1.f + expf(-to_f32(xr[j]))) + 1.f / (1.f + expf(-to_f32(xr[22 + j])));
    if (s > best) { best = s; arg = j; }
  }
  idx6[i] = arg;
}

}  // namespace xtag

using namespace xtag;

extern "C" int xtag_asl_fwd(const void* x, int x_dtype, const float* y, int rows, int cols, float gamma_neg,
                            float gamma_pos, float clip, float eps, float* loss_out, float* dx, int32_t* idx6,
                            void* stream) {
  XTAG_REQUIRE(x && y && loss_out && rows > 0 && cols > 0, XTAG_ERR_INVALID, "asl_fwd: bad arguments");
  XTAG_REQUIRE(x_dtype == XTAG_F32 || x_dtype == XTAG_BF16, XTAG_ERR_INVALID, "asl_fwd: bad dtype");
  XTAG_REQUIRE(idx6 == nullptr || cols == 44, XTAG_ERR_UNSUPPORTED, "asl_fwd: control-word indices need 44 logits per row");
  cudaStream_t st = (cudaStream_t)stream;
  const int n = rows * cols;
  if (x_dtype == XTAG_F32)
    asl_kernel<float><<<1, 1024, 0, st>>>((const float*)x, y, n, gamma_neg, gamma_pos, clip, eps, loss_out, dx);
  else
    asl_kernel<__nv_bfloat16><<<1, 1024, 0, st>>>((const __nv_bfloat16*)x, y, n, gamma_neg, gamma_pos, clip, eps, loss_out, dx);
  XTAG_CHECK_LAUNCH();
  if (idx6) {
    const int t = rows * 6;
    if (x_dtype == XTAG_F32)
      tag_top1_kernel<float><<<(t + 255) / 256, 256, 0, st>>>((const float*)x, rows, idx6);
    else
      tag_top1_kernel<__nv_bfloat16><<<(t + 255) / 256, 256, 0, st>>>((const __nv_bfloat16*)x, rows, idx6);
    XTAG_CHECK_LAUNCH();
  }
  return XTAG_OK;
}
