// K3: fused L2-normalise + cast (+ optional transposed copy) and its backward.
// Replaces F.normalize(features, dim=-1) of CLIP.encode_image / encode_text
// (reference src/open_clip/model.py:311-313, 332-333).
//
// HBM-bound: algorithmic bytes = rows*dim*(sizeof(in)+sizeof(out)).  One warp owns one row, reads it
// with 16-byte vector loads (second pass over the <= 8 KB row hits L1), writes 16-byte vectors.
// Grid = multiple of the SM count (grid-stride over rows).
#include "common.cuh"

namespace xtag {

template <typename T> struct Vec8;   // 8 elements
template <> struct Vec8<float> {
  float v[8];
  __device__ __forceinline__ void load(const float* p) {
    float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
  __device__ __forceinline__ void store(float* p) const {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
  }
};
template <> struct Vec8<__nv_bfloat16> {
  float v[8];
  __device__ __forceinline__ void load(const __nv_bfloat16* p) {
    uint4 r = *reinterpret_cast<const uint4*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float2 f = __bfloat1622float2(h[i]);
      v[2 * i] = f.x; v[2 * i + 1] = f.y;
    }
  }
  __device__ __forceinline__ void store(__nv_bfloat16* p) const {
    uint4 r;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    *reinterpret_cast<uint4*>(p) = r;
  }
};

template <typename TI, typename TO>
__global__ void __launch_bounds__(256) l2norm_fwd_kernel(const TI* __restrict__ x, TO* __restrict__ y,
                                                         float* __restrict__ inv_norm, int rows, int dim, float eps) {
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  const bool vec = (dim % 8 == 0);
  for (int r = blockIdx.x * warps_per_block + (threadIdx.x >> 5); r < rows; r += gridDim.x * warps_per_block) {
    const TI* xr = x + (size_t)r * dim;
    TO* yr = y + (size_t)r * dim;
    float ss = 0.f;
    if (vec) {
      for (int c = lane * 8; c < dim; c += 256) {
        Vec8<TI> a; a.load(xr + c);
#pragma unroll
        for (int i = 0; i < 8; ++i) ss = fmaf(a.v[i], a.v[i], ss);
      }
    } else {
      for (int c = lane; c < dim; c += 32) { float a = to_f32(xr[c]); ss = fmaf(a, a, ss); }
    }
    ss = warp_sum(ss);
    const float inv = 1.f / fmaxf(sqrtf(ss), eps);
    if (lane == 0 && inv_norm) inv_norm[r] = inv;
    if (vec) {
      for (int c = lane * 8; c < dim; c += 256) {
        Vec8<TI> a; a.load(xr + c);
        Vec8<TO> o;
#pragma unroll
        for (int i = 0; i < 8; ++i) o.v[i] = a.v[i] * inv;
        o.store(yr + c);
      }
    } else {
      for (int c = lane; c < dim; c += 32) yr[c] = from_f32<TO>(to_f32(xr[c]) * inv);
    }
  }
}

// gx = inv * (gy - y * (y . gy))   if ||x|| >= eps   (inv < 1/eps)
//      inv * gy                    otherwise         (clamp_min passes no gradient to the norm)
template <typename TG, typename TY, typename TO>
__global__ void __launch_bounds__(256) l2norm_bwd_kernel(const TG* __restrict__ gy, const TY* __restrict__ y,
                                                         const float* __restrict__ inv_norm, TO* __restrict__ gx,
                                                         int rows, int dim, float eps) {
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  const bool vec = (dim % 8 == 0);
  const float inv_clamped = 1.f / eps;
  for (int r = blockIdx.x * warps_per_block + (threadIdx.x >> 5); r < rows; r += gridDim.x * warps_per_block) {
    const TG* gr = gy + (size_t)r * dim;
    const TY* yr = y + (size_t)r * dim;
    TO* or_ = gx + (size_t)r * dim;
    const float inv = inv_norm[r];
    const bool clamped = inv >= inv_clamped;
    float dot = 0.f;
    if (!clamped) {
      if (vec) {
        for (int c = lane * 8; c < dim; c += 256) {
          Vec8<TG> a; a.load(gr + c);
          Vec8<TY> b; b.load(yr + c);
#pragma unroll
          for (int i = 0; i < 8; ++i) dot = fmaf(a.v[i], b.v[i], dot);
        }
      } else {
        for (int c = lane; c < dim; c += 32) dot = fmaf(to_f32(gr[c]), to_f32(yr[c]), dot);
      }
      dot = warp_sum(dot);
    }
    if (vec) {
      for (int c = lane * 8; c < dim; c += 256) {
        Vec8<TG> a; a.load(gr + c);
        Vec8<TY> b; b.load(yr + c);
        Vec8<TO> o;
#pragma unroll
        for (int i = 0; i < 8; ++i) o.v[i] = inv * (a.v[i] - b.v[i] * dot);
        o.store(or_ + c);
      }
    } else {
      for (int c = lane; c < dim; c += 32)
        or_[c] = from_f32<TO>(inv * (to_f32(gr[c]) - to_f32(yr[c]) * dot));
    }
  }
}

// out[c][r] = in[r][c]; 32x32 tiles through padded shared memory, coalesced both ways.
template <typename T>
__global__ void __launch_bounds__(256) transpose_kernel(const T* __restrict__ in, T* __restrict__ out, int rows, int cols) {
  __shared__ T tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
#pragma unroll
  for (int i = 0; i < 32; i += 8) {
    int r = r0 + ty + i, c = c0 + tx;
    if (r < rows && c < cols) tile[ty + i][tx] = in[(size_t)r * cols + c];
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 32; i += 8) {
    int c = c0 + ty + i, r = r0 + tx;
    if (r < rows && c < cols) out[(size_t)c * rows + r] = tile[tx][ty + i];
  }
}

int launch_transpose(const void* in, void* out, int dtype, int rows, int cols, cudaStream_t st) {
  dim3 grid((cols + 31) / 32, (rows + 31) / 32);
  if (dtype == XTAG_BF16)
    transpose_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)in, (__nv_bfloat16*)out, rows, cols);
  else
    transpose_kernel<float><<<grid, 256, 0, st>>>((const float*)in, (float*)out, rows, cols);
  XTAG_CHECK_LAUNCH();
  return XTAG_OK;
}

static int grid_for_rows(int rows) {
  int blocks = (rows + 7) / 8;
  int cap = num_sms() * 8;
  return blocks < cap ? (blocks > 0 ? blocks : 1) : cap;
}

}  // namespace xtag

using namespace xtag;

extern "C" int xtag_l2norm_fwd(const void* x, int x_dtype, void* y, int y_dtype, void* yT, float* inv_norm,
                               int rows, int dim, float eps, void* stream) {
  XTAG_REQUIRE(x && y && rows >= 0 && dim > 0, XTAG_ERR_INVALID, "l2norm_fwd: bad arguments");
  XTAG_REQUIRE((x_dtype == XTAG_F32 || x_dtype == XTAG_BF16) && (y_dtype == XTAG_F32 || y_dtype == XTAG_BF16),
               XTAG_ERR_INVALID, "l2norm_fwd: bad dtype");
  if (rows == 0) return XTAG_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = grid_for_rows(rows);
  if (x_dtype == XTAG_F32 && y_dtype == XTAG_F32)
    l2norm_fwd_kernel<float, float><<<grid, 256, 0, st>>>((const float*)x, (float*)y, inv_norm, rows, dim, eps);
  else if (x_dtype == XTAG_F32 && y_dtype == XTAG_BF16)
    l2norm_fwd_kernel<float, __nv_bfloat16><<<grid, 256, 0, st>>>((const float*)x, (__nv_bfloat16*)y, inv_norm, rows, dim, eps);
  else if (x_dtype == XTAG_BF16 && y_dtype == XTAG_F32)
    l2norm_fwd_kernel<__nv_bfloat16, float><<<grid, 256, 0, st>>>((const __nv_bfloat16*)x, (float*)y, inv_norm, rows, dim, eps);
  else
    l2norm_fwd_kernel<__nv_bfloat16, __nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)y, inv_norm, rows, dim, eps);
  XTAG_CHECK_LAUNCH();
  if (yT) return launch_transpose(y, yT, y_dtype, rows, dim, st);
  return XTAG_OK;
}

extern "C" int xtag_l2norm_bwd(const void* gy, int gy_dtype, const void* y, int y_dtype, const float* inv_norm,
                               void* gx, int gx_dtype, int rows, int dim, float eps, void* stream) {
  XTAG_REQUIRE(gy && y && inv_norm && gx && rows >= 0 && dim > 0, XTAG_ERR_INVALID, "l2norm_bwd: bad arguments");
  XTAG_REQUIRE(gy_dtype == y_dtype && (y_dtype == XTAG_F32 || y_dtype == XTAG_BF16) &&
                   (gx_dtype == XTAG_F32 || gx_dtype == XTAG_BF16),
               XTAG_ERR_UNSUPPORTED, "l2norm_bwd: gy must have y's dtype; dtypes must be f32 or bf16");
  if (rows == 0) return XTAG_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = grid_for_rows(rows);
  typedef __nv_bfloat16 bf16;
  if (y_dtype == XTAG_F32 && gx_dtype == XTAG_F32)
    l2norm_bwd_kernel<float, float, float><<<grid, 256, 0, st>>>((const float*)gy, (const float*)y, inv_norm, (float*)gx, rows, dim, eps);
  else if (y_dtype == XTAG_F32 && gx_dtype == XTAG_BF16)
    l2norm_bwd_kernel<float, float, bf16><<<grid, 256, 0, st>>>((const float*)gy, (const float*)y, inv_norm, (bf16*)gx, rows, dim, eps);
  else if (y_dtype == XTAG_BF16 && gx_dtype == XTAG_F32)
    l2norm_bwd_kernel<bf16, bf16, float><<<grid, 256, 0, st>>>((const bf16*)gy, (const bf16*)y, inv_norm, (float*)gx, rows, dim, eps);
  else
    l2norm_bwd_kernel<bf16, bf16, bf16><<<grid, 256, 0, st>>>((const bf16*)gy, (const bf16*)y, inv_norm, (bf16*)gx, rows, dim, eps);
  XTAG_CHECK_LAUNCH();
  return XTAG_OK;
}
