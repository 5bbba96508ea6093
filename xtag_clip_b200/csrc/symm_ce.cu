// Symmetric cross-entropy on a MATERIALISED square score matrix: the loss of the TQN fusion head
// (reference src/open_clip/tagging_heads/asymmetric_loss.py:54-65, DQNCOSLoss):
//     loss = ( CE(X, arange) + CE(X^T, arange) ) / 2
//          = 0.5 * [ mean_i(LSE_j X_ij - X_ii) + mean_j(LSE_i X_ij - X_jj) ]
//     dX   = g * ( softmax_row(X) + softmax_col(X) - 2 I ) / (2 n)
// Unlike the contrastive head the matrix exists in HBM (it is the output of an MLP over attention outputs), so the
// kernels are HBM-bound elementwise passes: the forward reads X ONCE (row log-sum-exps complete per strip of 32
// rows, column partials per strip in the log2 domain, finished by the shared lse_reduce kernel), the backward reads X
// once and writes dX once.  (The host composition this replaces made a scaled copy and a transposed copy of X in
// the forward and six eager passes over n x n fp32 temporaries in the backward.)
#include "common.cuh"

namespace xtag {

int launch_lse_reduce(const float* parts, int P, int n, float in_mul, float out_mul, float* out, cudaStream_t st);

constexpr int SCE_ROWS = 32;     // rows per CTA strip: 8 warps x 4 rows
constexpr int SCE_COLS = 128;    // columns per iteration: 4 per lane

template <typename T>
__device__ __forceinline__ void load4(const T* p, bool full, int valid, float (&v)[4]);
template <>
__device__ __forceinline__ void load4<float>(const float* p, bool full, int valid, float (&v)[4]) {
  if (full) {
    const float4 f = *reinterpret_cast<const float4*>(p);
    v[0] = f.x; v[1] = f.y; v[2] = f.z; v[3] = f.w;
  } else {
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] = (k < valid) ? p[k] : -INFINITY;
  }
}
template <>
__device__ __forceinline__ void load4<__nv_bfloat16>(const __nv_bfloat16* p, bool full, int valid, float (&v)[4]) {
  if (full) {
    const uint2 u = *reinterpret_cast<const uint2*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
    const float2 a = __bfloat1622float2(h[0]), b = __bfloat1622float2(h[1]);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
  } else {
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] = (k < valid) ? __bfloat162float(p[k]) : -INFINITY;
  }
}

// One CTA = rows [32*blockIdx.x, +32), all columns.  Warp w owns rows 4w .. 4w+3; lane l owns columns c0 + 4l .. +3 of
// every 128-column chunk.  vec: rows are 16-byte (fp32) / 8-byte (bf16) aligned and ld % 4 == 0.
template <typename T>
__global__ void __launch_bounds__(256) symm_ce_fwd_kernel(const T* __restrict__ x, int n, long ld, int vec,
                                                          float* __restrict__ row_lse, float* __restrict__ diag,
                                                          float* __restrict__ col_part) {
  __shared__ float sm_m[8][SCE_COLS], sm_l[8][SCE_COLS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r0 = blockIdx.x * SCE_ROWS + warp * 4;
  float rm[4], rl[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) { rm[i] = -INFINITY; rl[i] = 0.f; }
  for (int c0 = 0; c0 < n; c0 += SCE_COLS) {
    const int c = c0 + lane * 4;
    const int valid = n - c;                                   // columns of this lane inside the matrix
    float v[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (r0 + i < n && valid > 0) {
        load4<T>(x + (size_t)(r0 + i) * ld + c, vec && valid >= 4, valid, v[i]);
#pragma unroll
        for (int k = 0; k < 4; ++k) v[i][k] *= kLog2e;
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) v[i][k] = -INFINITY;
      }
    }
    // rows: online (max, sum) per lane
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float mx = fmaxf(fmaxf(v[i][0], v[i][1]), fmaxf(v[i][2], v[i][3]));
      if (mx > -INFINITY) {
        const float mn = fmaxf(rm[i], mx);
        float acc = rl[i] * exp2f(rm[i] - mn);                 // exp2f(-inf) == 0 covers the first chunk
#pragma unroll
        for (int k = 0; k < 4; ++k) acc += exp2f(v[i][k] - mn);
        rl[i] = acc;
        rm[i] = mn;
      }
    }
    // columns: this warp's 4 rows, then the 8 warps through shared memory
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float mx = fmaxf(fmaxf(v[0][k], v[1][k]), fmaxf(v[2][k], v[3][k]));
      float l = 0.f;
      if (mx > -INFINITY) {
#pragma unroll
        for (int i = 0; i < 4; ++i) l += exp2f(v[i][k] - mx);
      }
      sm_m[warp][lane * 4 + k] = mx;
      sm_l[warp][lane * 4 + k] = l;
    }
    __syncthreads();
    if (threadIdx.x < SCE_COLS && c0 + (int)threadIdx.x < n) {
      float mx = -INFINITY;
#pragma unroll
      for (int w = 0; w < 8; ++w) mx = fmaxf(mx, sm_m[w][threadIdx.x]);
      float tot = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w)
        tot += (sm_m[w][threadIdx.x] > -INFINITY) ? sm_l[w][threadIdx.x] * exp2f(sm_m[w][threadIdx.x] - mx) : 0.f;
      col_part[(size_t)blockIdx.x * n + c0 + threadIdx.x] = (tot > 0.f) ? mx + log2f(tot) : -INFINITY;
    }
    __syncthreads();
  }
  // rows: merge the 32 lanes
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float m = rm[i], l = rl[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float m2 = __shfl_xor_sync(0xffffffffu, m, o), l2 = __shfl_xor_sync(0xffffffffu, l, o);
      const float mn = fmaxf(m, m2);
      l = (mn > -INFINITY) ? l * exp2f(m - mn) + l2 * exp2f(m2 - mn) : 0.f;
      m = mn;
    }
    if (lane == 0 && r0 + i < n) {
      row_lse[r0 + i] = (m + log2f(l)) * kLn2;
      diag[r0 + i] = to_f32<T>(x[(size_t)(r0 + i) * ld + (r0 + i)]);
    }
  }
}

// dX_ij = gs * ( 2^(x_ij*log2e - rl2_i) + 2^(x_ij*log2e - cl2_j) - 2 [i == j] ),  gs = g / (2 n)
template <typename T>
__global__ void __launch_bounds__(256) symm_ce_bwd_kernel(const T* __restrict__ x, int n, long ld,
                                                          const float* __restrict__ row_lse,
                                                          const float* __restrict__ col_lse,
                                                          const float* __restrict__ grad_out, T* __restrict__ dx,
                                                          long lddx) {
  const float gs = grad_out[0] / (2.f * (float)n);
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const float ncl2 = -col_lse[j] * kLog2e;
  const int i0 = blockIdx.y * 16;
#pragma unroll 4
  for (int i = i0; i < i0 + 16 && i < n; ++i) {
    const float v = to_f32<T>(x[(size_t)i * ld + j]) * kLog2e;
    float d = exp2f(v - row_lse[i] * kLog2e) + exp2f(v + ncl2);
    if (i == j) d -= 2.f;
    dx[(size_t)i * lddx + j] = from_f32<T>(d * gs);
  }
}

}  // namespace xtag

using namespace xtag;

extern "C" size_t xtag_symm_ce_ws_bytes(int n) {
  if (n <= 0) return 0;
  return (size_t)((n + SCE_ROWS - 1) / SCE_ROWS) * (size_t)n * sizeof(float) + 256;
}

extern "C" int xtag_symm_ce_fwd(const void* x, int dtype, int n, long ld, float* row_lse, float* col_lse, float* diag,
                                float* loss_out, void* ws, size_t ws_bytes, void* stream) {
  XTAG_REQUIRE(x && row_lse && col_lse && diag && loss_out && n > 0 && ld >= n, XTAG_ERR_INVALID,
               "symm_ce_fwd: bad arguments");
  XTAG_REQUIRE(dtype == XTAG_F32 || dtype == XTAG_BF16, XTAG_ERR_INVALID, "symm_ce_fwd: bad dtype %d", dtype);
  XTAG_REQUIRE(ws && ws_bytes >= xtag_symm_ce_ws_bytes(n), XTAG_ERR_WORKSPACE, "symm_ce_fwd: workspace %zu < %zu",
               ws_bytes, xtag_symm_ce_ws_bytes(n));
  int rc = xtag_device_check();
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const int strips = (n + SCE_ROWS - 1) / SCE_ROWS;
  float* col_part = (float*)ws;
  const size_t esz = dtype == XTAG_F32 ? 4 : 2;
  const int vec = ((reinterpret_cast<uintptr_t>(x) % (4 * esz)) == 0 && ld % 4 == 0) ? 1 : 0;
  if (dtype == XTAG_F32)
    symm_ce_fwd_kernel<float><<<strips, 256, 0, st>>>((const float*)x, n, ld, vec, row_lse, diag, col_part);
  else
    symm_ce_fwd_kernel<__nv_bfloat16><<<strips, 256, 0, st>>>((const __nv_bfloat16*)x, n, ld, vec, row_lse, diag, col_part);
  XTAG_CHECK_LAUNCH();
  rc = launch_lse_reduce(col_part, strips, n, 1.f, kLn2, col_lse, st);
  if (rc) return rc;
  return xtag_clip_loss(row_lse, diag, col_lse, n, 0, loss_out, stream);
}

extern "C" int xtag_symm_ce_bwd(const void* x, int dtype, int n, long ld, const float* row_lse, const float* col_lse,
                                const float* grad_out, void* dx, long lddx, void* stream) {
  XTAG_REQUIRE(x && row_lse && col_lse && grad_out && dx && n > 0 && ld >= n && lddx >= n, XTAG_ERR_INVALID,
               "symm_ce_bwd: bad arguments");
  XTAG_REQUIRE(dtype == XTAG_F32 || dtype == XTAG_BF16, XTAG_ERR_INVALID, "symm_ce_bwd: bad dtype %d", dtype);
  int rc = xtag_device_check();
  if (rc) return rc;
  const dim3 grid((n + 255) / 256, (n + 15) / 16);
  if (dtype == XTAG_F32)
    symm_ce_bwd_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)x, n, ld, row_lse, col_lse, grad_out,
                                                                     (float*)dx, lddx);
  else
    symm_ce_bwd_kernel<__nv_bfloat16><<<grid, 256, 0, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)x, n, ld, row_lse, col_lse, grad_out, (__nv_bfloat16*)dx, lddx);
  XTAG_CHECK_LAUNCH();
  return XTAG_OK;
}
