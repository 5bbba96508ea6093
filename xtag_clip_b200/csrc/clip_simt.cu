// Contrastive head, fp32 SIMT path ("fp32 mode" of BASELINE.json: parity <= 1e-5) and the small
// reduction kernels shared with the tcgen05 path.
//
// Reference semantics: ClipLoss.get_logits / forward, src/open_clip/loss.py:104-139.
//   fwd : S = scale * A Bm^T tile by tile (64x64 per CTA, FFMA, fp32 accumulate); per tile the row and
//         column (max, sum-exp) partials are written as log2-domain LSE partials, S is never stored.
//   bwd : S tile recomputed, dS staged in fp32, two strided FFMA GEMMs.
// This path exists for exactness (fp32 inputs, ragged shapes); the bf16 production path is clip_tc.cu.
#include "common.cuh"

namespace xtag {

constexpr int TS = 64;     // tile edge
constexpr int TKS = 16;    // k chunk

// C(i,j) = sum_k A[i*sa_m + k*sa_k] * B[j*sb_n + k*sb_k]; 256 threads, 4x4 per thread.
template <typename TA, typename TB>
__device__ __forceinline__ void simt_tile_mainloop(float (&acc)[4][4], const TA* __restrict__ A, long sa_m, long sa_k,
                                                   const TB* __restrict__ B, long sb_n, long sb_k,
                                                   int M, int N, int K, int m0, int n0,
                                                   float (*As)[TS + 4], float (*Bs)[TS + 4]) {
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const bool a_kfast = (sa_k == 1), b_kfast = (sb_k == 1);
  for (int k0 = 0; k0 < K; k0 += TKS) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int idx = tid + e * 256;
      int r, kk;
      if (a_kfast) { r = idx >> 4; kk = idx & 15; } else { r = idx & 63; kk = idx >> 6; }
      const int gi = m0 + r, gk = k0 + kk;
      As[kk][r] = (gi < M && gk < K) ? to_f32(A[(long)gi * sa_m + (long)gk * sa_k]) : 0.f;
      if (b_kfast) { r = idx >> 4; kk = idx & 15; } else { r = idx & 63; kk = idx >> 6; }
      const int gj = n0 + r;
      const int gk2 = k0 + kk;
      Bs[kk][r] = (gj < N && gk2 < K) ? to_f32(B[(long)gj * sb_n + (long)gk2 * sb_k]) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < TKS; ++kk) {
      float a[4], b[4];
      *reinterpret_cast<float4*>(a) = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      *reinterpret_cast<float4*>(b) = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
}

// ---- forward: LSE partials ----------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) simt_clip_fwd_kernel(const T* __restrict__ A, const T* __restrict__ Bm,
                                                            int M, int N, int D, const float* __restrict__ scale_p, int label_offset,
                                                            float* __restrict__ row_part,   // [ntn][M] log2 domain
                                                            float* __restrict__ col_part,   // [ntm][N] log2 domain
                                                            float* __restrict__ diag) {     // [M] natural units
  __shared__ float As[TKS][TS + 4];
  __shared__ float Bs[TKS][TS + 4];
  __shared__ float Ts[TS][TS + 1];
  const int m0 = blockIdx.y * TS, n0 = blockIdx.x * TS;
  float acc[4][4];
  simt_tile_mainloop<T, T>(acc, A, D, 1, Bm, D, 1, M, N, D, m0, n0, As, Bs);
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const float sl2 = scale_p[0] * kLog2e;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) Ts[ty * 4 + i][tx * 4 + j] = acc[i][j] * sl2;
  __syncthreads();
  const int rows = min(TS, M - m0), cols = min(TS, N - n0);
  if (tid < TS) {
    const int r = tid;
    if (r < rows) {
      float mx = -INFINITY;
      for (int c = 0; c < cols; ++c) mx = fmaxf(mx, Ts[r][c]);
      float l = 0.f;
      for (int c = 0; c < cols; ++c) l += exp2f(Ts[r][c] - mx);
      row_part[(size_t)blockIdx.x * M + m0 + r] = mx + log2f(l);
      const int lab = (label_offset < 0) ? -1 : m0 + r + label_offset;
      if (lab >= n0 && lab < n0 + cols) diag[m0 + r] = Ts[r][lab - n0] * kLn2;
    }
  } else if (tid < 2 * TS) {
    const int c = tid - TS;
    if (c < cols) {
      float mx = -INFINITY;
      for (int r = 0; r < rows; ++r) mx = fmaxf(mx, Ts[r][c]);
      float l = 0.f;
      for (int r = 0; r < rows; ++r) l += exp2f(Ts[r][c] - mx);
      col_part[(size_t)blockIdx.y * N + n0 + c] = mx + log2f(l);
    }
  }
}

// out[i] = out_mul * log2( sum_p 2^(parts[p*n + i] * in_mul) )        (in_mul/out_mul convert ln <-> log2)
// 256 threads = 32 consecutive i (coalesced 128-byte rows of `parts`) x 8 interleaved groups of p; each thread keeps
// an online (max, sum) over its parts, the 8 groups are merged through shared memory.
__global__ void __launch_bounds__(256) lse_reduce_kernel(const float* __restrict__ parts, int P, int n,
                                                         float in_mul, float out_mul, float* __restrict__ out) {
  __shared__ float sm_m[8][33], sm_l[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int i0 = blockIdx.x * 32; i0 < n; i0 += gridDim.x * 32) {
    const int i = i0 + tx;
    float m = -INFINITY, l = 0.f;
    if (i < n) {
      for (int p = ty; p < P; p += 8) {
        const float v = parts[(size_t)p * n + i] * in_mul;
        if (v > m) { l = l * exp2f(m - v) + 1.f; m = v; }      // exp2f(-inf) == 0 covers the first element
        else if (v > -INFINITY) l += exp2f(v - m);
      }
    }
    sm_m[ty][tx] = m;
    sm_l[ty][tx] = l;
    __syncthreads();
    if (ty == 0 && i < n) {
      float mx = -INFINITY;
#pragma unroll
      for (int g = 0; g < 8; ++g) mx = fmaxf(mx, sm_m[g][tx]);
      float tot = 0.f;
#pragma unroll
      for (int g = 0; g < 8; ++g) tot += (sm_m[g][tx] > -INFINITY) ? sm_l[g][tx] * exp2f(sm_m[g][tx] - mx) : 0.f;
      out[i] = (mx + log2f(tot)) * out_mul;
    }
    __syncthreads();
  }
}

// Two independent reductions of log2-domain partials in ONE launch (the row and the column partials a K1 launch leaves
// behind): blocks [0, nb0) work on problem 0, the rest on problem 1.  512 threads = 32 consecutive outputs x 16 groups
// of parts; every thread first takes the maximum of its parts and then sums the exponentials, both with independent
// (unrolled) loads, instead of the serial online update of lse_reduce_kernel: the row reduction of a sharded forward
// (256 partials per row, 4096 rows) drops from 15 us to a few.
struct Reduce2 {
  const float* parts[2];
  float* out[2];
  int P[2], n[2], nb0;
};
__global__ void __launch_bounds__(512) lse_reduce2_kernel(const __grid_constant__ Reduce2 r) {
  __shared__ float sm_m[16][33], sm_l[16][33];
  const int which = ((int)blockIdx.x >= r.nb0) ? 1 : 0;
  const int blk = which ? (int)blockIdx.x - r.nb0 : (int)blockIdx.x;
  const float* __restrict__ parts = r.parts[which];
  const int P = r.P[which], n = r.n[which];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int i = blk * 32 + tx;
  float m = -INFINITY, l = 0.f;
  if (i < n) {
    for (int p0 = ty; p0 < P; p0 += 16 * 8) {
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int p = p0 + u * 16;
        v[u] = (p < P) ? parts[(size_t)p * n + i] : -INFINITY;
      }
      float mx = m;
#pragma unroll
      for (int u = 0; u < 8; ++u) mx = fmaxf(mx, v[u]);
      if (mx > -INFINITY) {
        float acc = (m > -INFINITY) ? l * exp2f(m - mx) : 0.f;
#pragma unroll
        for (int u = 0; u < 8; ++u) acc += exp2f(v[u] - mx);     // exp2f(-inf) == 0
        l = acc;
        m = mx;
      }
    }
  }
  sm_m[ty][tx] = m;
  sm_l[ty][tx] = l;
  __syncthreads();
  if (ty == 0 && i < n) {
    float mx = -INFINITY;
#pragma unroll
    for (int g = 0; g < 16; ++g) mx = fmaxf(mx, sm_m[g][tx]);
    float tot = 0.f;
#pragma unroll
    for (int g = 0; g < 16; ++g) tot += (sm_m[g][tx] > -INFINITY) ? sm_l[g][tx] * exp2f(sm_m[g][tx] - mx) : 0.f;
    r.out[which][i] = (mx + log2f(tot)) * kLn2;
  }
}

int launch_lse_reduce(const float* parts, int P, int n, float in_mul, float out_mul, float* out, cudaStream_t st) {
  int blocks = (n + 31) / 32;
  if (blocks > num_sms() * 8) blocks = num_sms() * 8;
  if (blocks < 1) blocks = 1;
  lse_reduce_kernel<<<blocks, 256, 0, st>>>(parts, P, n, in_mul, out_mul, out);
  XTAG_CHECK_LAUNCH();
  return XTAG_OK;
}

// loss = 0.5 * ( mean_i(row_lse_i - diag_i) + mean_i(col_lse[off+i] - diag_i) ); one CTA, fp64 accumulate
__global__ void __launch_bounds__(1024) clip_loss_kernel(const float* __restrict__ row_lse, const float* __restrict__ diag,
                                                         const float* __restrict__ col_lse, int M, int label_offset,
                                                         float* __restrict__ loss_out) {
  __shared__ double red[32];
  double s = 0.0;
  for (int i = threadIdx.x; i < M; i += blockDim.x)
    s += (double)row_lse[i] + (double)col_lse[label_offset + i] - 2.0 * (double)diag[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    double t = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (threadIdx.x == 0) loss_out[0] = (float)(0.5 * t / (double)M);
  }
}

// ---- backward: dS tile + dscale partial ----------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) simt_clip_ds_kernel(const T* __restrict__ A, const T* __restrict__ Bm,
                                                           int M, int N, int D, const float* __restrict__ scale_p, int label_offset,
                                                           const float* __restrict__ row_lse, const float* __restrict__ col_lse,
                                                           float w_row, float w_col, float w_diag,
                                                           const float* __restrict__ grad_out,
                                                           float* __restrict__ dS, float* __restrict__ dscale_part) {
  __shared__ float As[TKS][TS + 4];
  __shared__ float Bs[TKS][TS + 4];
  __shared__ float red[32];
  const int m0 = blockIdx.y * TS, n0 = blockIdx.x * TS;
  float acc[4][4];
  simt_tile_mainloop<T, T>(acc, A, D, 1, Bm, D, 1, M, N, D, m0, n0, As, Bs);
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const float g = grad_out[0];
  const float sl2 = scale_p[0] * kLog2e;
  float part = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gi = m0 + ty * 4 + i;
    if (gi >= M) continue;
    const float rl = row_lse[gi] * kLog2e;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gj = n0 + tx * 4 + j;
      if (gj >= N) continue;
      const float v = acc[i][j] * sl2;
      float d = w_row * exp2f(v - rl) + w_col * exp2f(v - col_lse[gj] * kLog2e);
      if (gj == gi + label_offset) d -= w_diag;
      d *= g;
      dS[(size_t)gi * N + gj] = d;
      part = fmaf(d, acc[i][j], part);
    }
  }
  part = block_sum(part, red);
  if (tid == 0) dscale_part[blockIdx.y * gridDim.x + blockIdx.x] = part;
}

__global__ void __launch_bounds__(1024) sum_into_kernel(const float* __restrict__ parts, int n, float* __restrict__ out) {
  __shared__ double red[32];
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += (double)parts[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    double t = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (threadIdx.x == 0) out[0] += (float)t;
  }
}

int launch_sum_into(const float* parts, int n, float* out, cudaStream_t st) {
  sum_into_kernel<<<1, 1024, 0, st>>>(parts, n, out);
  XTAG_CHECK_LAUNCH();
  return XTAG_OK;
}

// C[i*ldc + j] = alpha * sum_k A(i,k) B(j,k)   (strided operands, fp32 accumulate)
template <typename TA, typename TB, typename TC>
__global__ void __launch_bounds__(256) simt_gemm_kernel(const TA* __restrict__ A, long sa_m, long sa_k,
                                                        const TB* __restrict__ B, long sb_n, long sb_k,
                                                        TC* __restrict__ C, int ldc, int M, int N, int K, const float* __restrict__ alpha_p) {
  __shared__ float As[TKS][TS + 4];
  __shared__ float Bs[TKS][TS + 4];
  const int m0 = blockIdx.y * TS, n0 = blockIdx.x * TS;
  float acc[4][4];
  simt_tile_mainloop<TA, TB>(acc, A, sa_m, sa_k, B, sb_n, sb_k, M, N, K, m0, n0, As, Bs);
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const float alpha = alpha_p ? alpha_p[0] : 1.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gi = m0 + ty * 4 + i;
    if (gi >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gj = n0 + tx * 4 + j;
      if (gj < N) C[(size_t)gi * ldc + gj] = from_f32<TC>(alpha * acc[i][j]);
    }
  }
}

template <typename T>
static int simt_fwd(const T* A, const T* Bm, int M, int N, int D, const float* scale, int label_offset,
                    float* row_lse, float* col_lse, float* diag, void* ws, size_t ws_bytes, cudaStream_t st) {
  const int ntm = (M + TS - 1) / TS, ntn = (N + TS - 1) / TS;
  const size_t need = ((size_t)ntn * M + (size_t)ntm * N) * sizeof(float);
  XTAG_REQUIRE(ws && ws_bytes >= need, XTAG_ERR_WORKSPACE, "clip_fwd(simt): workspace %zu < %zu", ws_bytes, need);
  float* row_part = (float*)ws;
  float* col_part = row_part + (size_t)ntn * M;
  simt_clip_fwd_kernel<T><<<dim3(ntn, ntm), 256, 0, st>>>(A, Bm, M, N, D, scale, label_offset, row_part, col_part, diag);
  XTAG_CHECK_LAUNCH();
  int rc = launch_lse_reduce(row_part, ntn, M, 1.f, kLn2, row_lse, st);
  if (rc) return rc;
  return launch_lse_reduce(col_part, ntm, N, 1.f, kLn2, col_lse, st);
}

size_t simt_fwd_ws(int M, int N) {
  const size_t ntm = (M + TS - 1) / TS, ntn = (N + TS - 1) / TS;
  return (ntn * (size_t)M + ntm * (size_t)N) * sizeof(float) + 256;
}
size_t simt_bwd_ws(int M, int N) {
  const size_t ntm = (M + TS - 1) / TS, ntn = (N + TS - 1) / TS;
  return ((size_t)M * N + ntm * ntn) * sizeof(float) + 256;
}

int simt_clip_fwd(const void* A, const void* Bm, int dtype, int M, int N, int D, const float* scale, int label_offset,
                  float* row_lse, float* col_lse, float* diag, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (dtype == XTAG_F32)
    return simt_fwd<float>((const float*)A, (const float*)Bm, M, N, D, scale, label_offset, row_lse, col_lse, diag, ws, ws_bytes, st);
  return simt_fwd<__nv_bfloat16>((const __nv_bfloat16*)A, (const __nv_bfloat16*)Bm, M, N, D, scale, label_offset,
                                 row_lse, col_lse, diag, ws, ws_bytes, st);
}

template <typename T, typename TG>
static int simt_bwd(const T* A, const T* Bm, int M, int N, int D, const float* scale, int label_offset,
                    const float* row_lse, const float* col_lse, float w_row, float w_col, float w_diag,
                    const float* grad_out, TG* dA, TG* dB, float* dscale, void* ws, size_t ws_bytes, int flags,
                    cudaStream_t st) {
  const int ntm = (M + TS - 1) / TS, ntn = (N + TS - 1) / TS;
  const size_t need = ((size_t)M * N + (size_t)ntm * ntn) * sizeof(float);
  XTAG_REQUIRE(ws && ws_bytes >= need, XTAG_ERR_WORKSPACE, "clip_bwd(simt): workspace %zu < %zu", ws_bytes, need);
  float* dS = (float*)ws;
  float* part = dS + (size_t)M * N;
  if (!(flags & XTAG_BWD_REUSE_DS)) {
    simt_clip_ds_kernel<T><<<dim3(ntn, ntm), 256, 0, st>>>(A, Bm, M, N, D, scale, label_offset, row_lse, col_lse,
                                                            w_row, w_col, w_diag, grad_out, dS, part);
    XTAG_CHECK_LAUNCH();
    if (dscale) {
      int rc = launch_sum_into(part, ntm * ntn, dscale, st);
      if (rc) return rc;
    }
  }
  const int ntd = (D + TS - 1) / TS;
  if (dA) {   // dA[i,d] = scale * sum_j dS[i,j] Bm[j,d]
    simt_gemm_kernel<float, T, TG><<<dim3(ntd, ntm), 256, 0, st>>>(dS, N, 1, Bm, 1, D, dA, D, M, D, N, scale);
    XTAG_CHECK_LAUNCH();
  }
  if (dB) {   // dB[j,d] = scale * sum_i dS[i,j] A[i,d]
    simt_gemm_kernel<float, T, TG><<<dim3(ntd, ntn), 256, 0, st>>>(dS, 1, N, A, 1, D, dB, D, N, D, M, scale);
    XTAG_CHECK_LAUNCH();
  }
  return XTAG_OK;
}

int simt_clip_bwd(const void* A, const void* Bm, int dtype, int M, int N, int D, const float* scale, int label_offset,
                  const float* row_lse, const float* col_lse, float w_row, float w_col, float w_diag,
                  const float* grad_out, void* dA, void* dB, int grad_dtype, float* dscale,
                  void* ws, size_t ws_bytes, int flags, cudaStream_t st) {
  typedef __nv_bfloat16 bf16;
  if (dtype == XTAG_F32 && grad_dtype == XTAG_F32)
    return simt_bwd<float, float>((const float*)A, (const float*)Bm, M, N, D, scale, label_offset, row_lse, col_lse,
                                  w_row, w_col, w_diag, grad_out, (float*)dA, (float*)dB, dscale, ws, ws_bytes, flags, st);
  if (dtype == XTAG_BF16 && grad_dtype == XTAG_F32)
    return simt_bwd<bf16, float>((const bf16*)A, (const bf16*)Bm, M, N, D, scale, label_offset, row_lse, col_lse,
                                 w_row, w_col, w_diag, grad_out, (float*)dA, (float*)dB, dscale, ws, ws_bytes, flags, st);
  if (dtype == XTAG_BF16 && grad_dtype == XTAG_BF16)
    return simt_bwd<bf16, bf16>((const bf16*)A, (const bf16*)Bm, M, N, D, scale, label_offset, row_lse, col_lse,
                                w_row, w_col, w_diag, grad_out, (bf16*)dA, (bf16*)dB, dscale, ws, ws_bytes, flags, st);
  return simt_bwd<float, bf16>((const float*)A, (const float*)Bm, M, N, D, scale, label_offset, row_lse, col_lse,
                               w_row, w_col, w_diag, grad_out, (bf16*)dA, (bf16*)dB, dscale, ws, ws_bytes, flags, st);
}

// ---- peer-memory reductions (NVLink loads from symmetric buffers of the other ranks) -------------------------------
// out[j] = ln sum_w exp(parts[w][j]); parts[w] is rank w's buffer (peer-mapped device pointer)
__global__ void __launch_bounds__(256) lse_combine_ptrs_kernel(const float* const* __restrict__ parts, int W, int n,
                                                               float* __restrict__ out) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    float v[16];
    float mx = -INFINITY;
    for (int w = 0; w < W; ++w) {
      v[w] = parts[w][i] * kLog2e;
      mx = fmaxf(mx, v[w]);
    }
    float l = 0.f;
    for (int w = 0; w < W; ++w) l += exp2f(v[w] - mx);
    out[i] = (mx + log2f(l)) * kLn2;
  }
}

// Column-LSE combine over the W peer buffers + this rank's loss + the exchange epoch bump, in ONE launch.  Every block
// combines its columns (each thread reads the W peer values of one column: coalesced, all loads independent); the
// columns that are this rank's labels [label_offset, label_offset + M) also contribute
//   row_lse_i + col_lse[off + i] - 2 diag_i   to a per-block fp64 partial.  The blocks publish their partials in
// `scratch` and take a ticket; the last one sums them IN BLOCK ORDER (deterministic), writes the loss, bumps the epoch
// and resets the ticket.  scratch: 8 bytes ticket + gridDim.x doubles, zero-initialised once by the caller.
// (A first version let one extra block re-read its M label columns from the peers: 16 dependent rounds of NVLink loads
// per thread -- 215 us at 8 GPUs.)
__global__ void __launch_bounds__(256) lse_combine_loss_kernel(const float* const* __restrict__ parts, int W, int n,
                                                               float* __restrict__ out,
                                                               const float* __restrict__ row_lse,
                                                               const float* __restrict__ diag, int M, int label_offset,
                                                               float* __restrict__ loss_out, int* __restrict__ epoch,
                                                               unsigned long long* __restrict__ scratch) {
  __shared__ double red[8];
  __shared__ bool last;
  double s = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    float v[16];
    float mx = -INFINITY;
    for (int w = 0; w < W; ++w) {
      v[w] = parts[w][i] * kLog2e;
      mx = fmaxf(mx, v[w]);
    }
    float l = 0.f;
    for (int w = 0; w < W; ++w) l += exp2f(v[w] - mx);
    const float c = (mx + log2f(l)) * kLn2;
    out[i] = c;
    const int r = i - label_offset;
    if (r >= 0 && r < M) s += (double)row_lse[r] + (double)c - 2.0 * (double)diag[r];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  double* partials = reinterpret_cast<double*>(scratch + 1);
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int k = 0; k < 8; ++k) t += red[k];
    partials[blockIdx.x] = t;
    __threadfence();
    last = atomicAdd(scratch, 1ull) == (unsigned long long)(gridDim.x - 1);
  }
  __syncthreads();
  if (last) {
    // the whole block sums the partials: a fixed tree (deterministic), not one thread walking gridDim.x values
    __threadfence();
    double t = 0.0;
    for (unsigned k = threadIdx.x; k < gridDim.x; k += blockDim.x) t += reinterpret_cast<volatile double*>(partials)[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    __syncthreads();                        // red[] was read by thread 0 above
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = t;
    __syncthreads();
    if (threadIdx.x == 0) {
      double tot = 0.0;
      for (int k = 0; k < 8; ++k) tot += red[k];
      loss_out[0] = (float)(0.5 * tot / (double)M);
      if (epoch) epoch[0] += 1;
      scratch[0] = 0ull;                    // ready for the next launch
    }
  }
}

// out[i] = sum_w parts[w][i]  (bf16 in, fp32 accumulate, bf16 out): the reduce step of the pull-based reduce-scatter
__global__ void __launch_bounds__(256) sum_ptrs_bf16_kernel(const __nv_bfloat16* const* __restrict__ parts, int W,
                                                            size_t n8, __nv_bfloat16* __restrict__ out) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (size_t)gridDim.x * blockDim.x) {
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int w = 0; w < W; ++w) {
      const uint4 r = *reinterpret_cast<const uint4*>(parts[w] + i * 8);
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 f = __bfloat1622float2(h[k]);
        acc[2 * k] += f.x;
        acc[2 * k + 1] += f.y;
      }
    }
    uint4 o;
    __nv_bfloat162* ho = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
    for (int k = 0; k < 4; ++k) ho[k] = __floats2bfloat162_rn(acc[2 * k], acc[2 * k + 1]);
    *reinterpret_cast<uint4*>(out + i * 8) = o;
  }
}

}  // namespace xtag

using namespace xtag;

extern "C" int xtag_lse_combine_ptrs(const float* const* parts_dev, int W, int N, float* out, void* stream) {
  XTAG_REQUIRE(parts_dev && out && W > 0 && W <= 16 && N >= 0, XTAG_ERR_INVALID, "lse_combine_ptrs: bad arguments");
  if (N == 0) return XTAG_OK;
  int blocks = (N + 255) / 256;
  if (blocks > num_sms() * 4) blocks = num_sms() * 4;
  lse_combine_ptrs_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(parts_dev, W, N, out);
  XTAG_CHECK_LAUNCH();
  return XTAG_OK;
}

extern "C" size_t xtag_lse_combine_loss_scratch_bytes(void) { return 8 + 1024 * sizeof(double); }

extern "C" int xtag_lse_combine_ptrs_loss(const float* const* parts_dev, int W, int N, float* col_out,
                                          const float* row_lse, const float* diag, int M, int label_offset,
                                          float* loss_out, int* epoch, void* scratch, void* stream) {
  XTAG_REQUIRE(parts_dev && col_out && row_lse && diag && loss_out && scratch && W > 0 && W <= 16 && N > 0 && M > 0 &&
                   label_offset >= 0 && (long)label_offset + M <= (long)N,
               XTAG_ERR_INVALID, "lse_combine_ptrs_loss: bad arguments");
  int blocks = (N + 255) / 256;
  if (blocks > 1024) blocks = 1024;
  if (blocks > num_sms() * 4) blocks = num_sms() * 4;
  lse_combine_loss_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(parts_dev, W, N, col_out, row_lse, diag, M,
                                                                   label_offset, loss_out, epoch,
                                                                   (unsigned long long*)scratch);
  XTAG_CHECK_LAUNCH();
  return XTAG_OK;
}

extern "C" int xtag_lse_reduce2_log2(const float* parts0, int P0, int n0, float* out0, const float* parts1, int P1,
                                     int n1, float* out1, void* stream) {
  XTAG_REQUIRE(parts0 && out0 && parts1 && out1 && P0 > 0 && n0 > 0 && P1 > 0 && n1 > 0, XTAG_ERR_INVALID,
               "lse_reduce2_log2: bad arguments");
  Reduce2 r;
  r.parts[0] = parts0; r.parts[1] = parts1;
  r.out[0] = out0; r.out[1] = out1;
  r.P[0] = P0; r.P[1] = P1;
  r.n[0] = n0; r.n[1] = n1;
  r.nb0 = (n0 + 31) / 32;
  const int blocks = r.nb0 + (n1 + 31) / 32;
  lse_reduce2_kernel<<<blocks, 512, 0, (cudaStream_t)stream>>>(r);
  XTAG_CHECK_LAUNCH();
  return XTAG_OK;
}

extern "C" int xtag_sum_ptrs_bf16(const void* const* parts_dev, int W, size_t n, void* out, void* stream) {
  XTAG_REQUIRE(parts_dev && out && W > 0 && n % 8 == 0, XTAG_ERR_INVALID, "sum_ptrs_bf16: bad arguments (n %% 8 == 0)");
  if (n == 0) return XTAG_OK;
  const size_t n8 = n / 8;
  size_t blocks = (n8 + 255) / 256;
  if (blocks > (size_t)num_sms() * 8) blocks = (size_t)num_sms() * 8;
  sum_ptrs_bf16_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16* const*)parts_dev, W, n8,
                                                                     (__nv_bfloat16*)out);
  XTAG_CHECK_LAUNCH();
  return XTAG_OK;
}

extern "C" int xtag_lse_combine(const float* parts, int W, int N, float* out, void* stream) {
  XTAG_REQUIRE(parts && out && W > 0 && N >= 0, XTAG_ERR_INVALID, "lse_combine: bad arguments");
  if (N == 0) return XTAG_OK;
  return launch_lse_reduce(parts, W, N, kLog2e, kLn2, out, (cudaStream_t)stream);
}

extern "C" int xtag_clip_loss(const float* row_lse, const float* diag, const float* col_lse, int M, int label_offset,
                              float* loss_out, void* stream) {
  XTAG_REQUIRE(row_lse && diag && col_lse && loss_out && M > 0 && label_offset >= 0, XTAG_ERR_INVALID,
               "clip_loss: bad arguments");
  clip_loss_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(row_lse, diag, col_lse, M, label_offset, loss_out);
  XTAG_CHECK_LAUNCH();
  return XTAG_OK;
}
