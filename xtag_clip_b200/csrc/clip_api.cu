// C-ABI entry points of the contrastive head: argument validation and SIMT / tcgen05 selection.
#include "common.cuh"

namespace xtag {
size_t simt_fwd_ws(int M, int N);
size_t simt_bwd_ws(int M, int N);
int simt_clip_fwd(const void* A, const void* Bm, int dtype, int M, int N, int D, const float* scale, int label_offset,
                  float* row_lse, float* col_lse, float* diag, void* ws, size_t ws_bytes, cudaStream_t st);
int simt_clip_bwd(const void* A, const void* Bm, int dtype, int M, int N, int D, const float* scale, int label_offset,
                  const float* row_lse, const float* col_lse, float w_row, float w_col, float w_diag,
                  const float* grad_out, void* dA, void* dB, int grad_dtype, float* dscale,
                  void* ws, size_t ws_bytes, int flags, cudaStream_t st);
size_t tc_fwd_ws(int M, int N);
size_t tc_bwd_ws(int M, int N, int D);
int tc_clip_fwd(const void* A, const void* Bm, int M, int N, int D, const float* scale, int label_offset,
                float* row_lse, float* col_lse, float* diag, void* ws, size_t ws_bytes, cudaStream_t st);
int tc_clip_bwd(const void* A, const void* Bm, int M, int N, int D, const float* scale, int label_offset,
                const float* row_lse, const float* col_lse, float w_row, float w_col, float w_diag,
                const float* grad_out, void* dA, void* dB, int grad_dtype, float* dscale,
                void* ws, size_t ws_bytes, int flags, cudaStream_t st);

int tc_siglip_fwd(const void* A, const void* Bm, int M, int N, int D, const float* scale, const float* bias,
                  int label_offset, float w, float* out3, void* ws, size_t ws_bytes, int stage_ds, cudaStream_t st);
void tc_fwd_block_parts(int M, int N, int* row_parts, int* col_parts);
int tc_clip_fwd_block(const void* A, const void* Bm, int M, int N, int D, const float* scale, int label_offset,
                      float* row_part, float* col_part, int col_ld, float* diag, cudaStream_t st);
int tc_clip_fwd_stream(const void* A, const void* Bm, int M, int N, int D, const float* scale, int label_offset,
                       const int* order, const int* wait, int nblk, int blk_cols, const int* ready_flags,
                       const int* epoch, float* row_part, float* col_part, int col_ld, float* diag, cudaStream_t st);
int launch_lse_reduce(const float* parts, int P, int n, float in_mul, float out_mul, float* out, cudaStream_t st);

// tcgen05 path needs bf16 operands whose rows are 16-byte multiples (TMA global stride rule)
static bool tc_eligible(int dtype, int D) { return dtype == XTAG_BF16 && D % 8 == 0; }

static int resolve_impl(int impl, int dtype, int D, const char* who) {
  if (impl == XTAG_IMPL_AUTO) return tc_eligible(dtype, D) ? XTAG_IMPL_TC : XTAG_IMPL_SIMT;
  if (impl == XTAG_IMPL_TC && !tc_eligible(dtype, D)) {
    set_error("%s: XTAG_IMPL_TC needs bf16 inputs with D %% 8 == 0 (dtype=%d, D=%d)", who, dtype, D);
    return XTAG_ERR_UNSUPPORTED;
  }
  if (impl != XTAG_IMPL_SIMT && impl != XTAG_IMPL_TC) {
    set_error("%s: unknown impl %d", who, impl);
    return XTAG_ERR_INVALID;
  }
  return impl;
}
}  // namespace xtag

using namespace xtag;

extern "C" size_t xtag_clip_fwd_ws_bytes(int M, int N, int D, int dtype, int impl) {
  if (M <= 0 || N <= 0 || D <= 0) return 0;
  const int r = resolve_impl(impl, dtype, D, "clip_fwd_ws_bytes");
  if (r < 0) return 0;
  return r == XTAG_IMPL_TC ? tc_fwd_ws(M, N) : simt_fwd_ws(M, N);
}

extern "C" size_t xtag_clip_bwd_ws_bytes(int M, int N, int D, int dtype, int impl) {
  if (M <= 0 || N <= 0 || D <= 0) return 0;
  const int r = resolve_impl(impl, dtype, D, "clip_bwd_ws_bytes");
  if (r < 0) return 0;
  return r == XTAG_IMPL_TC ? tc_bwd_ws(M, N, D) : simt_bwd_ws(M, N);
}

extern "C" int xtag_clip_fwd(const void* A, const void* Bm, int dtype, int M, int N, int D, const float* scale,
                             int label_offset, float* row_lse, float* col_lse, float* diag, void* ws,
                             size_t ws_bytes, int impl, void* stream) {
  XTAG_REQUIRE(A && Bm && scale && row_lse && col_lse && diag, XTAG_ERR_INVALID, "clip_fwd: null pointer");
  XTAG_REQUIRE(M > 0 && N > 0 && D > 0, XTAG_ERR_INVALID, "clip_fwd: empty problem M=%d N=%d D=%d", M, N, D);
  XTAG_REQUIRE(dtype == XTAG_F32 || dtype == XTAG_BF16, XTAG_ERR_INVALID, "clip_fwd: bad dtype %d", dtype);
  XTAG_REQUIRE(label_offset == -1 || (label_offset >= 0 && (long)label_offset + M <= (long)N), XTAG_ERR_INVALID,
               "clip_fwd: labels [%d, %d) fall outside the %d columns", label_offset, label_offset + M, N);
  int rc = xtag_device_check();
  if (rc) return rc;
  const int r = resolve_impl(impl, dtype, D, "clip_fwd");
  if (r < 0) return r;
  if (r == XTAG_IMPL_TC)
    return tc_clip_fwd(A, Bm, M, N, D, scale, label_offset, row_lse, col_lse, diag, ws, ws_bytes, (cudaStream_t)stream);
  return simt_clip_fwd(A, Bm, dtype, M, N, D, scale, label_offset, row_lse, col_lse, diag, ws, ws_bytes,
                       (cudaStream_t)stream);
}

extern "C" int xtag_clip_bwd(const void* A, const void* Bm, int dtype, int M, int N, int D, const float* scale,
                             int label_offset, const float* row_lse, const float* col_lse, float w_row, float w_col,
                             float w_diag, const float* grad_out, void* dA, void* dB, int grad_dtype, float* dscale,
                             void* ws, size_t ws_bytes, int impl, int flags, void* stream) {
  XTAG_REQUIRE(A && Bm && scale && row_lse && col_lse && grad_out, XTAG_ERR_INVALID, "clip_bwd: null pointer");
  XTAG_REQUIRE(M > 0 && N > 0 && D > 0, XTAG_ERR_INVALID, "clip_bwd: empty problem M=%d N=%d D=%d", M, N, D);
  XTAG_REQUIRE((dtype == XTAG_F32 || dtype == XTAG_BF16) && (grad_dtype == XTAG_F32 || grad_dtype == XTAG_BF16),
               XTAG_ERR_INVALID, "clip_bwd: bad dtype");
  int rc = xtag_device_check();
  if (rc) return rc;
  const int r = resolve_impl(impl, dtype, D, "clip_bwd");
  if (r < 0) return r;
  if (r == XTAG_IMPL_TC)
    return tc_clip_bwd(A, Bm, M, N, D, scale, label_offset, row_lse, col_lse, w_row, w_col, w_diag, grad_out, dA, dB,
                       grad_dtype, dscale, ws, ws_bytes, flags, (cudaStream_t)stream);
  return simt_clip_bwd(A, Bm, dtype, M, N, D, scale, label_offset, row_lse, col_lse, w_row, w_col, w_diag, grad_out,
                       dA, dB, grad_dtype, dscale, ws, ws_bytes, flags, (cudaStream_t)stream);
}

// ---- forward with deferred reductions (one launch per arriving column block, two reductions per step) -----------
extern "C" int xtag_clip_fwd_block_parts(int M, int N, int* row_parts, int* col_parts) {
  XTAG_REQUIRE(M > 0 && N > 0 && row_parts && col_parts, XTAG_ERR_INVALID, "clip_fwd_block_parts: bad arguments");
  tc_fwd_block_parts(M, N, row_parts, col_parts);
  return XTAG_OK;
}

extern "C" int xtag_clip_fwd_block(const void* A, const void* Bm, int dtype, int M, int N, int D, const float* scale,
                                   int label_offset, float* row_part, float* col_part, int col_ld, float* diag,
                                   void* stream) {
  XTAG_REQUIRE(A && Bm && scale && row_part && col_part && diag, XTAG_ERR_INVALID, "clip_fwd_block: null pointer");
  XTAG_REQUIRE(M > 0 && N > 0 && D > 0 && col_ld >= N, XTAG_ERR_INVALID,
               "clip_fwd_block: bad shape M=%d N=%d D=%d col_ld=%d", M, N, D, col_ld);
  XTAG_REQUIRE(tc_eligible(dtype, D), XTAG_ERR_UNSUPPORTED,
               "clip_fwd_block: tcgen05 path only (bf16 inputs, D %% 8 == 0); use xtag_clip_fwd per block otherwise");
  XTAG_REQUIRE(label_offset == -1 || (label_offset >= 0 && (long)label_offset + M <= (long)N), XTAG_ERR_INVALID,
               "clip_fwd_block: labels [%d, %d) fall outside the %d columns", label_offset, label_offset + M, N);
  int rc = xtag_device_check();
  if (rc) return rc;
  return tc_clip_fwd_block(A, Bm, M, N, D, scale, label_offset, row_part, col_part, col_ld, diag, (cudaStream_t)stream);
}

extern "C" int xtag_lse_reduce_log2(const float* parts, int P, int n, float* out, void* stream) {
  XTAG_REQUIRE(parts && out && P > 0 && n > 0, XTAG_ERR_INVALID, "lse_reduce_log2: bad arguments");
  return launch_lse_reduce(parts, P, n, 1.f, kLn2, out, (cudaStream_t)stream);
}

// ---- streamed forward: ONE launch consumes the gather buffer block by block as the peers' features land ----------
extern "C" int xtag_clip_fwd_stream(const void* A, const void* Bm_all, int dtype, int M, int N, int D,
                                    const float* scale, int label_offset, const int* order_host,
                                    const int* wait_host, int nblk, int blk_cols, const int* ready_flags,
                                    const int* epoch, float* row_part, float* col_part, int col_ld, float* diag,
                                    void* stream) {
  XTAG_REQUIRE(A && Bm_all && scale && order_host && wait_host && row_part && col_part && diag, XTAG_ERR_INVALID,
               "clip_fwd_stream: null pointer");
  XTAG_REQUIRE(M > 0 && N > 0 && D > 0 && col_ld >= N, XTAG_ERR_INVALID, "clip_fwd_stream: bad shape");
  XTAG_REQUIRE(tc_eligible(dtype, D), XTAG_ERR_UNSUPPORTED, "clip_fwd_stream: tcgen05 path only (bf16, D %% 8 == 0)");
  XTAG_REQUIRE(label_offset >= 0 && (long)label_offset + M <= (long)N, XTAG_ERR_INVALID,
               "clip_fwd_stream: labels [%d, %d) fall outside the %d columns", label_offset, label_offset + M, N);
  bool any_wait = false;
  for (int k = 0; k < nblk && k < 16; ++k) any_wait |= wait_host[k] != 0;
  XTAG_REQUIRE(!any_wait || (ready_flags && epoch), XTAG_ERR_INVALID, "clip_fwd_stream: waiting blocks need flags");
  int rc = xtag_device_check();
  if (rc) return rc;
  return tc_clip_fwd_stream(A, Bm_all, M, N, D, scale, label_offset, order_host, wait_host, nblk, blk_cols, ready_flags,
                            epoch, row_part, col_part, col_ld, diag, (cudaStream_t)stream);
}

// ---- sigmoid (SigLIP) loss, sibling of the contrastive head on the same mainloop -----
extern "C" int xtag_siglip_fwd(const void* A, const void* Bm, int dtype, int M, int N, int D, const float* scale,
                               const float* bias, int label_offset, float weight, float* out3, void* ws, size_t ws_bytes,
                               int stage_ds, void* stream) {
  XTAG_REQUIRE(A && Bm && scale && bias && out3, XTAG_ERR_INVALID, "siglip_fwd: null pointer");
  XTAG_REQUIRE(M > 0 && N > 0 && D > 0, XTAG_ERR_INVALID, "siglip_fwd: empty problem M=%d N=%d D=%d", M, N, D);
  XTAG_REQUIRE(label_offset == -1 || (label_offset >= 0 && (long)label_offset + M <= (long)N), XTAG_ERR_INVALID,
               "siglip_fwd: labels [%d, %d) fall outside the %d columns", label_offset, label_offset + M, N);
  XTAG_REQUIRE(tc_eligible(dtype, D), XTAG_ERR_UNSUPPORTED,
               "siglip_fwd: tcgen05 path only (bf16 inputs, D %% 8 == 0)");
  int rc = xtag_device_check();
  if (rc) return rc;
  // label_offset == -1: a column block without positives -- every label is -1 (reference negative_only=True)
  return tc_siglip_fwd(A, Bm, M, N, D, scale, bias, label_offset < 0 ? -(1 << 30) : label_offset, weight, out3, ws,
                       ws_bytes, stage_ds, (cudaStream_t)stream);
}
