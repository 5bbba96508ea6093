// C-ABI entry points of the tag-head cross-attention core (K4).
#include "common.cuh"

namespace xtag {
template <typename T>
int xattn_simt_fwd(const void* q, const void* k, const void* v, void* o, float* lse, int b, int Lq, int Lk, int heads,
                   int dh, int ldq, int ldk, int ldv, float sm_scale, float p_drop, uint64_t seed, uint64_t offset,
                   cudaStream_t st);
template <typename T>
int xattn_simt_bwd(const void* q, const void* k, const void* v, const void* o, const void* d_o, const float* lse,
                   void* dq, void* dk, void* dv, int b, int Lq, int Lk, int heads, int dh, int ldq, int ldk, int ldv,
                   float sm_scale, float p_drop, uint64_t seed, uint64_t offset, cudaStream_t st);
bool xattn_mma_supported(int Lq, int dh);
int xattn_mma_fwd(const void* q, const void* k, const void* v, void* o, float* lse, int b, int Lq, int Lk, int heads,
                  int dh, int ldq, int ldk, int ldv, float sm_scale, float p_drop, uint64_t seed, uint64_t offset,
                  cudaStream_t st);
int xattn_mma_bwd(const void* q, const void* k, const void* v, const void* o, const void* d_o, const float* lse,
                  void* dq, void* dk, void* dv, float* delta_ws, int b, int Lq, int Lk, int heads, int dh, int ldq, int ldk,
                  int ldv, int lddk, int lddv, float sm_scale, float p_drop, uint64_t seed, uint64_t offset,
                  cudaStream_t st);
}  // namespace xtag

using namespace xtag;

// the tensor-core / TMA kernel needs bf16 operands that TMA can describe
static bool mma_eligible(const void* q, const void* k, const void* v, int dtype, int Lq, int dh, int heads, int ldq,
                         int ldk, int ldv) {
  auto al = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  return dtype == XTAG_BF16 && xattn_mma_supported(Lq, dh) && al(q) && al(k) && al(v) && ldq % 8 == 0 &&
         ldk % 8 == 0 && ldv % 8 == 0 && (heads * dh) % 8 == 0;
}

static int check_common(const char* who, int dtype, int b, int Lq, int Lk, int heads, int dh, int ldq, int ldk, int ldv,
                        float p) {
  XTAG_REQUIRE(dtype == XTAG_F32 || dtype == XTAG_BF16, XTAG_ERR_INVALID, "%s: bad dtype %d", who, dtype);
  XTAG_REQUIRE(b > 0 && Lq > 0 && Lk > 0 && heads > 0 && dh > 0, XTAG_ERR_INVALID, "%s: empty problem", who);
  XTAG_REQUIRE(ldq >= heads * dh && ldk >= heads * dh && ldv >= heads * dh, XTAG_ERR_INVALID,
               "%s: row strides smaller than heads*dh", who);
  XTAG_REQUIRE(p >= 0.f && p < 1.f, XTAG_ERR_INVALID, "%s: dropout_p must be in [0,1)", who);
  return xtag_device_check();
}

extern "C" int xtag_xattn_fwd(const void* q, const void* k, const void* v, int dtype, void* o, float* lse, int b, int Lq,
                              int Lk, int heads, int dh, int ldq, int ldk, int ldv, float softmax_scale, float dropout_p,
                              uint64_t seed, uint64_t offset, void* stream) {
  XTAG_REQUIRE(q && k && v && o && lse, XTAG_ERR_INVALID, "xattn_fwd: null pointer");
  int rc = check_common("xattn_fwd", dtype, b, Lq, Lk, heads, dh, ldq, ldk, ldv, dropout_p);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if (mma_eligible(q, k, v, dtype, Lq, dh, heads, ldq, ldk, ldv))
    return xattn_mma_fwd(q, k, v, o, lse, b, Lq, Lk, heads, dh, ldq, ldk, ldv, softmax_scale, dropout_p, seed, offset, st);
  if (dtype == XTAG_F32)
    return xattn_simt_fwd<float>(q, k, v, o, lse, b, Lq, Lk, heads, dh, ldq, ldk, ldv, softmax_scale, dropout_p, seed, offset, st);
  return xattn_simt_fwd<__nv_bfloat16>(q, k, v, o, lse, b, Lq, Lk, heads, dh, ldq, ldk, ldv, softmax_scale, dropout_p, seed,
                                       offset, st);
}

extern "C" int xtag_xattn_bwd(const void* q, const void* k, const void* v, const void* o, const void* d_o,
                              const float* lse, int dtype, void* dq, void* dk, void* dv, int b, int Lq, int Lk, int heads,
                              int dh, int ldq, int ldk, int ldv, float softmax_scale, float dropout_p, uint64_t seed,
                              uint64_t offset, void* ws, size_t ws_bytes, void* stream) {
  return xtag_xattn_bwd_ld(q, k, v, o, d_o, lse, dtype, dq, dk, dv, b, Lq, Lk, heads, dh, ldq, ldk, ldv, heads * dh,
                           heads * dh, softmax_scale, dropout_p, seed, offset, ws, ws_bytes, stream);
}

extern "C" int xtag_xattn_bwd_ld(const void* q, const void* k, const void* v, const void* o, const void* d_o,
                                 const float* lse, int dtype, void* dq, void* dk, void* dv, int b, int Lq, int Lk,
                                 int heads, int dh, int ldq, int ldk, int ldv, int lddk, int lddv, float softmax_scale,
                                 float dropout_p, uint64_t seed, uint64_t offset, void* ws, size_t ws_bytes,
                                 void* stream) {
  XTAG_REQUIRE(q && k && v && o && d_o && lse && dq && dk && dv, XTAG_ERR_INVALID, "xattn_bwd: null pointer");
  XTAG_REQUIRE(lddk >= heads * dh && lddv >= heads * dh, XTAG_ERR_INVALID, "xattn_bwd: dK / dV row strides too small");
  int rc = check_common("xattn_bwd", dtype, b, Lq, Lk, heads, dh, ldq, ldk, ldv, dropout_p);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  auto al = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  if (mma_eligible(q, k, v, dtype, Lq, dh, heads, ldq, ldk, ldv) && al(o) && al(d_o) && al(dq) && al(dk) && al(dv)) {
    XTAG_REQUIRE(ws && ws_bytes >= (size_t)b * heads * Lq * sizeof(float), XTAG_ERR_WORKSPACE,
                 "xattn_bwd: workspace of b*heads*Lq floats required");
    if (lddk % 8 == 0 && lddv % 8 == 0)
      return xattn_mma_bwd(q, k, v, o, d_o, lse, dq, dk, dv, (float*)ws, b, Lq, Lk, heads, dh, ldq, ldk, ldv, lddk, lddv,
                           softmax_scale, dropout_p, seed, offset, st);
  }
  XTAG_REQUIRE(lddk == heads * dh && lddv == heads * dh, XTAG_ERR_UNSUPPORTED,
               "xattn_bwd: strided dK / dV outputs are only supported by the bf16 tensor-core path");
  if (dtype == XTAG_F32)
    return xattn_simt_bwd<float>(q, k, v, o, d_o, lse, dq, dk, dv, b, Lq, Lk, heads, dh, ldq, ldk, ldv, softmax_scale,
                                 dropout_p, seed, offset, st);
  return xattn_simt_bwd<__nv_bfloat16>(q, k, v, o, d_o, lse, dq, dk, dv, b, Lq, Lk, heads, dh, ldq, ldk, ldv, softmax_scale,
                                       dropout_p, seed, offset, st);
}
