// K4, bf16 production path: flash-style cross-attention core of the XTag tag head.
// Reference: BertSelfAttention.forward cross branch (src/open_clip/tagging_heads/bert.py:219-274):
//   P = softmax(q k^T / sqrt(dh)) ; P = dropout(P) ; ctx = P v      per (sample, head); Lq = 44 tag queries,
//   Lk = 50 / 197 / 257 ViT tokens, 4 heads x 192.
//
// HBM-bound ((88 + 2 N) * 1536 bytes per sample-layer, ~36 flop/byte), so the design goal is to
// stream K and V exactly once at full bandwidth:
//   * one CTA per (sample, head), 16 query rows per warp (3 warps for Lq = 44), several CTAs resident per SM
//   * Q, K, V tiles arrive by TMA (cp.async.bulk.tensor.3d, 128B swizzle, zero-filled past Lq / Lk) into a
//     2-stage smem ring signalled through mbarriers; the next K/V tile is in flight while the current one is used
//   * QK^T and PV on tensor cores (mma.sync m16n8k16 bf16, fp32 accumulate; operands via ldmatrix on the swizzled
//     tiles); online softmax in the log2 domain with warp-shuffle row reductions; P never leaves registers
//   * Philox keep-mask for the attention-probability dropout (training), regenerated in the backward
// tcgen05 is deliberately not used here: M = 44 rows per problem cannot fill a 128-row UMMA tile and the kernel is
// bandwidth-bound, not tensor-bound.
#include <cuda.h>

#include "common.cuh"
#include "philox.cuh"
#include "tc_ptx.cuh"

namespace xtag {

using namespace ptx;

__device__ __forceinline__ void tma_load_3d(uint32_t dst_smem, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst_smem), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
// byte offset of element (row, col) inside a [rows][64 bf16] tile written by TMA with CU_TENSOR_MAP_SWIZZLE_128B
// (col must be a multiple of 8: one 16-byte unit)
__device__ __forceinline__ uint32_t swz(int row, int col) {
  return (uint32_t)(row * 128 + ((((col >> 3) ^ row) & 7) << 4));
}

constexpr int XA_KT = 32;        // keys per pipeline stage
constexpr int kTuneXattnFusedBwd = 0x800;   // xtag_set_tune bit 11: single-pass K4 backward
constexpr int XA_MAXW = 4;       // up to 64 query rows

// Dropout keep-bits for this thread's elements of a [16 rows x XA_KT keys] accumulator tile (query-major kernels):
// element (row r in {row0, row0+8}, n8 tile i, column 2*t4 + c) <-> bit (2*i + c) of bits[r].
// A Philox block covers 4 consecutive keys = the columns of two neighbouring lanes (t4, t4^1) of one n8 tile: the even
// lane computes the block of tile 2*ip, the odd lane the block of tile 2*ip + 1, and they swap keep-bits with one
// shuffle -- 4 elements per Philox call instead of 1.  key0 (first key of the tile) is a multiple of 4.
__device__ __forceinline__ void dropout_bits_tile(uint32_t (&bits)[2], uint64_t seed, uint64_t offset, uint64_t row_id0,
                                                  int key0, int kblocks, uint32_t thr, int t4) {
  const int odd = t4 & 1;
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    uint32_t acc = 0;
#pragma unroll
    for (int ip = 0; ip < XA_KT / 16; ++ip) {
      const int i_mine = 2 * ip + odd;                               // the n8 tile whose block this lane computes
      const int kblk = (key0 >> 2) + 2 * i_mine + (t4 >> 1);
      const uint32_t own = philox_keep4(seed, offset, row_id0 + (uint64_t)(r * 8), kblk, kblocks, thr);
      const uint32_t other = __shfl_xor_sync(0xffffffffu, own, 1);
      const uint32_t even_tile = odd ? other : own;                  // block of tile 2*ip   (computed by the even lane)
      const uint32_t odd_tile = odd ? own : other;                   // block of tile 2*ip+1 (computed by the odd lane)
      acc |= ((even_tile >> (2 * odd)) & 3u) << (2 * (2 * ip));
      acc |= ((odd_tile >> (2 * odd)) & 3u) << (2 * (2 * ip + 1));
    }
    bits[r] = acc;
  }
}

// NCH = dh / 64
template <int NCH>
__global__ void __launch_bounds__(32 * XA_MAXW)
xattn_fwd_mma_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                     const __grid_constant__ CUtensorMap tmV, __nv_bfloat16* __restrict__ o, float* __restrict__ lse,
                     int Lq, int Lk, int heads, float sl2, float p_drop, uint64_t seed, uint64_t offset) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int nw = blockDim.x >> 5;
  const int QROWS = 16 * nw;
  const uint32_t q_bytes = (uint32_t)NCH * QROWS * 128;
  constexpr uint32_t kv_chunk = XA_KT * 128;                  // one 64-column chunk of a K or V tile
  constexpr uint32_t stage_bytes = 2 * NCH * kv_chunk;        // K then V
  uint8_t* Qs = smem;
  uint8_t* KVs = smem + ((q_bytes + 1023) & ~1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(KVs + 2 * stage_bytes);   // [0] = Q, [1..2] = stages

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int bh = blockIdx.x, b = bh / heads, h = bh % heads;
  const int q0 = blockIdx.y * QROWS;                        // first query row of this CTA (query sets > 64 rows: one
  const int dh = NCH * 64;                                  // CTA per 64-row chunk, ONE launch for all of them)
  const int num_tiles = (Lk + XA_KT - 1) / XA_KT;

  if (tid == 0) {
    prefetch_tmap(&tmQ); prefetch_tmap(&tmK); prefetch_tmap(&tmV);
    mbar_init(smem_u32(&bars[0]), 1);
    mbar_init(smem_u32(&bars[1]), 1);
    mbar_init(smem_u32(&bars[2]), 1);
    fence_barrier_init();
  }
  __syncthreads();

  auto issue_tile = [&](int t) {
    const int st = t & 1;
    const uint32_t fb = smem_u32(&bars[1 + st]);
    const uint32_t base = smem_u32(KVs + st * stage_bytes);
    mbar_arrive_expect_tx(fb, stage_bytes);
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      tma_load_3d(base + c * kv_chunk, &tmK, fb, h * dh + c * 64, t * XA_KT, b);
      tma_load_3d(base + (NCH + c) * kv_chunk, &tmV, fb, h * dh + c * 64, t * XA_KT, b);
    }
  };
  if (tid == 0) {
    const uint32_t qb = smem_u32(&bars[0]);
    mbar_arrive_expect_tx(qb, q_bytes);
#pragma unroll
    for (int c = 0; c < NCH; ++c) tma_load_3d(smem_u32(Qs) + c * QROWS * 128, &tmQ, qb, h * dh + c * 64, q0, b);
    issue_tile(0);
  }

  const int g = lane >> 2, t4 = lane & 3;
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
  float oacc[NCH * 8][4];
#pragma unroll
  for (int i = 0; i < NCH * 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) oacc[i][j] = 0.f;

  mbar_wait(smem_u32(&bars[0]), 0);
  const uint32_t q_base = smem_u32(Qs);
  const int qrow = warp * 16 + (lane & 15);                 // ldmatrix row supplied by this lane (A operand)
  const int row0 = warp * 16 + g;                           // accumulator rows owned: row0 and row0 + 8
  const float keep_scale = p_drop > 0.f ? 1.f / (1.f - p_drop) : 1.f;
  const uint32_t drop_thr = philox_drop_threshold(p_drop);

  for (int t = 0; t < num_tiles; ++t) {
    const int st = t & 1;
    if (tid == 0 && t + 1 < num_tiles) issue_tile(t + 1);   // stage st^1 was released by the barrier below
    mbar_wait(smem_u32(&bars[1 + st]), (uint32_t)((t >> 1) & 1));
    const uint32_t k_base = smem_u32(KVs + st * stage_bytes);
    const uint32_t v_base = k_base + NCH * kv_chunk;

    // ---- S = Q K^T for this warp's 16 rows x XA_KT keys ----
    float sacc[XA_KT / 8][4];
#pragma unroll
    for (int i = 0; i < XA_KT / 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) sacc[i][j] = 0.f;
#pragma unroll
    for (int kk = 0; kk < NCH * 4; ++kk) {
      const int c = kk >> 2, kx = (kk & 3) * 16;
      uint32_t a[4];
      ldsm_x4(q_base + c * QROWS * 128 + swz(qrow, kx + ((lane >> 4) << 3)), a);
#pragma unroll
      for (int np = 0; np < XA_KT / 16; ++np) {
        // four 8x8 blocks: (keys 0-7, k 0-7), (keys 0-7, k 8-15), (keys 8-15, k 0-7), (keys 8-15, k 8-15)
        const int krow = np * 16 + ((lane >> 4) << 3) + (lane & 7);
        const int kcol = kx + (((lane >> 3) & 1) << 3);
        uint32_t bb[4];
        ldsm_x4(k_base + c * kv_chunk + swz(krow, kcol), bb);
        mma_bf16(sacc[2 * np], a, bb[0], bb[1]);
        mma_bf16(sacc[2 * np + 1], a, bb[2], bb[3]);
      }
    }
    // ---- scale (log2 domain), mask keys past Lk, online softmax ----
    const int key0 = t * XA_KT;
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int i = 0; i < XA_KT / 8; ++i) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int key = key0 + i * 8 + t4 * 2 + (j & 1);
        const float s = (key < Lk) ? sacc[i][j] * sl2 : -INFINITY;
        sacc[i][j] = s;
        mx[j >> 1] = fmaxf(mx[j >> 1], s);
      }
    }
    float corr[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
      const float m_new = fmaxf(m_run[r], mx[r]);            // every tile holds >= 1 valid key, so m_new is finite
      corr[r] = fast_exp2(m_run[r] - m_new);
      m_run[r] = m_new;
    }
    float rs[2] = {0.f, 0.f};
#pragma unroll
    for (int i = 0; i < XA_KT / 8; ++i) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float p = fast_exp2(sacc[i][j] - m_run[j >> 1]);
        rs[j >> 1] += p;
        sacc[i][j] = p;
      }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) l_run[r] = l_run[r] * corr[r] + rs[r];      // quad-partial sums, reduced at the end
    if (p_drop > 0.f) {
      // dropout acts on the normalised probabilities; the normaliser is linear, so masking the unnormalised
      // numerators (after the row sum above) is equivalent
      uint32_t keep[2];
      dropout_bits_tile(keep, seed, offset, (uint64_t)bh * Lq + q0 + row0, key0, (Lk + 3) >> 2, drop_thr, t4);
#pragma unroll
      for (int i = 0; i < XA_KT / 8; ++i) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          sacc[i][j] = ((keep[j >> 1] >> (2 * i + (j & 1))) & 1u) ? sacc[i][j] * keep_scale : 0.f;
      }
    }
#pragma unroll
    for (int i = 0; i < NCH * 8; ++i) {
      oacc[i][0] *= corr[0]; oacc[i][1] *= corr[0];
      oacc[i][2] *= corr[1]; oacc[i][3] *= corr[1];
    }
    // ---- O += P V ----
#pragma unroll
    for (int ks = 0; ks < XA_KT / 16; ++ks) {
      uint32_t a[4];
      a[0] = pack_bf16(sacc[2 * ks][0], sacc[2 * ks][1]);
      a[1] = pack_bf16(sacc[2 * ks][2], sacc[2 * ks][3]);
      a[2] = pack_bf16(sacc[2 * ks + 1][0], sacc[2 * ks + 1][1]);
      a[3] = pack_bf16(sacc[2 * ks + 1][2], sacc[2 * ks + 1][3]);
#pragma unroll
      for (int dn = 0; dn < NCH * 4; ++dn) {
        // transposed 8x8 blocks: (keys 0-7, d 0-7), (keys 8-15, d 0-7), (keys 0-7, d 8-15), (keys 8-15, d 8-15)
        const int vrow = ks * 16 + (((lane >> 3) & 1) << 3) + (lane & 7);
        const int vcol = dn * 16 + ((lane >> 4) << 3);
        uint32_t bb[4];
        ldsm_x4_t(v_base + (vcol >> 6) * kv_chunk + swz(vrow, vcol & 63), bb);
        mma_bf16(oacc[2 * dn], a, bb[0], bb[1]);
        mma_bf16(oacc[2 * dn + 1], a, bb[2], bb[3]);
      }
    }
    __syncthreads();                                         // stage st may be refilled from here on
  }

  // ---- finalize: normalise, store ctx (bf16) and the row log-sum-exp (natural log) ----
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
  }
  const int HD = heads * dh;
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int row = q0 + row0 + r * 8;
    if (row >= Lq) continue;
    const float inv = 1.f / l_run[r];
    __nv_bfloat16* orow = o + ((size_t)b * Lq + row) * HD + h * dh;
#pragma unroll
    for (int i = 0; i < NCH * 8; ++i)
      *reinterpret_cast<uint32_t*>(orow + i * 8 + t4 * 2) = pack_bf16(oacc[i][2 * r] * inv, oacc[i][2 * r + 1] * inv);
    if (t4 == 0) lse[(size_t)bh * Lq + row] = (m_run[r] + fast_log2(l_run[r])) * kLn2;
  }
}

// ---- host side ----------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled xa_get_encode() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

// [b][L][cols] bf16 tensor, row stride ld elements, batch stride L*ld; box = [1][box_rows][64]
static int xa_make_tmap(CUtensorMap* tm, const void* base, int b, int L, int cols, long ld, int box_rows) {
  PFN_encodeTiled enc = xa_get_encode();
  XTAG_REQUIRE(enc != nullptr, XTAG_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  XTAG_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0 && (ld * 2) % 16 == 0, XTAG_ERR_UNSUPPORTED,
               "xattn(mma): operands must be 16-byte aligned with 16-byte multiple row strides");
  cuuint64_t gdim[3] = {(cuuint64_t)cols, (cuuint64_t)L, (cuuint64_t)b};
  cuuint64_t gstride[2] = {(cuuint64_t)ld * 2, (cuuint64_t)L * (cuuint64_t)ld * 2};
  cuuint32_t box[3] = {64u, (cuuint32_t)box_rows, 1u};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  XTAG_REQUIRE(r == CUDA_SUCCESS, XTAG_ERR_CUDA, "cuTensorMapEncodeTiled(3d) failed with CUresult %d", (int)r);
  return XTAG_OK;
}

bool xattn_mma_supported(int Lq, int dh) { return (dh == 64 || dh == 128 || dh == 192 || dh == 256) && Lq >= 1; }
// warps per CTA / query rows per CTA for a query set of Lq rows (more than 64 rows: chunks of 64 on blockIdx.y)
static int xa_nw(int Lq) { const int nw = (Lq + 15) / 16; return nw > XA_MAXW ? XA_MAXW : nw; }

template <int NCH>
static int launch_fwd(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, void* o, float* lse, int b,
                      int Lq, int Lk, int heads, float sl2, float p, uint64_t seed, uint64_t offset, cudaStream_t st) {
  const int nw = xa_nw(Lq);
  const size_t q_bytes = ((size_t)NCH * 16 * nw * 128 + 1023) & ~(size_t)1023;
  const size_t smem = 1024 + q_bytes + 2 * (2 * NCH * XA_KT * 128) + 64;
  XTAG_CUDA(sync_spin_timeout());
  XTAG_CUDA(cudaFuncSetAttribute(xattn_fwd_mma_kernel<NCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const dim3 grid((unsigned)(b * heads), (unsigned)((Lq + 16 * nw - 1) / (16 * nw)));
  xattn_fwd_mma_kernel<NCH><<<grid, 32 * nw, smem, st>>>(tq, tk, tv, (__nv_bfloat16*)o, lse, Lq, Lk, heads, sl2, p, seed,
                                                         offset);
  XTAG_CHECK_LAUNCH();
  return XTAG_OK;
}

int xattn_mma_fwd(const void* q, const void* k, const void* v, void* o, float* lse, int b, int Lq, int Lk, int heads,
                  int dh, int ldq, int ldk, int ldv, float sm_scale, float p_drop, uint64_t seed, uint64_t offset,
                  cudaStream_t st) {
  CUtensorMap tq, tk, tv;
  const int nw = xa_nw(Lq);
  int rc = xa_make_tmap(&tq, q, b, Lq, heads * dh, ldq, 16 * nw);
  if (rc) return rc;
  rc = xa_make_tmap(&tk, k, b, Lk, heads * dh, ldk, XA_KT);
  if (rc) return rc;
  rc = xa_make_tmap(&tv, v, b, Lk, heads * dh, ldv, XA_KT);
  if (rc) return rc;
  const float sl2 = sm_scale * kLog2e;
  switch (dh / 64) {
    case 1: return launch_fwd<1>(tq, tk, tv, o, lse, b, Lq, Lk, heads, sl2, p_drop, seed, offset, st);
    case 2: return launch_fwd<2>(tq, tk, tv, o, lse, b, Lq, Lk, heads, sl2, p_drop, seed, offset, st);
    case 3: return launch_fwd<3>(tq, tk, tv, o, lse, b, Lq, Lk, heads, sl2, p_drop, seed, offset, st);
    default: return launch_fwd<4>(tq, tk, tv, o, lse, b, Lq, Lk, heads, sl2, p_drop, seed, offset, st);
  }
}



// delta_r = sum_c dO[r,c] * O[r,c] for this warp's 16 query rows.  O comes straight from global memory (it is needed
// nowhere else): all 16 row loads of a lane are issued back to back (one 16-byte unit per lane and row), so the warp
// pays ONE global round trip instead of sixteen dependent ones; dO is read from its TMA-swizzled smem tile.
// Phase 1 (before the wait on the Q/dO barrier) fills ov, phase 2 (after it) reduces.
template <int NCH>
__device__ __forceinline__ void delta_load_o(uint4 (&ov)[16], const __nv_bfloat16* __restrict__ o, int warp, int lane,
                                             int b, int h, int Lq, int HD, int q0 = 0) {
  const bool act = lane < NCH * 8;
#pragma unroll
  for (int rr = 0; rr < 16; ++rr) {
    const int row = q0 + warp * 16 + rr;
    ov[rr] = make_uint4(0u, 0u, 0u, 0u);
    if (act && row < Lq)
      ov[rr] = __ldg(reinterpret_cast<const uint4*>(o + ((size_t)b * Lq + row) * HD + h * (NCH * 64) + lane * 8));
  }
}
template <int NCH>
__device__ __forceinline__ void delta_reduce(const uint4 (&ov)[16], uint32_t do_base, int qrows, int warp, int lane,
                                             int Lq, float* delta_s, float* delta_g, int q0 = 0) {
  const bool act = lane < NCH * 8;
#pragma unroll
  for (int rr = 0; rr < 16; ++rr) {
    const int row = warp * 16 + rr;
    float sacc = 0.f;
    if (act) {
      uint32_t d0, d1, d2, d3;
      const uint32_t addr = do_base + (uint32_t)((lane >> 3) * qrows * 128) + swz(row, (lane & 7) * 8);
      asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(d0), "=r"(d1), "=r"(d2), "=r"(d3) : "r"(addr));
      const uint32_t dw[4] = {d0, d1, d2, d3};
      const uint32_t ow[4] = {ov[rr].x, ov[rr].y, ov[rr].z, ov[rr].w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&ow[k]));
        const float2 d2f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&dw[k]));
        sacc = fmaf(a.x, d2f.x, fmaf(a.y, d2f.y, sacc));
      }
    }
    sacc = warp_sum(sacc);
    if (lane == 0) {
      delta_s[row] = sacc;
      if (delta_g != nullptr && q0 + row < Lq) delta_g[q0 + row] = sacc;
    }
  }
  __syncwarp();
}

// =============================================================================================================
// Backward.  Two kernels, each a forward-shaped pipeline (TMA tiles, ldmatrix, mma.sync):
//   xattn_bwd_dq_kernel   query-major, one CTA per (sample, head): recompute S and dP = dO V^T tile by tile,
//                         dS = P (dP*mask - delta) * scale, dQ += dS K.  Also emits delta = rowsum(dO * O).
//   xattn_bwd_dkv_kernel  key-major, one CTA per (sample, head, 64-key tile), 16 keys per warp: S^T = K Q^T and
//                         dP^T = V dO^T over all Lq queries, dV = Pd^T dO, dK = dS^T Q -- complete per key tile, so
//                         there is no cross-CTA accumulation and no atomics.
// K and V are streamed twice (once per kernel): ~3x the forward's bytes in total, still bandwidth-shaped.
// =============================================================================================================
template <int NCH>
__global__ void __launch_bounds__(32 * XA_MAXW)
xattn_bwd_dq_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmDO,
                    const __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ d_o,
                    const float* __restrict__ lse, __nv_bfloat16* __restrict__ dq, float* __restrict__ delta,
                    int Lq, int Lk, int heads, float sl2, float sm_scale, float p_drop, uint64_t seed, uint64_t offset) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int nw = blockDim.x >> 5;
  const int QROWS = 16 * nw;
  const uint32_t q_bytes = (uint32_t)NCH * QROWS * 128;
  const uint32_t q_pad = (q_bytes + 1023) & ~1023u;
  constexpr uint32_t kv_chunk = XA_KT * 128;
  constexpr uint32_t stage_bytes = 2 * NCH * kv_chunk;
  uint8_t* Qs = smem;
  uint8_t* dOs = smem + q_pad;
  uint8_t* KVs = smem + 2 * q_pad;
  uint64_t* bars = reinterpret_cast<uint64_t*>(KVs + 2 * stage_bytes);
  float* delta_s = reinterpret_cast<float*>(bars + 4);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int bh = blockIdx.x, b = bh / heads, h = bh % heads;
  const int q0 = blockIdx.y * QROWS;                    // query chunk of this CTA (query sets of more than 64 rows)
  const int dh = NCH * 64, HD = heads * dh;
  const int num_tiles = (Lk + XA_KT - 1) / XA_KT;

  if (tid == 0) {
    prefetch_tmap(&tmQ); prefetch_tmap(&tmK); prefetch_tmap(&tmV); prefetch_tmap(&tmDO);
    mbar_init(smem_u32(&bars[0]), 1);
    mbar_init(smem_u32(&bars[1]), 1);
    mbar_init(smem_u32(&bars[2]), 1);
    fence_barrier_init();
  }
  __syncthreads();
  auto issue_tile = [&](int t) {
    const int st = t & 1;
    const uint32_t fb = smem_u32(&bars[1 + st]);
    const uint32_t base = smem_u32(KVs + st * stage_bytes);
    mbar_arrive_expect_tx(fb, stage_bytes);
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      tma_load_3d(base + c * kv_chunk, &tmK, fb, h * dh + c * 64, t * XA_KT, b);
      tma_load_3d(base + (NCH + c) * kv_chunk, &tmV, fb, h * dh + c * 64, t * XA_KT, b);
    }
  };
  if (tid == 0) {
    const uint32_t qb = smem_u32(&bars[0]);
    mbar_arrive_expect_tx(qb, 2 * q_bytes);
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      tma_load_3d(smem_u32(Qs) + c * QROWS * 128, &tmQ, qb, h * dh + c * 64, q0, b);
      tma_load_3d(smem_u32(dOs) + c * QROWS * 128, &tmDO, qb, h * dh + c * 64, q0, b);
    }
    issue_tile(0);
  }
  {
    uint4 ov[16];
    delta_load_o<NCH>(ov, o, warp, lane, b, h, Lq, HD, q0);
    mbar_wait(smem_u32(&bars[0]), 0);
    delta_reduce<NCH>(ov, smem_u32(dOs), QROWS, warp, lane, Lq, delta_s, delta + (size_t)bh * Lq, q0);
  }

  const int g = lane >> 2, t4 = lane & 3;
  const int row0 = warp * 16 + g;
  float lse2[2], dl[2];
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int row = row0 + r * 8;
    lse2[r] = (q0 + row < Lq) ? lse[(size_t)bh * Lq + q0 + row] * kLog2e : INFINITY;
    dl[r] = delta_s[row];
  }
  float qacc[NCH * 8][4];
#pragma unroll
  for (int i = 0; i < NCH * 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) qacc[i][j] = 0.f;

  mbar_wait(smem_u32(&bars[0]), 0);
  const uint32_t q_base = smem_u32(Qs), do_base = smem_u32(dOs);
  const int qrow = warp * 16 + (lane & 15);
  const float keep_scale = p_drop > 0.f ? 1.f / (1.f - p_drop) : 1.f;
  const uint32_t drop_thr = philox_drop_threshold(p_drop);

  for (int t = 0; t < num_tiles; ++t) {
    const int st = t & 1;
    if (tid == 0 && t + 1 < num_tiles) issue_tile(t + 1);
    mbar_wait(smem_u32(&bars[1 + st]), (uint32_t)((t >> 1) & 1));
    const uint32_t k_base = smem_u32(KVs + st * stage_bytes);
    const uint32_t v_base = k_base + NCH * kv_chunk;
    float sacc[XA_KT / 8][4], pacc[XA_KT / 8][4];
#pragma unroll
    for (int i = 0; i < XA_KT / 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) { sacc[i][j] = 0.f; pacc[i][j] = 0.f; }
#pragma unroll
    for (int kk = 0; kk < NCH * 4; ++kk) {
      const int c = kk >> 2, kx = (kk & 3) * 16;
      uint32_t aq[4], ad[4];
      ldsm_x4(q_base + c * QROWS * 128 + swz(qrow, kx + ((lane >> 4) << 3)), aq);
      ldsm_x4(do_base + c * QROWS * 128 + swz(qrow, kx + ((lane >> 4) << 3)), ad);
#pragma unroll
      for (int np = 0; np < XA_KT / 16; ++np) {
        const int krow = np * 16 + ((lane >> 4) << 3) + (lane & 7);
        const int kcol = kx + (((lane >> 3) & 1) << 3);
        uint32_t bk[4], bv[4];
        ldsm_x4(k_base + c * kv_chunk + swz(krow, kcol), bk);
        ldsm_x4(v_base + c * kv_chunk + swz(krow, kcol), bv);
        mma_bf16(sacc[2 * np], aq, bk[0], bk[1]);
        mma_bf16(sacc[2 * np + 1], aq, bk[2], bk[3]);
        mma_bf16(pacc[2 * np], ad, bv[0], bv[1]);
        mma_bf16(pacc[2 * np + 1], ad, bv[2], bv[3]);
      }
    }
    const int key0 = t * XA_KT;
    uint32_t keep[2] = {0xffffffffu, 0xffffffffu};
    if (p_drop > 0.f)
      dropout_bits_tile(keep, seed, offset, (uint64_t)bh * Lq + q0 + row0, key0, (Lk + 3) >> 2, drop_thr, t4);
#pragma unroll
    for (int i = 0; i < XA_KT / 8; ++i) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int key = key0 + i * 8 + t4 * 2 + (j & 1);
        const int r = j >> 1;
        float p = (key < Lk) ? fast_exp2(fmaf(sacc[i][j], sl2, -lse2[r])) : 0.f;
        const float m = ((keep[r] >> (2 * i + (j & 1))) & 1u) ? keep_scale : 0.f;
        sacc[i][j] = p * (pacc[i][j] * m - dl[r]) * sm_scale;        // dS (w.r.t. q.k before the 1/sqrt(dh) scale)
      }
    }
#pragma unroll
    for (int ks = 0; ks < XA_KT / 16; ++ks) {
      uint32_t a[4];
      a[0] = pack_bf16(sacc[2 * ks][0], sacc[2 * ks][1]);
      a[1] = pack_bf16(sacc[2 * ks][2], sacc[2 * ks][3]);
      a[2] = pack_bf16(sacc[2 * ks + 1][0], sacc[2 * ks + 1][1]);
      a[3] = pack_bf16(sacc[2 * ks + 1][2], sacc[2 * ks + 1][3]);
#pragma unroll
      for (int dn = 0; dn < NCH * 4; ++dn) {
        const int krow = ks * 16 + (((lane >> 3) & 1) << 3) + (lane & 7);
        const int kcol = dn * 16 + ((lane >> 4) << 3);
        uint32_t bb[4];
        ldsm_x4_t(k_base + (kcol >> 6) * kv_chunk + swz(krow, kcol & 63), bb);
        mma_bf16(qacc[2 * dn], a, bb[0], bb[1]);
        mma_bf16(qacc[2 * dn + 1], a, bb[2], bb[3]);
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int row = q0 + row0 + r * 8;
    if (row >= Lq) continue;
    __nv_bfloat16* orow = dq + ((size_t)b * Lq + row) * HD + h * dh;
#pragma unroll
    for (int i = 0; i < NCH * 8; ++i)
      *reinterpret_cast<uint32_t*>(orow + i * 8 + t4 * 2) = pack_bf16(qacc[i][2 * r], qacc[i][2 * r + 1]);
  }
}

constexpr int XA_KTB = 64;       // keys per CTA in the dK/dV kernel (4 warps x 16 keys)

template <int NCH, int NQT>      // NQT = number of 16-row query tiles (Lq <= 16 * NQT)
__global__ void __launch_bounds__(128)
xattn_bwd_dkv_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                     const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmDO,
                     const float* __restrict__ lse, const float* __restrict__ delta,
                     __nv_bfloat16* __restrict__ dk, __nv_bfloat16* __restrict__ dv,
                     int Lq, int Lk, int heads, int tiles_per_bh, float sl2, float sm_scale, float p_drop, uint64_t seed,
                     uint64_t offset) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  constexpr int QROWS = 16 * NQT;
  constexpr uint32_t q_bytes = (uint32_t)NCH * QROWS * 128;
  constexpr uint32_t q_pad = (q_bytes + 1023) & ~1023u;
  constexpr uint32_t kv_chunk = XA_KTB * 128;
  uint8_t* Qs = smem;
  uint8_t* dOs = smem + q_pad;
  uint8_t* Ks = smem + 2 * q_pad;
  uint8_t* Vs = Ks + NCH * kv_chunk;
  uint64_t* bars = reinterpret_cast<uint64_t*>(Vs + NCH * kv_chunk);
  float* lse2_s = reinterpret_cast<float*>(bars + 2);
  float* delta_s = lse2_s + QROWS;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int bh = blockIdx.x / tiles_per_bh, kt = blockIdx.x % tiles_per_bh;
  const int b = bh / heads, h = bh % heads;
  const int dh = NCH * 64, HD = heads * dh;
  const int key_base = kt * XA_KTB;

  if (tid == 0) {
    prefetch_tmap(&tmQ); prefetch_tmap(&tmK); prefetch_tmap(&tmV); prefetch_tmap(&tmDO);
    mbar_init(smem_u32(&bars[0]), 1);
    fence_barrier_init();
  }
  for (int i = tid; i < QROWS; i += blockDim.x) {
    lse2_s[i] = (i < Lq) ? lse[(size_t)bh * Lq + i] * kLog2e : INFINITY;      // +inf: padded queries get P = 0
    delta_s[i] = (i < Lq) ? delta[(size_t)bh * Lq + i] : 0.f;
  }
  __syncthreads();
  if (tid == 0) {
    const uint32_t fb = smem_u32(&bars[0]);
    mbar_arrive_expect_tx(fb, 2 * q_bytes + 2 * NCH * kv_chunk);
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      tma_load_3d(smem_u32(Qs) + c * QROWS * 128, &tmQ, fb, h * dh + c * 64, 0, b);
      tma_load_3d(smem_u32(dOs) + c * QROWS * 128, &tmDO, fb, h * dh + c * 64, 0, b);
      tma_load_3d(smem_u32(Ks) + c * kv_chunk, &tmK, fb, h * dh + c * 64, key_base, b);
      tma_load_3d(smem_u32(Vs) + c * kv_chunk, &tmV, fb, h * dh + c * 64, key_base, b);
    }
  }
  mbar_wait(smem_u32(&bars[0]), 0);

  const int g = lane >> 2, t4 = lane & 3;
  const uint32_t q_base = smem_u32(Qs), do_base = smem_u32(dOs), k_base = smem_u32(Ks), v_base = smem_u32(Vs);
  const int krow_a = warp * 16 + (lane & 15);                    // A-operand row (key) supplied by this lane
  const float keep_scale = p_drop > 0.f ? 1.f / (1.f - p_drop) : 1.f;
  const uint32_t drop_thr = philox_drop_threshold(p_drop);

  // S^T = K_w Q^T and dP^T = V_w dO^T : 16 keys x QROWS queries
  float st[2 * NQT][4], pt[2 * NQT][4];
#pragma unroll
  for (int i = 0; i < 2 * NQT; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) { st[i][j] = 0.f; pt[i][j] = 0.f; }
#pragma unroll
  for (int kk = 0; kk < NCH * 4; ++kk) {
    const int c = kk >> 2, kx = (kk & 3) * 16;
    uint32_t ak[4], av[4];
    ldsm_x4(k_base + c * kv_chunk + swz(krow_a, kx + ((lane >> 4) << 3)), ak);
    ldsm_x4(v_base + c * kv_chunk + swz(krow_a, kx + ((lane >> 4) << 3)), av);
#pragma unroll
    for (int np = 0; np < NQT; ++np) {
      const int qr = np * 16 + ((lane >> 4) << 3) + (lane & 7);
      const int qc = kx + (((lane >> 3) & 1) << 3);
      uint32_t bq[4], bd[4];
      ldsm_x4(q_base + c * QROWS * 128 + swz(qr, qc), bq);
      ldsm_x4(do_base + c * QROWS * 128 + swz(qr, qc), bd);
      mma_bf16(st[2 * np], ak, bq[0], bq[1]);
      mma_bf16(st[2 * np + 1], ak, bq[2], bq[3]);
      mma_bf16(pt[2 * np], av, bd[0], bd[1]);
      mma_bf16(pt[2 * np + 1], av, bd[2], bd[3]);
    }
  }
  // elementwise: rows = keys (g, g+8), columns = queries
#pragma unroll
  for (int i = 0; i < 2 * NQT; ++i) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int qi = i * 8 + t4 * 2 + (j & 1);
      const int key = key_base + warp * 16 + g + ((j >> 1) << 3);
      const float p = fast_exp2(fmaf(st[i][j], sl2, -lse2_s[qi]));
      float m = 1.f;
      if (p_drop > 0.f) {
        m = (qi < Lq && key < Lk && philox_keep(seed, offset, (uint64_t)bh * Lq + qi, key, (Lk + 3) >> 2, drop_thr))
                ? keep_scale : 0.f;
      }
      st[i][j] = p * (pt[i][j] * m - delta_s[qi]) * sm_scale;    // dS^T
      pt[i][j] = p * m;                                          // dropped probabilities Pd^T
    }
  }
  const int key_lo = key_base + warp * 16 + g;
  // two passes over the head dimension with one accumulator set: dV = Pd^T dO, then dK = dS^T Q
#pragma unroll 1
  for (int pass = 0; pass < 2; ++pass) {
    float acc[NCH * 8][4];
#pragma unroll
    for (int i = 0; i < NCH * 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    const uint32_t b_base = pass == 0 ? do_base : q_base;
#pragma unroll
    for (int ks = 0; ks < NQT; ++ks) {
      uint32_t a[4];
      if (pass == 0) {
        a[0] = pack_bf16(pt[2 * ks][0], pt[2 * ks][1]);         a[1] = pack_bf16(pt[2 * ks][2], pt[2 * ks][3]);
        a[2] = pack_bf16(pt[2 * ks + 1][0], pt[2 * ks + 1][1]); a[3] = pack_bf16(pt[2 * ks + 1][2], pt[2 * ks + 1][3]);
      } else {
        a[0] = pack_bf16(st[2 * ks][0], st[2 * ks][1]);         a[1] = pack_bf16(st[2 * ks][2], st[2 * ks][3]);
        a[2] = pack_bf16(st[2 * ks + 1][0], st[2 * ks + 1][1]); a[3] = pack_bf16(st[2 * ks + 1][2], st[2 * ks + 1][3]);
      }
#pragma unroll
      for (int dn = 0; dn < NCH * 4; ++dn) {
        const int qr = ks * 16 + (((lane >> 3) & 1) << 3) + (lane & 7);
        const int dc = dn * 16 + ((lane >> 4) << 3);
        uint32_t bb[4];
        ldsm_x4_t(b_base + (dc >> 6) * QROWS * 128 + swz(qr, dc & 63), bb);
        mma_bf16(acc[2 * dn], a, bb[0], bb[1]);
        mma_bf16(acc[2 * dn + 1], a, bb[2], bb[3]);
      }
    }
    __nv_bfloat16* outp = pass == 0 ? dv : dk;
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int key = key_lo + r * 8;
      if (key >= Lk) continue;
      __nv_bfloat16* orow = outp + ((size_t)b * Lk + key) * HD + h * dh;
#pragma unroll
      for (int i = 0; i < NCH * 8; ++i)
        *reinterpret_cast<uint32_t*>(orow + i * 8 + t4 * 2) = pack_bf16(acc[i][2 * r], acc[i][2 * r + 1]);
    }
  }
}

// =============================================================================================================
// dK / dV for query sets of MORE than 64 rows (the TQN fusion head attends with Lq = B queries per sample,
// reference CAR_heads/transformer_decoder.py:146-240 via model.py:552-561): one CTA per (sample, head, 64-key tile)
// keeps its K / V tile in shared memory and LOOPS over the queries in chunks of 64 rows (Q / dO chunk by TMA), so the
// whole backward of any Lq is two launches (this kernel + the chunked dQ kernel) and dK / dV are complete per CTA --
// no atomics, no per-chunk partial gradients to add up on the host.  Two sweeps over the chunks with ONE accumulator
// set (registers): sweep 0 accumulates dV = Pd^T dO, sweep 1 dK = dS^T Q.
template <int NCH>
__global__ void __launch_bounds__(128)
xattn_bwd_dkv_loop_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                          const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmDO,
                          const float* __restrict__ lse, const float* __restrict__ delta,
                          __nv_bfloat16* __restrict__ dk, __nv_bfloat16* __restrict__ dv,
                          int Lq, int Lk, int heads, int tiles_per_bh, float sl2, float sm_scale, float p_drop,
                          uint64_t seed, uint64_t offset) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  constexpr int NQT = 4, QROWS = 64;
  constexpr uint32_t q_bytes = (uint32_t)NCH * QROWS * 128;
  constexpr uint32_t kv_chunk = XA_KTB * 128;
  uint8_t* Qs = smem;
  uint8_t* dOs = smem + q_bytes;
  uint8_t* Ks = smem + 2 * q_bytes;
  uint8_t* Vs = Ks + NCH * kv_chunk;
  uint64_t* bars = reinterpret_cast<uint64_t*>(Vs + NCH * kv_chunk);      // [0] K/V tile, [1] Q/dO chunk
  float* lse2_s = reinterpret_cast<float*>(bars + 2);
  float* delta_s = lse2_s + QROWS;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int bh = blockIdx.x / tiles_per_bh, kt = blockIdx.x % tiles_per_bh;
  const int b = bh / heads, h = bh % heads;
  const int dh = NCH * 64, HD = heads * dh;
  const int key_base = kt * XA_KTB;
  const int nchunks = (Lq + QROWS - 1) / QROWS;

  if (tid == 0) {
    prefetch_tmap(&tmQ); prefetch_tmap(&tmK); prefetch_tmap(&tmV); prefetch_tmap(&tmDO);
    mbar_init(smem_u32(&bars[0]), 1);
    mbar_init(smem_u32(&bars[1]), 1);
    fence_barrier_init();
  }
  __syncthreads();
  if (tid == 0) {
    const uint32_t fb = smem_u32(&bars[0]);
    mbar_arrive_expect_tx(fb, 2 * NCH * kv_chunk);
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      tma_load_3d(smem_u32(Ks) + c * kv_chunk, &tmK, fb, h * dh + c * 64, key_base, b);
      tma_load_3d(smem_u32(Vs) + c * kv_chunk, &tmV, fb, h * dh + c * 64, key_base, b);
    }
  }
  mbar_wait(smem_u32(&bars[0]), 0);

  const int g = lane >> 2, t4 = lane & 3;
  const uint32_t q_base = smem_u32(Qs), do_base = smem_u32(dOs), k_base = smem_u32(Ks), v_base = smem_u32(Vs);
  const int krow_a = warp * 16 + (lane & 15);
  const float keep_scale = p_drop > 0.f ? 1.f / (1.f - p_drop) : 1.f;
  const uint32_t drop_thr = philox_drop_threshold(p_drop);
  const int key_lo = key_base + warp * 16 + g;
  uint32_t load_idx = 0;

#pragma unroll 1
  for (int sweep = 0; sweep < 2; ++sweep) {
    float acc[NCH * 8][4];
#pragma unroll
    for (int i = 0; i < NCH * 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
#pragma unroll 1
    for (int ch = 0; ch < nchunks; ++ch) {
      const int q0 = ch * QROWS;
      __syncthreads();                                   // every warp is done with the previous chunk's tiles
      if (tid == 0) {
        const uint32_t fb = smem_u32(&bars[1]);
        mbar_arrive_expect_tx(fb, 2 * q_bytes);
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
          tma_load_3d(q_base + c * QROWS * 128, &tmQ, fb, h * dh + c * 64, q0, b);
          tma_load_3d(do_base + c * QROWS * 128, &tmDO, fb, h * dh + c * 64, q0, b);
        }
      }
      for (int i = tid; i < QROWS; i += blockDim.x) {
        lse2_s[i] = (q0 + i < Lq) ? lse[(size_t)bh * Lq + q0 + i] * kLog2e : INFINITY;      // padded queries: P = 0
        delta_s[i] = (q0 + i < Lq) ? delta[(size_t)bh * Lq + q0 + i] : 0.f;
      }
      __syncthreads();
      mbar_wait(smem_u32(&bars[1]), load_idx & 1u);
      ++load_idx;

      // S^T = K_w Q^T (and, for dK, dP^T = V_w dO^T): 16 keys x 64 queries
      float st[2 * NQT][4], pt[2 * NQT][4];
#pragma unroll
      for (int i = 0; i < 2 * NQT; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { st[i][j] = 0.f; pt[i][j] = 0.f; }
#pragma unroll
      for (int kk = 0; kk < NCH * 4; ++kk) {
        const int c = kk >> 2, kx = (kk & 3) * 16;
        uint32_t ak[4], av[4];
        ldsm_x4(k_base + c * kv_chunk + swz(krow_a, kx + ((lane >> 4) << 3)), ak);
        if (sweep == 1) ldsm_x4(v_base + c * kv_chunk + swz(krow_a, kx + ((lane >> 4) << 3)), av);
#pragma unroll
        for (int np = 0; np < NQT; ++np) {
          const int qr = np * 16 + ((lane >> 4) << 3) + (lane & 7);
          const int qc = kx + (((lane >> 3) & 1) << 3);
          uint32_t bq[4];
          ldsm_x4(q_base + c * QROWS * 128 + swz(qr, qc), bq);
          mma_bf16(st[2 * np], ak, bq[0], bq[1]);
          mma_bf16(st[2 * np + 1], ak, bq[2], bq[3]);
          if (sweep == 1) {
            uint32_t bd[4];
            ldsm_x4(do_base + c * QROWS * 128 + swz(qr, qc), bd);
            mma_bf16(pt[2 * np], av, bd[0], bd[1]);
            mma_bf16(pt[2 * np + 1], av, bd[2], bd[3]);
          }
        }
      }
      // elementwise: rows = keys (g, g+8), columns = queries; st <- the A operand of this sweep's product
#pragma unroll
      for (int i = 0; i < 2 * NQT; ++i) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int qi = i * 8 + t4 * 2 + (j & 1);
          const int key = key_base + warp * 16 + g + ((j >> 1) << 3);
          const float p = fast_exp2(fmaf(st[i][j], sl2, -lse2_s[qi]));
          float m = 1.f;
          if (p_drop > 0.f) {
            m = (q0 + qi < Lq && key < Lk &&
                 philox_keep(seed, offset, (uint64_t)bh * Lq + q0 + qi, key, (Lk + 3) >> 2, drop_thr))
                    ? keep_scale : 0.f;
          }
          st[i][j] = (sweep == 0) ? p * m                                          // Pd^T
                                  : p * (pt[i][j] * m - delta_s[qi]) * sm_scale;   // dS^T
        }
      }
      const uint32_t b_base = sweep == 0 ? do_base : q_base;
#pragma unroll
      for (int ks = 0; ks < NQT; ++ks) {
        uint32_t a[4];
        a[0] = pack_bf16(st[2 * ks][0], st[2 * ks][1]);         a[1] = pack_bf16(st[2 * ks][2], st[2 * ks][3]);
        a[2] = pack_bf16(st[2 * ks + 1][0], st[2 * ks + 1][1]); a[3] = pack_bf16(st[2 * ks + 1][2], st[2 * ks + 1][3]);
#pragma unroll
        for (int dn = 0; dn < NCH * 4; ++dn) {
          const int qr = ks * 16 + (((lane >> 3) & 1) << 3) + (lane & 7);
          const int dc = dn * 16 + ((lane >> 4) << 3);
          uint32_t bb[4];
          ldsm_x4_t(b_base + (dc >> 6) * QROWS * 128 + swz(qr, dc & 63), bb);
          mma_bf16(acc[2 * dn], a, bb[0], bb[1]);
          mma_bf16(acc[2 * dn + 1], a, bb[2], bb[3]);
        }
      }
    }
    __nv_bfloat16* outp = sweep == 0 ? dv : dk;
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int key = key_lo + r * 8;
      if (key >= Lk) continue;
      __nv_bfloat16* orow = outp + ((size_t)b * Lk + key) * HD + h * dh;
#pragma unroll
      for (int i = 0; i < NCH * 8; ++i)
        *reinterpret_cast<uint32_t*>(orow + i * 8 + t4 * 2) = pack_bf16(acc[i][2 * r], acc[i][2 * r + 1]);
    }
  }
}

// =============================================================================================================
// Backward, single pass (default).  One CTA per (sample, head), NW = ceil(Lq / 16) warps, K and V streamed ONCE:
// per 32-key tile
//   S = Q K^T, dP = dO V^T                         16 query rows per warp (as the forward)
//   P = 2^(S*sl2 - lse2), Pd = P*mask, dS = P (dP*mask - delta) * scale
//   dQ += dS K                                      A operand straight from the accumulator registers
//   Pd, dS -> shared memory (bf16, [query][key])    so that every warp sees all Lq queries of the tile
//   dV_tile = Pd^T dO, dK_tile = dS^T Q             contraction over the queries; the head dimension is split across
//                                                   the warps (A = ldmatrix.trans of the staged Pd / dS)
//   dV_tile / dK_tile -> swizzled staging tile -> TMA tile store (full 128-byte rows, rows past Lk clipped)
// HBM traffic is the algorithmic minimum: q, k, v, o, dO read once, dq, dk, dv written once.
// =============================================================================================================
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, uint32_t src_smem, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(src_smem), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
// byte offset of (row, 16-byte unit) in a [rows][32 keys] bf16 staging tile (64-byte rows, XOR swizzle so that
// both the fragment stores and the transposed ldmatrix reads are bank-conflict free)
__device__ __forceinline__ uint32_t swz64(int row, int unit) {
  return (uint32_t)(row * 64 + ((unit ^ ((row >> 1) & 3)) << 4));
}

template <int NCH, int NW>
__global__ void __launch_bounds__(32 * NW)
xattn_bwd_fused_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                       const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmDO,
                       const __grid_constant__ CUtensorMap tmDK, const __grid_constant__ CUtensorMap tmDV,
                       const __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ d_o,
                       const float* __restrict__ lse, __nv_bfloat16* __restrict__ dq,
                       int Lq, int Lk, int heads, float sl2, float sm_scale, float p_drop, uint64_t seed,
                       uint64_t offset) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  constexpr int QROWS = 16 * NW;
  constexpr uint32_t q_bytes = (uint32_t)NCH * QROWS * 128;
  constexpr uint32_t q_pad = (q_bytes + 1023) & ~1023u;
  constexpr uint32_t kv_chunk = XA_KT * 128;
  constexpr uint32_t stage_bytes = 2 * NCH * kv_chunk;
  constexpr int NPAIR = NCH * 4;                       // 16-column pairs of n8 tiles across the head dimension
  constexpr int NP = (NPAIR + NW - 1) / NW;            // pairs owned by one warp in the dV / dK products
  uint8_t* Qs = smem;
  uint8_t* dOs = smem + q_pad;
  uint8_t* KVs = smem + 2 * q_pad;
  uint8_t* Os = KVs + 2 * stage_bytes;                 // [NCH][32][64] bf16 output staging (TMA 128B swizzle)
  uint8_t* Ps = Os + NCH * kv_chunk;                   // [QROWS][32] bf16 dropped probabilities
  uint8_t* dSs = Ps + QROWS * 64;                      // [QROWS][32] bf16 dS
  uint64_t* bars = reinterpret_cast<uint64_t*>(dSs + QROWS * 64);
  float* delta_s = reinterpret_cast<float*>(bars + 4);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int bh = blockIdx.x, b = bh / heads, h = bh % heads;
  constexpr int dh = NCH * 64;
  const int HD = heads * dh;
  const int num_tiles = (Lk + XA_KT - 1) / XA_KT;

  if (tid == 0) {
    prefetch_tmap(&tmQ); prefetch_tmap(&tmK); prefetch_tmap(&tmV); prefetch_tmap(&tmDO);
    prefetch_tmap(&tmDK); prefetch_tmap(&tmDV);
    mbar_init(smem_u32(&bars[0]), 1);
    mbar_init(smem_u32(&bars[1]), 1);
    mbar_init(smem_u32(&bars[2]), 1);
    fence_barrier_init();
  }
  __syncthreads();
  auto issue_tile = [&](int t) {
    const int st = t & 1;
    const uint32_t fb = smem_u32(&bars[1 + st]);
    const uint32_t base = smem_u32(KVs + st * stage_bytes);
    mbar_arrive_expect_tx(fb, stage_bytes);
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      tma_load_3d(base + c * kv_chunk, &tmK, fb, h * dh + c * 64, t * XA_KT, b);
      tma_load_3d(base + (NCH + c) * kv_chunk, &tmV, fb, h * dh + c * 64, t * XA_KT, b);
    }
  };
  if (tid == 0) {
    const uint32_t qb = smem_u32(&bars[0]);
    mbar_arrive_expect_tx(qb, 2 * q_bytes);
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      tma_load_3d(smem_u32(Qs) + c * QROWS * 128, &tmQ, qb, h * dh + c * 64, 0, b);
      tma_load_3d(smem_u32(dOs) + c * QROWS * 128, &tmDO, qb, h * dh + c * 64, 0, b);
    }
    issue_tile(0);
  }
  {
    uint4 ov[16];
    delta_load_o<NCH>(ov, o, warp, lane, b, h, Lq, HD);
    mbar_wait(smem_u32(&bars[0]), 0);
    delta_reduce<NCH>(ov, smem_u32(dOs), QROWS, warp, lane, Lq, delta_s, nullptr);
  }

  const int g = lane >> 2, t4 = lane & 3;
  const int row0 = warp * 16 + g;
  float lse2[2], dl[2];
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int row = row0 + r * 8;
    lse2[r] = (row < Lq) ? lse[(size_t)bh * Lq + row] * kLog2e : INFINITY;    // padded query rows: P = 0
    dl[r] = delta_s[row];
  }
  float qacc[NCH * 8][4];
#pragma unroll
  for (int i = 0; i < NCH * 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) qacc[i][j] = 0.f;

  mbar_wait(smem_u32(&bars[0]), 0);
  const uint32_t q_base = smem_u32(Qs), do_base = smem_u32(dOs);
  const uint32_t ps_base = smem_u32(Ps), ds_base = smem_u32(dSs), os_base = smem_u32(Os);
  const int qrow = warp * 16 + (lane & 15);
  const float keep_scale = p_drop > 0.f ? 1.f / (1.f - p_drop) : 1.f;
  const uint32_t drop_thr = philox_drop_threshold(p_drop);

  for (int t = 0; t < num_tiles; ++t) {
    const int st = t & 1;
    if (tid == 0 && t + 1 < num_tiles) issue_tile(t + 1);      // stage st^1 was released by the last barrier of t-1
    mbar_wait(smem_u32(&bars[1 + st]), (uint32_t)((t >> 1) & 1));
    const uint32_t k_base = smem_u32(KVs + st * stage_bytes);
    const uint32_t v_base = k_base + NCH * kv_chunk;
    const int key0 = t * XA_KT;
    // ---- S = Q K^T and dP = dO V^T for this warp's 16 query rows ----
    float sacc[XA_KT / 8][4], pacc[XA_KT / 8][4];
#pragma unroll
    for (int i = 0; i < XA_KT / 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) { sacc[i][j] = 0.f; pacc[i][j] = 0.f; }
#pragma unroll
    for (int kk = 0; kk < NCH * 4; ++kk) {
      const int c = kk >> 2, kx = (kk & 3) * 16;
      uint32_t aq[4], ad[4];
      ldsm_x4(q_base + c * QROWS * 128 + swz(qrow, kx + ((lane >> 4) << 3)), aq);
      ldsm_x4(do_base + c * QROWS * 128 + swz(qrow, kx + ((lane >> 4) << 3)), ad);
#pragma unroll
      for (int np = 0; np < XA_KT / 16; ++np) {
        const int krow = np * 16 + ((lane >> 4) << 3) + (lane & 7);
        const int kcol = kx + (((lane >> 3) & 1) << 3);
        uint32_t bk[4], bv[4];
        ldsm_x4(k_base + c * kv_chunk + swz(krow, kcol), bk);
        ldsm_x4(v_base + c * kv_chunk + swz(krow, kcol), bv);
        mma_bf16(sacc[2 * np], aq, bk[0], bk[1]);
        mma_bf16(sacc[2 * np + 1], aq, bk[2], bk[3]);
        mma_bf16(pacc[2 * np], ad, bv[0], bv[1]);
        mma_bf16(pacc[2 * np + 1], ad, bv[2], bv[3]);
      }
    }
    // ---- elementwise: sacc <- dS, pacc <- dropped probabilities ----
    uint32_t keep[2] = {0xffffffffu, 0xffffffffu};
    if (p_drop > 0.f) dropout_bits_tile(keep, seed, offset, (uint64_t)bh * Lq + row0, key0, (Lk + 3) >> 2, drop_thr, t4);
#pragma unroll
    for (int i = 0; i < XA_KT / 8; ++i) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int key = key0 + i * 8 + t4 * 2 + (j & 1);
        const int r = j >> 1;
        const float p = (key < Lk) ? fast_exp2(fmaf(sacc[i][j], sl2, -lse2[r])) : 0.f;
        const float m = ((keep[r] >> (2 * i + (j & 1))) & 1u) ? keep_scale : 0.f;
        sacc[i][j] = p * (pacc[i][j] * m - dl[r]) * sm_scale;    // dS (w.r.t. q.k before the 1/sqrt(dh) scale)
        pacc[i][j] = p * m;
      }
    }
    // ---- stage Pd and dS ([query][key], bf16) for the key-major products ----
#pragma unroll
    for (int i = 0; i < XA_KT / 8; ++i) {
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int row = row0 + r * 8;
        const uint32_t off = swz64(row, i) + (uint32_t)(t4 * 4);
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(ps_base + off), "r"(pack_bf16(pacc[i][2 * r], pacc[i][2 * r + 1]))
                     : "memory");
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(ds_base + off), "r"(pack_bf16(sacc[i][2 * r], sacc[i][2 * r + 1]))
                     : "memory");
      }
    }
    // ---- dQ += dS K (A operand from the accumulator registers) ----
#pragma unroll
    for (int ks = 0; ks < XA_KT / 16; ++ks) {
      uint32_t a[4];
      a[0] = pack_bf16(sacc[2 * ks][0], sacc[2 * ks][1]);
      a[1] = pack_bf16(sacc[2 * ks][2], sacc[2 * ks][3]);
      a[2] = pack_bf16(sacc[2 * ks + 1][0], sacc[2 * ks + 1][1]);
      a[3] = pack_bf16(sacc[2 * ks + 1][2], sacc[2 * ks + 1][3]);
#pragma unroll
      for (int dn = 0; dn < NCH * 4; ++dn) {
        const int krow = ks * 16 + (((lane >> 3) & 1) << 3) + (lane & 7);
        const int kcol = dn * 16 + ((lane >> 4) << 3);
        uint32_t bb[4];
        ldsm_x4_t(k_base + (kcol >> 6) * kv_chunk + swz(krow, kcol & 63), bb);
        mma_bf16(qacc[2 * dn], a, bb[0], bb[1]);
        mma_bf16(qacc[2 * dn + 1], a, bb[2], bb[3]);
      }
    }
    // the previous tile's dK store must have drained the staging tile before it is overwritten below
    if (tid == 0) tma_store_wait_read();
    __syncthreads();                                             // (E) Pd / dS of all warps staged; Os free

    // ---- key-major products: pass 0 dV_tile = Pd^T dO, pass 1 dK_tile = dS^T Q ----
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
      const uint32_t a_base = pass == 0 ? ps_base : ds_base;
      const uint32_t b_base = pass == 0 ? do_base : q_base;
      float acc[2][2 * NP][4];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int i = 0; i < 2 * NP; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[mt][i][j] = 0.f;
#pragma unroll
      for (int ks = 0; ks < NW; ++ks) {
        uint32_t a[2][4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          // A[key][query] = staged[query][key]: four transposed 8x8 blocks (keys 0-7 | 8-15) x (queries 0-7 | 8-15)
          const int mat = lane >> 3;
          const int qr = ks * 16 + ((mat >> 1) << 3) + (lane & 7);
          ldsm_x4_t(a_base + swz64(qr, mt * 2 + (mat & 1)), a[mt]);
        }
#pragma unroll
        for (int pl = 0; pl < NP; ++pl) {
          const int dn = warp * NP + pl;
          if (dn < NPAIR) {
            const int qr = ks * 16 + (((lane >> 3) & 1) << 3) + (lane & 7);
            const int dc = dn * 16 + ((lane >> 4) << 3);
            uint32_t bb[4];
            ldsm_x4_t(b_base + (dc >> 6) * QROWS * 128 + swz(qr, dc & 63), bb);
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
              mma_bf16(acc[mt][2 * pl], a[mt], bb[0], bb[1]);
              mma_bf16(acc[mt][2 * pl + 1], a[mt], bb[2], bb[3]);
            }
          }
        }
      }
      if (pass == 1) {
        if (tid == 0) tma_store_wait_read();                     // the dV store has drained the staging tile
        __syncthreads();
      }
      // accumulators -> staging tile in the TMA 128B-swizzle layout [NCH][32 keys][64]
#pragma unroll
      for (int pl = 0; pl < NP; ++pl) {
        const int dn = warp * NP + pl;
        if (dn < NPAIR) {
#pragma unroll
          for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nn = 0; nn < 2; ++nn)
#pragma unroll
              for (int r = 0; r < 2; ++r) {
                const int row = mt * 16 + g + r * 8;
                const int col = dn * 16 + nn * 8;
                const uint32_t off = (uint32_t)((col >> 6) * kv_chunk) + swz(row, col & 63) + (uint32_t)(t4 * 4);
                asm volatile("st.shared.b32 [%0], %1;" ::"r"(os_base + off),
                             "r"(pack_bf16(acc[mt][2 * pl + nn][2 * r], acc[mt][2 * pl + nn][2 * r + 1]))
                             : "memory");
              }
        }
      }
      fence_proxy_async();                                       // generic-proxy writes -> visible to the TMA engine
      __syncthreads();
      if (tid == 0) {
        const CUtensorMap* tmO = pass == 0 ? &tmDV : &tmDK;
#pragma unroll
        for (int c = 0; c < NCH; ++c) tma_store_3d(tmO, os_base + c * kv_chunk, h * dh + c * 64, key0, b);
        tma_store_commit();
      }
    }
    // the barrier inside pass 1 ordered every warp's reads of stage st, Ps and dSs (pass 0 and the S / dP / dQ
    // products) before this point only for pass 0; close the tile for pass 1's operand reads too
    __syncthreads();
  }
  if (tid == 0) tma_store_wait_all();
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int row = row0 + r * 8;
    if (row >= Lq) continue;
    __nv_bfloat16* orow = dq + ((size_t)b * Lq + row) * HD + h * dh;
#pragma unroll
    for (int i = 0; i < NCH * 8; ++i)
      *reinterpret_cast<uint32_t*>(orow + i * 8 + t4 * 2) = pack_bf16(qacc[i][2 * r], qacc[i][2 * r + 1]);
  }
}

template <int NCH, int NW>
static int launch_bwd_fused(const void* q, const void* k, const void* v, const void* o, const void* d_o,
                            const float* lse, void* dq, void* dk, void* dv, int b, int Lq, int Lk, int heads, int ldq,
                            int ldk, int ldv, int lddk, int lddv, float sm_scale, float p, uint64_t seed, uint64_t offset,
                            cudaStream_t st) {
  const int dh = NCH * 64, HD = heads * dh;
  const float sl2 = sm_scale * kLog2e;
  CUtensorMap tq, tdo, tk, tv, tdk, tdv;
  int rc;
  if ((rc = xa_make_tmap(&tq, q, b, Lq, HD, ldq, 16 * NW))) return rc;
  if ((rc = xa_make_tmap(&tdo, d_o, b, Lq, HD, HD, 16 * NW))) return rc;
  if ((rc = xa_make_tmap(&tk, k, b, Lk, HD, ldk, XA_KT))) return rc;
  if ((rc = xa_make_tmap(&tv, v, b, Lk, HD, ldv, XA_KT))) return rc;
  if ((rc = xa_make_tmap(&tdk, dk, b, Lk, HD, lddk, XA_KT))) return rc;      // dK / dV may be column slices of a wider
  if ((rc = xa_make_tmap(&tdv, dv, b, Lk, HD, lddv, XA_KT))) return rc;      // buffer (row stride lddk / lddv)
  const size_t q_pad = ((size_t)NCH * 16 * NW * 128 + 1023) & ~(size_t)1023;
  const size_t smem = 1024 + 2 * q_pad + 2 * (2 * NCH * XA_KT * 128) + NCH * XA_KT * 128 + 2 * (16 * NW * 64) + 64 +
                      16 * NW * 4;
  XTAG_CUDA(sync_spin_timeout());
  XTAG_CUDA(cudaFuncSetAttribute(xattn_bwd_fused_kernel<NCH, NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  xattn_bwd_fused_kernel<NCH, NW><<<b * heads, 32 * NW, smem, st>>>(
      tq, tk, tv, tdo, tdk, tdv, (const __nv_bfloat16*)o, (const __nv_bfloat16*)d_o, lse, (__nv_bfloat16*)dq, Lq, Lk,
      heads, sl2, sm_scale, p, seed, offset);
  XTAG_CHECK_LAUNCH();
  return XTAG_OK;
}

template <int NCH>
static int launch_bwd_fused_nch(int nw, const void* q, const void* k, const void* v, const void* o, const void* d_o,
                                const float* lse, void* dq, void* dk, void* dv, int b, int Lq, int Lk, int heads,
                                int ldq, int ldk, int ldv, int lddk, int lddv, float sm_scale, float p, uint64_t seed,
                                uint64_t offset, cudaStream_t st) {
  switch (nw) {
    case 1: return launch_bwd_fused<NCH, 1>(q, k, v, o, d_o, lse, dq, dk, dv, b, Lq, Lk, heads, ldq, ldk, ldv, lddk, lddv, sm_scale, p, seed, offset, st);
    case 2: return launch_bwd_fused<NCH, 2>(q, k, v, o, d_o, lse, dq, dk, dv, b, Lq, Lk, heads, ldq, ldk, ldv, lddk, lddv, sm_scale, p, seed, offset, st);
    case 3: return launch_bwd_fused<NCH, 3>(q, k, v, o, d_o, lse, dq, dk, dv, b, Lq, Lk, heads, ldq, ldk, ldv, lddk, lddv, sm_scale, p, seed, offset, st);
    default: return launch_bwd_fused<NCH, 4>(q, k, v, o, d_o, lse, dq, dk, dv, b, Lq, Lk, heads, ldq, ldk, ldv, lddk, lddv, sm_scale, p, seed, offset, st);
  }
}

template <int NCH>
static int launch_bwd(const void* q, const void* k, const void* v, const void* o, const void* d_o, const float* lse,
                      void* dq, void* dk, void* dv, float* delta, int b, int Lq, int Lk, int heads, int ldq, int ldk,
                      int ldv, float sm_scale, float p, uint64_t seed, uint64_t offset, cudaStream_t st) {
  const int dh = NCH * 64, HD = heads * dh;
  const int nw = (Lq + 15) / 16;
  const float sl2 = sm_scale * kLog2e;
  CUtensorMap tq, tdo, tk, tv, tk2, tv2;
  int rc;
  if ((rc = xa_make_tmap(&tq, q, b, Lq, HD, ldq, 16 * nw))) return rc;
  if ((rc = xa_make_tmap(&tdo, d_o, b, Lq, HD, HD, 16 * nw))) return rc;
  if ((rc = xa_make_tmap(&tk, k, b, Lk, HD, ldk, XA_KT))) return rc;
  if ((rc = xa_make_tmap(&tv, v, b, Lk, HD, ldv, XA_KT))) return rc;
  if ((rc = xa_make_tmap(&tk2, k, b, Lk, HD, ldk, XA_KTB))) return rc;
  if ((rc = xa_make_tmap(&tv2, v, b, Lk, HD, ldv, XA_KTB))) return rc;
  {
    const size_t q_pad = ((size_t)NCH * 16 * nw * 128 + 1023) & ~(size_t)1023;
    const size_t smem = 1024 + 2 * q_pad + 2 * (2 * NCH * XA_KT * 128) + 64 + 16 * nw * 4;
    XTAG_CUDA(sync_spin_timeout());
    XTAG_CUDA(cudaFuncSetAttribute(xattn_bwd_dq_kernel<NCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    xattn_bwd_dq_kernel<NCH><<<b * heads, 32 * nw, smem, st>>>(tq, tk, tv, tdo, (const __nv_bfloat16*)o,
                                                               (const __nv_bfloat16*)d_o, lse, (__nv_bfloat16*)dq, delta,
                                                               Lq, Lk, heads, sl2, sm_scale, p, seed, offset);
    XTAG_CHECK_LAUNCH();
  }
  {
    const int tiles = (Lk + XA_KTB - 1) / XA_KTB;
    auto launch = [&](auto kern, int nqt) -> int {
      const size_t q_pad = ((size_t)NCH * 16 * nqt * 128 + 1023) & ~(size_t)1023;
      const size_t smem = 1024 + 2 * q_pad + 2 * NCH * XA_KTB * 128 + 64 + 2 * 16 * nqt * 4;
      XTAG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      kern<<<b * heads * tiles, 128, smem, st>>>(tq, tk2, tv2, tdo, lse, delta, (__nv_bfloat16*)dk, (__nv_bfloat16*)dv, Lq,
                                                 Lk, heads, tiles, sl2, sm_scale, p, seed, offset);
      XTAG_CHECK_LAUNCH();
      return XTAG_OK;
    };
    switch (nw) {
      case 1: rc = launch(xattn_bwd_dkv_kernel<NCH, 1>, 1); break;
      case 2: rc = launch(xattn_bwd_dkv_kernel<NCH, 2>, 2); break;
      case 3: rc = launch(xattn_bwd_dkv_kernel<NCH, 3>, 3); break;
      default: rc = launch(xattn_bwd_dkv_kernel<NCH, 4>, 4); break;
    }
    if (rc) return rc;
  }
  return XTAG_OK;
}

// delta_ws: caller scratch of b*heads*Lq floats
// Backward for query sets of more than 64 rows: the chunked dQ kernel + the looping dK / dV kernel (two launches).
template <int NCH>
static int launch_bwd_loop(const void* q, const void* k, const void* v, const void* o, const void* d_o, const float* lse,
                           void* dq, void* dk, void* dv, float* delta, int b, int Lq, int Lk, int heads, int ldq, int ldk,
                           int ldv, float sm_scale, float p, uint64_t seed, uint64_t offset, cudaStream_t st) {
  const int dh = NCH * 64, HD = heads * dh;
  const int nw = XA_MAXW;
  const float sl2 = sm_scale * kLog2e;
  CUtensorMap tq, tdo, tk, tv, tk2, tv2;
  int rc;
  if ((rc = xa_make_tmap(&tq, q, b, Lq, HD, ldq, 16 * nw))) return rc;
  if ((rc = xa_make_tmap(&tdo, d_o, b, Lq, HD, HD, 16 * nw))) return rc;
  if ((rc = xa_make_tmap(&tk, k, b, Lk, HD, ldk, XA_KT))) return rc;
  if ((rc = xa_make_tmap(&tv, v, b, Lk, HD, ldv, XA_KT))) return rc;
  if ((rc = xa_make_tmap(&tk2, k, b, Lk, HD, ldk, XA_KTB))) return rc;
  if ((rc = xa_make_tmap(&tv2, v, b, Lk, HD, ldv, XA_KTB))) return rc;
  XTAG_CUDA(sync_spin_timeout());
  {
    const size_t q_pad = ((size_t)NCH * 16 * nw * 128 + 1023) & ~(size_t)1023;
    const size_t smem = 1024 + 2 * q_pad + 2 * (2 * NCH * XA_KT * 128) + 64 + 16 * nw * 4;
    XTAG_CUDA(cudaFuncSetAttribute(xattn_bwd_dq_kernel<NCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const dim3 grid((unsigned)(b * heads), (unsigned)((Lq + 16 * nw - 1) / (16 * nw)));
    xattn_bwd_dq_kernel<NCH><<<grid, 32 * nw, smem, st>>>(tq, tk, tv, tdo, (const __nv_bfloat16*)o,
                                                          (const __nv_bfloat16*)d_o, lse, (__nv_bfloat16*)dq, delta, Lq, Lk,
                                                          heads, sl2, sm_scale, p, seed, offset);
    XTAG_CHECK_LAUNCH();
  }
  {
    const int tiles = (Lk + XA_KTB - 1) / XA_KTB;
    const size_t smem = 1024 + 2 * (size_t)NCH * 64 * 128 + 2 * (size_t)NCH * XA_KTB * 128 + 64 + 2 * 64 * 4;
    XTAG_CUDA(cudaFuncSetAttribute(xattn_bwd_dkv_loop_kernel<NCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    xattn_bwd_dkv_loop_kernel<NCH><<<b * heads * tiles, 128, smem, st>>>(tq, tk2, tv2, tdo, lse, delta, (__nv_bfloat16*)dk,
                                                                         (__nv_bfloat16*)dv, Lq, Lk, heads, tiles, sl2,
                                                                         sm_scale, p, seed, offset);
    XTAG_CHECK_LAUNCH();
  }
  return XTAG_OK;
}

int xattn_mma_bwd(const void* q, const void* k, const void* v, const void* o, const void* d_o, const float* lse,
                  void* dq, void* dk, void* dv, float* delta_ws, int b, int Lq, int Lk, int heads, int dh, int ldq, int ldk,
                  int ldv, int lddk, int lddv, float sm_scale, float p_drop, uint64_t seed, uint64_t offset,
                  cudaStream_t st) {
  // the single-pass kernel keeps dQ [16 x dh] and one [32 x dh/NW] product tile per warp in registers: shapes whose
  // accumulators would spill (few warps with a wide head) stay on the two-kernel path
  if (Lq > 16 * XA_MAXW) {
    XTAG_REQUIRE(lddk == heads * dh && lddv == heads * dh, XTAG_ERR_UNSUPPORTED,
                 "xattn_bwd: strided dK / dV outputs are not supported for query sets of more than %d rows", 16 * XA_MAXW);
    switch (dh / 64) {
      case 1: return launch_bwd_loop<1>(q, k, v, o, d_o, lse, dq, dk, dv, delta_ws, b, Lq, Lk, heads, ldq, ldk, ldv, sm_scale, p_drop, seed, offset, st);
      case 2: return launch_bwd_loop<2>(q, k, v, o, d_o, lse, dq, dk, dv, delta_ws, b, Lq, Lk, heads, ldq, ldk, ldv, sm_scale, p_drop, seed, offset, st);
      case 3: return launch_bwd_loop<3>(q, k, v, o, d_o, lse, dq, dk, dv, delta_ws, b, Lq, Lk, heads, ldq, ldk, ldv, sm_scale, p_drop, seed, offset, st);
      default: return launch_bwd_loop<4>(q, k, v, o, d_o, lse, dq, dk, dv, delta_ws, b, Lq, Lk, heads, ldq, ldk, ldv, sm_scale, p_drop, seed, offset, st);
    }
  }
  const int nw = (Lq + 15) / 16, nch = dh / 64;
  const int acc_regs = 16 * ((nch * 4 + nw - 1) / nw) + 32 * nch;
  if ((tc_tune() & kTuneXattnFusedBwd) && acc_regs <= 224) {
    switch (nch) {
      case 1: return launch_bwd_fused_nch<1>(nw, q, k, v, o, d_o, lse, dq, dk, dv, b, Lq, Lk, heads, ldq, ldk, ldv, lddk, lddv, sm_scale, p_drop, seed, offset, st);
      case 2: return launch_bwd_fused_nch<2>(nw, q, k, v, o, d_o, lse, dq, dk, dv, b, Lq, Lk, heads, ldq, ldk, ldv, lddk, lddv, sm_scale, p_drop, seed, offset, st);
      case 3: return launch_bwd_fused_nch<3>(nw, q, k, v, o, d_o, lse, dq, dk, dv, b, Lq, Lk, heads, ldq, ldk, ldv, lddk, lddv, sm_scale, p_drop, seed, offset, st);
      default: return launch_bwd_fused_nch<4>(nw, q, k, v, o, d_o, lse, dq, dk, dv, b, Lq, Lk, heads, ldq, ldk, ldv, lddk, lddv, sm_scale, p_drop, seed, offset, st);
    }
  }
  XTAG_REQUIRE(lddk == heads * dh && lddv == heads * dh, XTAG_ERR_UNSUPPORTED,
               "xattn_bwd: strided dK / dV outputs need the single-pass kernel (tune bit 11, accumulators that fit)");
  switch (dh / 64) {
    case 1: return launch_bwd<1>(q, k, v, o, d_o, lse, dq, dk, dv, delta_ws, b, Lq, Lk, heads, ldq, ldk, ldv, sm_scale, p_drop, seed, offset, st);
    case 2: return launch_bwd<2>(q, k, v, o, d_o, lse, dq, dk, dv, delta_ws, b, Lq, Lk, heads, ldq, ldk, ldv, sm_scale, p_drop, seed, offset, st);
    case 3: return launch_bwd<3>(q, k, v, o, d_o, lse, dq, dk, dv, delta_ws, b, Lq, Lk, heads, ldq, ldk, ldv, sm_scale, p_drop, seed, offset, st);
    default: return launch_bwd<4>(q, k, v, o, d_o, lse, dq, dk, dv, delta_ws, b, Lq, Lk, heads, ldq, ldk, ldv, sm_scale, p_drop, seed, offset, st);
  }
}

}  // namespace xtag
