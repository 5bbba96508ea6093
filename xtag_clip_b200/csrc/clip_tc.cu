// Contrastive head, bf16 production path: tcgen05 / TMEM / TMA (sm_100a).
//
// One persistent, warp-specialised kernel template computes  acc[M,N] = A[M,K] * B[N,K]^T  (both operands
// bf16, K-major, fp32 accumulation in TMEM) over 128x256 tiles and hands every finished accumulator to one
// of three fused epilogues:
//
//   EPI_LSE    K1 forward   S = scale*acc never leaves the SM: per tile the epilogue emits log2-domain
//                           (max,sum-exp) partials per row and per column and picks the label logit
//                           (reference src/open_clip/loss.py:116-124 + F.cross_entropy at :134-137)
//   EPI_DS     K2 backward  S recomputed, dS = g*(w_row*softmax_row + w_col*softmax_col - w_diag*1[label])
//                           written once as bf16 by per-warp TMA tile stores + d(logit_scale) partials
//   EPI_STORE  plain GEMM   C = alpha*acc (fp32 or bf16): dA = scale*dS*Bm and dB = scale*dS^T*A, with dS and the
//                           features read in place (MN-major UMMA operands)
//
// Pipeline per CTA (384 threads, 1 CTA / SM, grid = #SMs, static tile schedule with grouped rasterisation):
//   warp 0      TMA producer  : cp.async.bulk.tensor 2-D tiles (128B swizzle) into a 4-stage smem ring
//   warp 1      MMA issuer    : one elected lane issues tcgen05.mma 128x256x16, commits to mbarriers
//   warp 2      TMEM allocator: 512 columns = two 128x256 fp32 accumulators (epilogue of tile i overlaps
//                               the MMAs of tile i+1)
//   warps 4-11  epilogue      : tcgen05.ld 32x32b (lane = row), 2 warps per TMEM lane quadrant, each owning
//                               one 128-column half of the accumulator
#include <cuda.h>
#include <string.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace xtag {

using namespace ptx;

constexpr int BM = 128, BN = 256, BK = 64;
constexpr int kABytes = BM * BK * 2;          // 16 KB
constexpr int kEpiWarps = 8;
constexpr int kThreads = 128 + 32 * kEpiWarps;   // 384
constexpr int kTmemCols = 512;
constexpr int kGroupM = 16;
constexpr int kEpiScratchFloats = 2 /*acc*/ * 2 /*half*/ * 4 /*quadrant*/ * 128;   // 8 KB
constexpr int kOutStageBytes = kEpiWarps * 2048;   // per epilogue warp: one 32-row x 64-byte tile for the dS TMA store
// Shared-memory geometry per CTA.  CG = 1: one CTA per 128 x 256 tile (tcgen05.mma.cta_group::1).  CG = 2: a CTA PAIR
// (two SMs of one TPC) per 256 x 256 tile (tcgen05.mma.cta_group::2, M = 256): every CTA stages its own 128 rows of A
// and only ITS HALF of the B tile, so a k-block costs 32 KB instead of 48 KB of shared memory (6 ring stages instead of
// 4: 50 % more look-ahead against L2 / DRAM latency) and each SM reads a third less operand data per MMA.
template <int CG>
struct Geo {
  static constexpr int kBRows = BN / CG;
  static constexpr int kBBytes = kBRows * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (CG == 2) ? 6 : 4;
  static constexpr int kSmemBytes = 1024 /*align slack*/ + kStages * kStageBytes + 256 /*barriers*/ +
                                    kEpiScratchFloats * 4 + kOutStageBytes;
};
constexpr int kMaxStages = 6;

enum { EPI_LSE = 0, EPI_DS = 1, EPI_STORE = 2, EPI_BIAS = 3, EPI_SIG = 4 };

struct EpiParams {
  // EPI_LSE
  const float* scale_p;   // device scalar: logit scale (natural units)
  int label_offset;
  float* row_part;        // [2*num_n][M]   log2-domain LSE partial per (n tile, column half)
  float* col_part;        // [num_m][col_ld] log2-domain LSE partial per m tile (col_ld >= N: several column blocks
  int col_ld;             //                 of one forward may share a buffer, each at its own column offset)
  float* diag;            // [M]            natural units
  // EPI_DS
  const float* row_lse;   // [M] natural
  const float* col_lse;   // [N] natural
  float w_row, w_col, w_diag;
  const float* grad_out;  // device scalar
  __nv_bfloat16* dS;      // [M][ldds]
  int ldds;
  float* dscale_part;     // [grid * kEpiWarps]
  // EPI_STORE
  void* C;
  int ldc;
  int c_is_bf16;
  float alpha;            // C = alpha * (alpha_p ? *alpha_p : 1) * acc
  const float* alpha_p;
  // EPI_BIAS: C (bf16, row stride ldc, written by TMA tile stores) = alpha * acc + bias[column]
  const float* bias;      // [N] fp32 or nullptr
  // EPI_STORE split-K: the K loop is cut into split_k slices, scheduled as split_k x tiles work items; slice ks writes
  // its raw fp32 partial to C + ks * split_stride (a workspace), a second kernel sums the slices into the real output
  int group_m;            // m tiles per group of the tile schedule (0: default kGroupM / CG)
  int split_k;            // 0 / 1: no split
  long split_stride;      // elements between consecutive partial slabs
  // EPI_SIG (sigmoid loss): z = scale * acc + *bias_p; label +1 on column gi + label_offset, -1 elsewhere;
  // loss = w_sig * sum softplus(-label * z); dS (for a unit upstream gradient) = -label * w_sig * sigmoid(-label * z),
  // staged as bf16 like EPI_DS when sig_store != 0.  Partials per epilogue warp: loss_part, dscale_part (sum dS * acc),
  // dbias_part (sum dS).
  const float* bias_p;
  float w_sig;
  int sig_store;
  float* loss_part;
  float* dbias_part;
  // streamed forward (EPI_LSE over a gather buffer that fills up block by block): column block k of blk_tiles n tiles
  // is is one slab of the tile schedule, slabs are visited in arrival order; ready_flags[blk] == *epoch_p once block blk has landed
  const int* ready_flags; // device [nblk] or nullptr
  const int* epoch_p;     // device scalar
  int blk_tiles;          // n tiles per column block (0: not streamed)
  int blk_order[16];      // slab s of the schedule works on column block blk_order[s]
  int blk_wait[16];       // 1: slab s must wait for its ready flag
  // all epilogues: runtime tuning bits (xtag_set_tune, documented in include/xtag_b200.h)
  int tune;
};
constexpr int kTunePrefetchMask = 0xff, kTuneStoreEvictFirst = 0x100, kTuneStreamAEvictFirst = 0x200,
              kTuneDsTwoExp = 0x400,
              kTuneCluster2 = 0x4000,       // clusters of 2 CTAs along M with TMA multicast of the shared B tile
              kTuneCluster4 = 0x8000,       // clusters of 4
              kTuneDbgNoDsStore = 0x1000,   // diagnostics only (wrong results): dS tile neither staged nor stored
              kTuneDbgNoDsTma = 0x2000,     // diagnostics only (wrong results): dS tile staged in smem, TMA store skipped
              kTuneNoSplitK = 0x8000000,    // plain GEMMs: never split the K loop
              kTuneNoPair = 0x1000000,      // do NOT use the CTA-pair (cta_group::2) kernels
              kTuneBEvictLast = 0x2000000,  // B operand tiles: L2 evict_last
              kTuneAEvictLast = 0x4000000;  // A operand tiles: L2 evict_last

// Static tile schedule.  Tiles are ordered  n-slab  >  group of kGroupM m tiles  >  n tile inside the slab  >  m tile
// inside the group, so that at any time the CTAs work on a [kGroupM m tiles] x [~#SM / kGroupM n tiles] patch (operand
// tiles shared through L2 by the CTAs of a wave), a group's A tiles are re-used across one slab, and a slab's B tiles are
// re-used by every group before the schedule moves on: with `slab` n tiles per slab the B slab (slab * BN rows) stays
// L2-resident while the groups stream past it even when the kernel also writes a large output (the dS producer writes
// 2*BM*BN bytes per tile; without slabs every group pass re-fetched the whole B matrix from HBM and each k-block load
// paid DRAM latency, which the 4-stage ring cannot hide: measured 77 % vs 95 % tensor-pipe activity).
__host__ __device__ __forceinline__ void tile_coords(int tile, int num_m, int num_n, int slab, const int* slab_order,
                                                     int& m_blk, int& n_blk, int* slab_idx = nullptr,
                                                     int group_m = kGroupM) {
  const int per_slab = num_m * slab;
  const int s = tile / per_slab;
  if (slab_idx) *slab_idx = s;
  const int n0 = (slab_order ? slab_order[s] : s) * slab;
  const int ns = (slab < num_n - n0) ? slab : num_n - n0;
  const int r = tile - s * per_slab;
  const int per_group = group_m * ns;
  const int group = r / per_group;
  const int first_m = group * group_m;
  const int gsize = (num_m - first_m < group_m) ? num_m - first_m : group_m;
  const int in = r - group * per_group;
  m_blk = first_m + in % gsize;
  n_blk = n0 + in / gsize;
}

// Transpose-reduce over the 32 lanes of a warp: on entry x[j] is lane-local value for column j; on exit
// x[0] on lane L holds op-reduction over all lanes of column L.  31 shuffles.
template <bool kMax>
__device__ __forceinline__ void warp_transpose_reduce(float (&x)[32], int lane) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    const bool upper = (lane & o) != 0;
#pragma unroll
    for (int t = 0; t < o; ++t) {
      const float send = upper ? x[t] : x[t + o];
      const float keep = upper ? x[t + o] : x[t];
      const float recv = __shfl_xor_sync(0xffffffffu, send, o);
      x[t] = kMax ? fmaxf(keep, recv) : (keep + recv);
    }
  }
}

// r[idx] for a per-thread idx in [0,32) as a 5-level select tree (31 selects, registers only: a dynamic
// subscript would push the whole array to local memory)
__device__ __forceinline__ float select32(const uint32_t (&r)[32], int idx) {
  float t16[16], t8[8], t4[4], t2[2];
#pragma unroll
  for (int k = 0; k < 16; ++k) t16[k] = __uint_as_float((idx & 16) ? r[k + 16] : r[k]);
#pragma unroll
  for (int k = 0; k < 8; ++k) t8[k] = (idx & 8) ? t16[k + 8] : t16[k];
#pragma unroll
  for (int k = 0; k < 4; ++k) t4[k] = (idx & 4) ? t8[k + 4] : t8[k];
#pragma unroll
  for (int k = 0; k < 2; ++k) t2[k] = (idx & 2) ? t4[k + 2] : t4[k];
  return (idx & 1) ? t2[1] : t2[0];
}

__device__ __forceinline__ uint32_t pack2_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ---------------------------------------------------------------------------------------------------------
// Epilogue bodies.  One call handles this warp's 32 rows x 128 columns of a finished accumulator (4 chunks of
// 32 columns read with tcgen05.ld 32x32b: thread = row).  FULL = no ragged edge inside the block, so every
// bounds predicate folds away at compile time.
// ---------------------------------------------------------------------------------------------------------

// K1: log2-domain (max, sum-exp) partials per row and per column + label logit.
//   one exponential per element, referenced to the row's chunk maximum cm_i:
//     row   : sum_j 2^(v_ij - cm_i), folded into the running (max, sum) with two scalar exps per chunk
//     column: 2^(v_ij - W) = e_ij * 2^(cm_i - W), W = max of cm over the warp's 32 rows -> the same exponentials
//             serve the column sums after one multiply and a 31-shuffle transpose-reduce.
//   Exact unless an element lies more than 2^-kRange below W (it would flush to zero while possibly dominating
//   its own column): such chunks -- a 32x32 block spanning > 83 nats -- take the exact two-exp path.
// "accumulator drained": in CTA-pair mode both CTAs' epilogue warps report to the LEADER's barrier (its MMA thread
// writes both halves of the next accumulator)
template <int CG>
__device__ __forceinline__ void arrive_tempty(uint32_t tempty) {
  if constexpr (CG == 2) mbar_arrive_cluster(tempty, 0);
  else                   mbar_arrive(tempty);
}

template <bool FULL, int CG>
__device__ __forceinline__ void lse_tile(const EpiParams& ep, uint32_t taddr, float* scratch, int gi, int n_base,
                                         int m_blk, int n_blk, int h, int q, int lane, int M, int N,
                                         uint32_t tempty, int bar_id) {
  constexpr float kRange = 120.f;
  const float sl2 = ep.scale_p[0] * kLog2e;
  const bool row_ok = FULL || gi < M;
  const int lab = (ep.label_offset < 0) ? -1 : gi + ep.label_offset;      // -1: no labels in this column block
  float m_run = -INFINITY, l_run = 0.f;
#pragma unroll 1
  for (int ch = 0; ch < 4; ++ch) {
    uint32_t r[32];
    tmem_ld_32x32(taddr + ch * 32, r);
    tmem_ld_wait();
    const int c0 = n_base + ch * 32;
    // raw accumulators a_j = <A_i, B_j>; v_j = a_j * sl2.  max / min on the raw values in 4 interleaved chains
    // (two epilogue warps per scheduler hide little latency); the scale is folded into the exponent FFMA.
    float amx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
    float amn[4] = {INFINITY, INFINITY, INFINITY, INFINITY};
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const float a = __uint_as_float(r[j]);
      if (FULL || (c0 + j) < N) {
        amx[j & 3] = fmaxf(amx[j & 3], a);
        amn[j & 3] = fminf(amn[j & 3], a);
      }
    }
    const float amax = fmaxf(fmaxf(amx[0], amx[1]), fmaxf(amx[2], amx[3]));
    const float amin = fminf(fminf(amn[0], amn[1]), fminf(amn[2], amn[3]));
    const bool any_col = FULL || amax > -INFINITY;                // false only for a chunk entirely past N
    const float cm = any_col ? ((sl2 >= 0.f) ? amax * sl2 : amin * sl2) : -INFINITY;
    const float vmin = any_col ? ((sl2 >= 0.f) ? amin * sl2 : amax * sl2) : INFINITY;
    if (row_ok && lab >= c0 && lab < c0 + 32) {
      const float dv = select32(r, lab - c0);
      // same rounding sequence as the row LSE ((a*sl2 + log2 l) * ln2): lse - diag cancels exactly for rows
      // dominated by their label
      ep.diag[gi] = (dv * sl2) * kLn2;
    }
    float ex[32];
    float rsp[4] = {0.f, 0.f, 0.f, 0.f};
    const float ncm = any_col ? -cm : 0.f;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const float t = fast_exp2(fmaf(__uint_as_float(r[j]), sl2, ncm));
      ex[j] = (FULL || (c0 + j) < N) ? t : 0.f;
      rsp[j & 3] += ex[j];
    }
    const float rsum = (rsp[0] + rsp[1]) + (rsp[2] + rsp[3]);
    if (any_col) {
      const float m_new = fmaxf(m_run, cm);
      l_run = l_run * fast_exp2(m_run - m_new) + rsum * fast_exp2(cm - m_new);
      m_run = m_new;
    }
    const float W = warp_max(row_ok ? cm : -INFINITY);
    const float lo = -warp_max(row_ok ? -vmin : -INFINITY);
    float* cslot = scratch + q * 128 + ch * 32;
    if (W - lo <= kRange) {
      // fast path (warp-uniform branch)
      const float f = (row_ok && any_col) ? fast_exp2(cm - W) : 0.f;
#pragma unroll
      for (int j = 0; j < 32; ++j) ex[j] *= f;
      warp_transpose_reduce<false>(ex, lane);             // ex[0] = sum over rows of column `lane`
      cslot[lane] = (ex[0] > 0.f) ? W + fast_log2(ex[0]) : -INFINITY;
    } else {
      // exact path: per-column maximum, then a second exponential
      // (re-reads the chunk from TMEM so the fast path need not keep the raw accumulators alive)
      float x[32], v[32];
      uint32_t rr[32];
      tmem_ld_32x32(taddr + ch * 32, rr);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        v[j] = (FULL || (c0 + j) < N) ? __uint_as_float(rr[j]) * sl2 : -INFINITY;
        x[j] = row_ok ? v[j] : -INFINITY;
      }
      warp_transpose_reduce<true>(x, lane);               // x[0] = max of column `lane`
      cslot[lane] = x[0];
      __syncwarp();
#pragma unroll
      for (int j4 = 0; j4 < 8; ++j4) {
        const float4 c4 = *reinterpret_cast<const float4*>(cslot + j4 * 4);
        const float cmx[4] = {c4.x, c4.y, c4.z, c4.w};
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int j = j4 * 4 + u;
          // column with no valid entry: max = -inf -> contributes nothing
          ex[j] = (row_ok && cmx[u] > -INFINITY) ? fast_exp2(v[j] - cmx[u]) : 0.f;
        }
      }
      __syncwarp();
      const float my_cmax = x[0];
      warp_transpose_reduce<false>(ex, lane);
      cslot[lane] = (ex[0] > 0.f) ? my_cmax + fast_log2(ex[0]) : -INFINITY;
    }
  }
  // row partial for this (n tile, half)
  if (row_ok) ep.row_part[(size_t)(n_blk * 2 + h) * M + gi] = (l_run > 0.f) ? m_run + fast_log2(l_run) : -INFINITY;
  // TMEM reads of this warp are complete: release the accumulator before the cross-warp combine
  tc_fence_before();
  __syncwarp();
  if (lane == 0) arrive_tempty<CG>(tempty);
  named_bar_sync(bar_id, 128);
  {
    const int c = q * 32 + lane;
    const float p0 = scratch[0 * 128 + c], p1 = scratch[1 * 128 + c];
    const float p2 = scratch[2 * 128 + c], p3 = scratch[3 * 128 + c];
    const float mx = fmaxf(fmaxf(p0, p1), fmaxf(p2, p3));
    float out = -INFINITY;
    if (mx > -INFINITY)
      out = mx + fast_log2(fast_exp2(p0 - mx) + fast_exp2(p1 - mx) + fast_exp2(p2 - mx) + fast_exp2(p3 - mx));
    const int gj = n_base + c;
    if (FULL || gj < N) ep.col_part[(size_t)m_blk * ep.col_ld + gj] = out;
  }
}

// K2a: dS = g * ( w_row * 2^(a*sl2 - rl2_i) + w_col * 2^(a*sl2 - cl2_j) ) - g*w_diag*[label]   (log2 domain).
//   FAST  one exponential per element: with e_ij = 2^(a*sl2 - rl2_i),
//           dS_ij = e_ij * ( g*w_row + (g*w_col*2^(rl2_i - nu)) * 2^(nu - cl2_j) )
//         the bracket is one FFMA of a per-row and a per-column factor (nu = a reference common to the 128 columns
//         of this half).  Exact up to rounding while |rl2_i - cl2_j| <= kDsFastRange for the whole 32 x 128 block
//         (no factor over/underflows); the caller checks that per warp and otherwise takes
//   !FAST the (non-negative) weights folded into two exponent offsets: one FFMA + one MUFU per term.
// scratch[0..127] = per-column offsets cl2' (two-exp path), scratch[128..255] = per-column factors 2^(nu - cl2_j).
// Returns this thread's partial of sum dS_ij * a_ij (= d logit_scale).
constexpr float kDsFastRange = 60.f;
template <bool FULL, bool FAST>
__device__ __forceinline__ float ds_tile(const EpiParams& ep, const CUtensorMap* tmC, uint32_t taddr,
                                         const float* scratch, uint32_t ostage, int gi, int row_blk, int n_base,
                                         float nrl2, float gwr, float gwc_r, uint64_t store_policy, int lane, int M,
                                         int N) {
  const float sl2 = ep.scale_p[0] * kLog2e;
  const float g = ep.grad_out[0];
  const float gwd = g * ep.w_diag;
  const bool row_ok = FULL || gi < M;
  const int lab = gi + ep.label_offset;
  float dsp[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
  for (int ch = 0; ch < 4; ++ch) {
    uint32_t r[32];
    tmem_ld_32x32(taddr + ch * 32, r);
    tmem_ld_wait();
    const int c0 = n_base + ch * 32;
    float d[32];
#pragma unroll
    for (int j4 = 0; j4 < 8; ++j4) {
      const float4 c4 = *reinterpret_cast<const float4*>(scratch + (FAST ? 128 : 0) + ch * 32 + j4 * 4);
      const float cc[4] = {c4.x, c4.y, c4.z, c4.w};
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int j = j4 * 4 + u;
        const float a = __uint_as_float(r[j]);
        if (FAST) {
          // out-of-range rows carry nrl2 = -inf (e = 0, finite factors); out-of-range columns are never stored
          // and multiply a zero accumulator in the d(logit_scale) sum
          d[j] = fast_exp2(fmaf(a, sl2, nrl2)) * fmaf(gwc_r, cc[u], gwr);
        } else {
          // out-of-range rows / columns carry offset -inf: their terms are exactly 0 (and are never stored)
          d[j] = g * (fast_exp2(fmaf(a, sl2, nrl2)) + fast_exp2(fmaf(a, sl2, -cc[u])));
        }
      }
    }
    if (row_ok && lab >= c0 && lab < c0 + 32) {
#pragma unroll
      for (int j = 0; j < 32; ++j) d[j] = (lab - c0 == j) ? d[j] - gwd : d[j];
    }
#pragma unroll
    for (int j = 0; j < 32; ++j) dsp[j & 3] = fmaf(d[j], __uint_as_float(r[j]), dsp[j & 3]);
    // ---- store: 32 rows x 64 B of bf16 through shared memory and one TMA tile store per warp and chunk.
    // Direct STG would write 16 B per 128-B line per instruction (32 L2 requests per warp instruction, measured as
    // the kernel's dominant stall); the TMA store issues full 64-B row segments and clips ragged edges itself.
    if (!(ep.tune & kTuneDbgNoDsStore)) {
      if (lane == 0) tma_store_wait_read();                 // previous tile store has drained this buffer
      __syncwarp();
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        uint32_t p0 = pack2_bf16(d[u * 8 + 0], d[u * 8 + 1]), p1 = pack2_bf16(d[u * 8 + 2], d[u * 8 + 3]);
        uint32_t p2 = pack2_bf16(d[u * 8 + 4], d[u * 8 + 5]), p3 = pack2_bf16(d[u * 8 + 6], d[u * 8 + 7]);
        // CU_TENSOR_MAP_SWIZZLE_64B: 16-byte unit index (bits 4-5) ^= address bits 7-8 (= (row >> 1) & 3)
        const uint32_t off = (uint32_t)(lane * 64 + ((u ^ ((lane >> 1) & 3)) << 4));
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(ostage + off), "r"(p0), "r"(p1), "r"(p2), "r"(p3)
                     : "memory");
      }
      fence_proxy_async();                                  // generic-proxy writes -> visible to the TMA engine
      __syncwarp();
      if (lane == 0 && !(ep.tune & kTuneDbgNoDsTma)) {
        if (store_policy) tma_store_2d_hint(tmC, ostage, c0, row_blk, store_policy);
        else              tma_store_2d(tmC, ostage, c0, row_blk);
        tma_store_commit();
      }
    }
  }
  return (dsp[0] + dsp[1]) + (dsp[2] + dsp[3]);
}

// plain GEMM epilogue: C = alpha * acc (fp32 or bf16)
__device__ __forceinline__ void store_tile(const EpiParams& ep, uint32_t taddr, int gi, int n_base, int M, int N,
                                           long c_off) {
  const float alpha = ep.alpha * (ep.alpha_p ? ep.alpha_p[0] : 1.f);
  const bool row_ok = gi < M;
#pragma unroll 1
  for (int ch = 0; ch < 4; ++ch) {
    uint32_t r[32];
    tmem_ld_32x32(taddr + ch * 32, r);
    tmem_ld_wait();
    const int c0 = n_base + ch * 32;
    if (!row_ok) continue;
    if (ep.c_is_bf16) {
      __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(ep.C) + (size_t)gi * ep.ldc + c0;
      if (c0 + 32 <= N && (ep.ldc % 8) == 0) {
#pragma unroll
        for (int j8 = 0; j8 < 4; ++j8) {
          uint4 pk;
          __nv_bfloat162* hp = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
          for (int u = 0; u < 4; ++u)
            hp[u] = __floats2bfloat162_rn(__uint_as_float(r[j8 * 8 + 2 * u]) * alpha,
                                          __uint_as_float(r[j8 * 8 + 2 * u + 1]) * alpha);
          *reinterpret_cast<uint4*>(dst + j8 * 8) = pk;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (c0 + j < N) dst[j] = __float2bfloat16_rn(__uint_as_float(r[j]) * alpha);
      }
    } else {
      float* dst = reinterpret_cast<float*>(ep.C) + c_off + (size_t)gi * ep.ldc + c0;
      if (c0 + 32 <= N && (ep.ldc % 4) == 0) {
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4)
          *reinterpret_cast<float4*>(dst + j4 * 4) =
              make_float4(__uint_as_float(r[j4 * 4]) * alpha, __uint_as_float(r[j4 * 4 + 1]) * alpha,
                          __uint_as_float(r[j4 * 4 + 2]) * alpha, __uint_as_float(r[j4 * 4 + 3]) * alpha);
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (c0 + j < N) dst[j] = __uint_as_float(r[j]) * alpha;
      }
    }
  }
}

// dense-layer epilogue: C = bf16(alpha * acc + bias[col]) through the same per-warp staging + TMA tile stores as the dS
// producer (a [b*N, 3072] K|V projection writes > 1 GB: direct 16-byte stores per thread-row would issue 32 L2 requests
// per warp instruction); rows / columns past the edge are clipped by the TMA unit
__device__ __forceinline__ void bias_tile(const EpiParams& ep, const CUtensorMap* tmC, uint32_t taddr, uint32_t ostage,
                                          int row_blk, int n_base, int lane, int N) {
  const float alpha = ep.alpha;
#pragma unroll 1
  for (int ch = 0; ch < 4; ++ch) {
    uint32_t r[32];
    tmem_ld_32x32(taddr + ch * 32, r);
    tmem_ld_wait();
    const int c0 = n_base + ch * 32;
    if (c0 >= N) continue;                                  // warp-uniform
    float d[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const float bj = (ep.bias && c0 + j < N) ? __ldg(ep.bias + c0 + j) : 0.f;
      d[j] = fmaf(__uint_as_float(r[j]), alpha, bj);
    }
    if (lane == 0) tma_store_wait_read();
    __syncwarp();
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      uint32_t p0 = pack2_bf16(d[u * 8 + 0], d[u * 8 + 1]), p1 = pack2_bf16(d[u * 8 + 2], d[u * 8 + 3]);
      uint32_t p2 = pack2_bf16(d[u * 8 + 4], d[u * 8 + 5]), p3 = pack2_bf16(d[u * 8 + 6], d[u * 8 + 7]);
      const uint32_t off = (uint32_t)(lane * 64 + ((u ^ ((lane >> 1) & 3)) << 4));
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(ostage + off), "r"(p0), "r"(p1), "r"(p2), "r"(p3)
                   : "memory");
    }
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) {
      tma_store_2d(tmC, ostage, c0, row_blk);
      tma_store_commit();
    }
  }
}

// Sigmoid (SigLIP) loss epilogue, reference src/open_clip/loss.py:345-359: ONE pass yields the loss AND the complete
// logit gradient (no row / column normaliser exists, so nothing has to be known before the pass): the backward is
// just the two gradient GEMMs on the staged dS -- 6 B^2 D executed FLOPs per step, the algorithmic count.
// acc3: this thread's partial sums (loss, sum dS * acc, sum dS).
template <bool FULL>
__device__ __forceinline__ void sig_tile(const EpiParams& ep, const CUtensorMap* tmC, uint32_t taddr, uint32_t ostage,
                                         int gi, int row_blk, int n_base, int lane, int M, int N, float (&acc3)[3]) {
  const float s = ep.scale_p[0], bias = ep.bias_p[0], w = ep.w_sig;
  const bool row_ok = FULL || gi < M;
  const int lab = gi + ep.label_offset;
  float lsum[2] = {0.f, 0.f}, dsa[2] = {0.f, 0.f}, dsb[2] = {0.f, 0.f};
#pragma unroll 1
  for (int ch = 0; ch < 4; ++ch) {
    uint32_t r[32];
    tmem_ld_32x32(taddr + ch * 32, r);
    tmem_ld_wait();
    const int c0 = n_base + ch * 32;
    float d[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const float a = __uint_as_float(r[j]);
      const float z = fmaf(a, s, bias);
      const bool pos = (lab == c0 + j);
      const float t = pos ? -z : z;                          // -label * z
      const float e = fast_exp2(-fabsf(t) * kLog2e);         // exp(-|t|) in (0, 1]
      const float inv = __frcp_rn(1.f + e);
      const float sp = fmaxf(t, 0.f) + fast_log2(1.f + e) * kLn2;      // softplus(t) = -logsigmoid(label * z)
      const float sg = (t >= 0.f) ? inv : e * inv;                      // sigmoid(t)
      const bool ok = row_ok && (FULL || (c0 + j) < N);
      const float dv = ok ? (pos ? -w * sg : w * sg) : 0.f;             // d loss / d z
      d[j] = dv;
      lsum[j & 1] += ok ? sp : 0.f;
      dsa[j & 1] = fmaf(dv, a, dsa[j & 1]);
      dsb[j & 1] += dv;
    }
    if (ep.sig_store) {
      if (lane == 0) tma_store_wait_read();
      __syncwarp();
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        uint32_t p0 = pack2_bf16(d[u * 8 + 0], d[u * 8 + 1]), p1 = pack2_bf16(d[u * 8 + 2], d[u * 8 + 3]);
        uint32_t p2 = pack2_bf16(d[u * 8 + 4], d[u * 8 + 5]), p3 = pack2_bf16(d[u * 8 + 6], d[u * 8 + 7]);
        const uint32_t off = (uint32_t)(lane * 64 + ((u ^ ((lane >> 1) & 3)) << 4));
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(ostage + off), "r"(p0), "r"(p1), "r"(p2), "r"(p3)
                     : "memory");
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        tma_store_2d(tmC, ostage, c0, row_blk);
        tma_store_commit();
      }
    }
  }
  acc3[0] += (lsum[0] + lsum[1]) * w;
  acc3[1] += dsa[0] + dsa[1];
  acc3[2] += dsb[0] + dsb[1];
}

// A_MN / B_MN: the operand is stored "MN-major": global tensor [K rows][M or N contiguous] (e.g. dS read as the
// A operand of dB = dS^T A, or row-major features read as the B operand [N=D][K] of dA = dS Bm).  TMA then loads
// 64-element (128 B) wide boxes of BK rows; UMMA reads them through an MN-major 128B-swizzle descriptor.
// CL = thread-block cluster size along M (1, 2 or 4): the CL CTAs of a cluster work on CL consecutive m tiles of the
// SAME n tile, so the B operand tile is shared: each CTA fetches 1/CL of it and TMA-multicasts that slice into the
// shared memory of all CL CTAs.  Per CTA the TMA unit then moves BM + BN/CL rows per k-block instead of BM + BN and
// the L2 -> SM operand traffic drops by the same factor (measured: the TMA/L2 request rate, not the tensor pipe, is
// what the 128x256 tiles saturate first; an added L2 prefetch stream slows the kernels by 1.4x).
// CG = 2: CTA-pair kernel (see Geo): cluster = the pair, rank 0 = leader.  Both CTAs run a TMA producer (own A rows, own
// half of B, bytes credited to the leader's "full" barrier) and the epilogue of their 128 accumulator rows; only the
// leader's MMA thread issues tcgen05.mma.cta_group::2 and its commits are multicast onto both CTAs' barriers.
template <int EPI, bool A_MN, bool B_MN, int CL, int CG>
__global__ void __launch_bounds__(kThreads, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmC, int M, int N, int K, const __grid_constant__ EpiParams ep) {
  // 1024-byte alignment is required by the 128B-swizzle atoms; align inside the shared window with pointer
  // arithmetic on the __shared__ array itself so the compiler keeps emitting LDS/STS (a uintptr_t round trip turns
  // every access into a generic LD/ST: measured as "lg" stalls in the epilogue)
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* stage_base = smem;
  // layout: [operand ring | dS store staging (both 1024-byte aligned: swizzle patterns use absolute address bits)
  //          | mbarriers + TMEM pointer (256 B) | epilogue scratch]
  static_assert(CG == 1 || CL == 1, "the CTA pair is the cluster: no additional multicast clusters");
  using G = Geo<CG>;
  constexpr int kStages = G::kStages, kStageBytes = G::kStageBytes;
  constexpr bool kClustered = (CL > 1) || (CG == 2);
  uint8_t* out_stage = smem + kStages * kStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(out_stage + kOutStageBytes);
  uint64_t* full_bar = bars;                      // [kStages]
  uint64_t* empty_bar = bars + kStages;           // [kStages]
  uint64_t* tfull_bar = bars + 2 * kStages;       // [2]
  uint64_t* tempty_bar = bars + 2 * kStages + 2;  // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);
  float* epi_scratch = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  // scheduled tiles: BM*CG rows x BN columns (a CTA pair works on two vertically adjacent 128-row blocks)
  const int num_m = (M + BM * CG - 1) / (BM * CG), num_n = (N + BN - 1) / BN;
  const int num_k_all = (K + BK - 1) / BK;
  // split-K (plain GEMM only): work item = tile * S + slice; the S slices of a tile run side by side
  const int S = (EPI == EPI_STORE && ep.split_k > 1) ? ep.split_k : 1;
  const int kb_per = (num_k_all + S - 1) / S;
  const int num_tiles = num_m * num_n * S;
  // m tiles per schedule group: kGroupM / CG by default (same operand footprint per wave in both modes); 1 = the n
  // tiles of one m tile are consecutive (outputs with few n tiles: all CTAs that share an A row block run together)
  const int kGroup = ep.group_m > 0 ? ep.group_m : kGroupM / CG;
  const int tile_first = (CG == 2) ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int tile_step = (CG == 2) ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int slab_req = (ep.tune >> 16) & 0xff;                    // n tiles per slab of the tile schedule (0: one slab)
  const bool streamed = (EPI == EPI_LSE) && ep.blk_tiles > 0;
  const int slab = streamed ? ep.blk_tiles : ((slab_req > 0 && slab_req < num_n) ? slab_req : num_n);
  const int* slab_order = streamed ? ep.blk_order : nullptr;

  if (warp == 0 && elect_one()) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
  }
  if (warp == 1 && elect_one()) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(smem_u32(&full_bar[i]), 1);
      // multicast clusters: a slot is rewritten by every CTA of the cluster (B multicast), each of which commits;
      // CTA pair: one multicast commit of the leader arrives once on each CTA's barrier
      mbar_init(smem_u32(&empty_bar[i]), CL);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&tfull_bar[i]), 1);
      mbar_init(smem_u32(&tempty_bar[i]), kEpiWarps * CG);   // pair: both CTAs' epilogue warps report to the leader
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc<CG>(smem_u32(tmem_ptr), kTmemCols);
    tmem_relinquish<CG>();
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (kClustered) cluster_sync();        // peers' barriers are initialised before any multicast can arrive
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  [[maybe_unused]] const uint32_t cta_rank = kClustered ? cluster_ctarank() : 0u;
  [[maybe_unused]] constexpr uint16_t kClMask = (uint16_t)((1u << (CL * CG)) - 1u);

  if (warp == 0) {
    // ===================================== TMA producer =====================================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      const int pf_dist = (CG == 1 && S == 1) ? (ep.tune & kTunePrefetchMask) : 0;
      // bit 9: the A operand of the plain GEMMs is the staged dS, read exactly once: mark it evict_first so that the
      // stream does not push the re-used B operand (the feature matrix) out of L2
      const uint64_t a_policy = (EPI == EPI_STORE && (ep.tune & kTuneStreamAEvictFirst)) ? l2_policy_evict_first()
                                : (ep.tune & kTuneAEvictLast)                            ? l2_policy_evict_last()
                                                                                         : 0;
      const uint64_t b_policy = (ep.tune & kTuneBEvictLast) ? l2_policy_evict_last() : 0;
      // one 64-element-wide box of operand X into this CTA's shared memory; pair mode credits the leader's barrier
      auto load_box = [&](uint32_t dst, const CUtensorMap* tm, uint32_t fb, int c0, int c1, uint64_t policy) {
        if constexpr (CG == 2) {
          if (policy) tma_load_2d_pair_hint(dst, tm, fb, c0, c1, policy);
          else        tma_load_2d_pair(dst, tm, fb, c0, c1);
        } else {
          if (policy) tma_load_2d_hint(dst, tm, fb, c0, c1, policy);
          else        tma_load_2d(dst, tm, fb, c0, c1);
        }
      };
      // m_blk: this CTA's own 128-row block
      auto load_a = [&](uint32_t sa, uint32_t fb, int m_blk, int kb) {
        if constexpr (A_MN) {
#pragma unroll
          for (int u = 0; u < BM / 64; ++u) load_box(sa + u * (BK * 128), &tmA, fb, m_blk * BM + u * 64, kb * BK, a_policy);
        } else {
          load_box(sa, &tmA, fb, kb * BK, m_blk * BM, a_policy);
        }
      };
      auto load_b = [&](uint32_t sb, uint32_t fb, int n_blk, int kb) {
        if constexpr (CG == 2) {
          // this CTA's half of the B tile (N rows [rank*128, rank*128 + 128) of the 256), stored at the start of its
          // own B slot: the pair MMA reads the two halves from the two CTAs
          if constexpr (B_MN) {
            constexpr int per = BN / 64 / 2;
#pragma unroll
            for (int uu = 0; uu < per; ++uu)
              load_box(sb + uu * (BK * 128), &tmB, fb, n_blk * BN + ((int)cta_rank * per + uu) * 64, kb * BK, b_policy);
          } else {
            load_box(sb, &tmB, fb, kb * BK, n_blk * BN + (int)cta_rank * (BN / 2), b_policy);
          }
        } else if constexpr (CL > 1) {
          // this CTA's 1/CL slice of the shared B tile, multicast to the whole cluster
          if constexpr (B_MN) {
            constexpr int per = BN / 64 / CL;
#pragma unroll
            for (int uu = 0; uu < per; ++uu) {
              const int u = (int)cta_rank * per + uu;
              tma_load_2d_mc(sb + u * (BK * 128), &tmB, fb, n_blk * BN + u * 64, kb * BK, kClMask);
            }
          } else {
            constexpr int rows = BN / CL;
            tma_load_2d_mc(sb + cta_rank * (rows * 128), &tmB, fb, kb * BK, n_blk * BN + (int)cta_rank * rows, kClMask);
          }
        } else if constexpr (B_MN) {
#pragma unroll
          for (int u = 0; u < BN / 64; ++u) load_box(sb + u * (BK * 128), &tmB, fb, n_blk * BN + u * 64, kb * BK, b_policy);
        } else {
          load_box(sb, &tmB, fb, kb * BK, n_blk * BN, b_policy);
        }
      };
      auto prefetch = [&](int m_blk, int n_blk, int kb) {
        if constexpr (A_MN) {
#pragma unroll
          for (int u = 0; u < BM / 64; ++u) tma_prefetch_2d(&tmA, m_blk * BM + u * 64, kb * BK);
        } else {
          tma_prefetch_2d(&tmA, kb * BK, m_blk * BM);
        }
        if constexpr (B_MN) {
#pragma unroll
          for (int u = 0; u < BN / 64; ++u) tma_prefetch_2d(&tmB, n_blk * BN + u * 64, kb * BK);
        } else {
          tma_prefetch_2d(&tmB, kb * BK, n_blk * BN);
        }
      };
      int slab_seen = -1;
      const int epoch = (streamed && ep.ready_flags) ? ep.epoch_p[0] : 0;
      for (int tile = tile_first; tile < num_tiles; tile += tile_step) {
        int m_blk, n_blk, s_idx;
        const int t_idx = tile / S, ks = tile - t_idx * S;
        const int kb_begin = ks * kb_per, num_k = min(num_k_all, kb_begin + kb_per);
        tile_coords(t_idx, num_m, num_n, slab, slab_order, m_blk, n_blk, &s_idx, kGroup);
        if constexpr (CG == 2) m_blk = m_blk * 2 + (int)cta_rank;
        if (streamed && s_idx != slab_seen) {
          // first tile of this CTA in a new column block: its rows of the gather buffer must have landed (the copy
          // engine writes the flag right behind the block); order the TMA (async proxy) reads after the acquire
          if (ep.ready_flags && ep.blk_wait[s_idx]) {
            wait_flag_eq(ep.ready_flags + ep.blk_order[s_idx], epoch);
            asm volatile("fence.proxy.async;" ::: "memory");
          }
          slab_seen = s_idx;
        }
        int m_nxt = -1, n_nxt = -1;
        if (pf_dist && tile + tile_step < num_tiles)
          tile_coords(tile + tile_step, num_m, num_n, slab, slab_order, m_nxt, n_nxt, nullptr, kGroup);
        for (int kb = kb_begin; kb < num_k; ++kb) {
          if (pf_dist == 0xff) {
            // de-duplicated next-tile prefetch: the CTAs of a wave share operand tiles (16 CTAs per B tile, ~9 per A
            // tile), so only ONE of them asks L2 for each tile of the wave's next position, a whole tile ahead:
            // the B tile by the CTA whose next tile is the first m of its group, the A tile by the CTAs whose next
            // tile opens a group pass (first n of the slab) or that sit in the first kGroupM slots of the wave
            if (m_nxt >= 0) {
              if (m_nxt % kGroupM == 0) {
                if constexpr (B_MN) {
#pragma unroll
                  for (int u = 0; u < BN / 64; ++u) tma_prefetch_2d(&tmB, n_nxt * BN + u * 64, kb * BK);
                } else {
                  tma_prefetch_2d(&tmB, kb * BK, n_nxt * BN);
                }
              }
              if (n_nxt % slab == 0 || (int)blockIdx.x < kGroupM) {
                if constexpr (A_MN) {
#pragma unroll
                  for (int u = 0; u < BM / 64; ++u) tma_prefetch_2d(&tmA, m_nxt * BM + u * 64, kb * BK);
                } else {
                  tma_prefetch_2d(&tmA, kb * BK, m_nxt * BM);
                }
              }
            }
          } else if (pf_dist) {
            // naive variant (every CTA prefetches its own operands pf_dist k-blocks ahead): measured 1.4x SLOWER --
            // it doubles the L2 request rate because 16 / 9 CTAs ask for the same tile
            const int pk = kb + pf_dist;
            if (pk < num_k) prefetch(m_blk, n_blk, pk);
            else if (m_nxt >= 0 && pk - num_k < num_k) prefetch(m_nxt, n_nxt, pk - num_k);
          }
          mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1);
          const uint32_t fb = smem_u32(&full_bar[stage]);
          // pair mode: the leader expects both CTAs' bytes on its barrier (the peer's boxes may complete first: the
          // transaction count is signed within a phase)
          if (CG == 1 || cta_rank == 0) mbar_arrive_expect_tx(fb, kStageBytes * CG);
          const uint32_t sa = smem_u32(stage_base + stage * kStageBytes);
          load_a(sa, fb, m_blk, kb);
          load_b(sa + kABytes, fb, n_blk, kb);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1 && (CG == 1 || cta_rank == 0)) {
    // ===================================== MMA issuer =======================================
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc_bf16_f32(BM * CG, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
      // K-major : rows of 64 K-elements (128 B), 8-row swizzle atoms 1024 B apart (SBO); LBO unused
      // MN-major: 64 MN-elements per 128 B row, one row per K index; 8 K-rows = one 1024 B atom (SBO),
      //           the next 64 MN-elements live in the next TMA box, BK*128 B further (LBO)
      constexpr uint64_t desc_hi_k = make_smem_desc_hi(16, 1024, kSwizzle128B);
      constexpr uint64_t desc_hi_mn = make_smem_desc_hi(BK * 128, 1024, kSwizzle128B);
      constexpr uint64_t a_hi = A_MN ? desc_hi_mn : desc_hi_k;
      constexpr uint64_t b_hi = B_MN ? desc_hi_mn : desc_hi_k;
      // descriptor start-address step per 16-wide K slice (encoded >> 4): 32 B inside a K-major swizzle row,
      // 16 rows * 128 B for MN-major
      constexpr uint64_t a_step = A_MN ? (16 * 128) >> 4 : 2;
      constexpr uint64_t b_step = B_MN ? (16 * 128) >> 4 : 2;
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = tile_first; tile < num_tiles; tile += tile_step) {
        mbar_wait(smem_u32(&tempty_bar[acc]), acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        const int ks = tile % S;
        const int num_k = min(num_k_all, ks * kb_per + kb_per) - ks * kb_per;     // k-blocks of this slice (>= 1)
        for (int kb = 0; kb < num_k; ++kb) {
          mbar_wait(smem_u32(&full_bar[stage]), phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(stage_base + stage * kStageBytes);
          const uint64_t da = make_smem_desc(sa, a_hi);
          const uint64_t db = make_smem_desc(sa + kABytes, b_hi);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            if constexpr (CG == 2)
              umma_bf16_pair(d_tmem, da + (uint64_t)k * a_step, db + (uint64_t)k * b_step, idesc, (uint32_t)((kb | k) != 0));
            else
              umma_bf16(d_tmem, da + (uint64_t)k * a_step, db + (uint64_t)k * b_step, idesc, (uint32_t)((kb | k) != 0));
          }
          // frees the smem slot when the MMAs retire -- in every CTA of the cluster, whose producers all write it
          if constexpr (CG == 2)     umma_commit_pair(smem_u32(&empty_bar[stage]), kClMask);
          else if constexpr (CL > 1) umma_commit_mc(smem_u32(&empty_bar[stage]), kClMask);
          else                       umma_commit(smem_u32(&empty_bar[stage]));
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        // accumulator ready for the epilogue (pair: of both CTAs, each waits on its own barrier)
        if constexpr (CG == 2) umma_commit_pair(smem_u32(&tfull_bar[acc]), kClMask);
        else                   umma_commit(smem_u32(&tfull_bar[acc]));
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else if (warp >= 4) {
    // ===================================== epilogue =========================================
    const int e = warp - 4;
    const int q = warp & 3;            // TMEM lane quadrant this warp may access
    const int h = e >> 2;              // which 128-column half of the accumulator
    const int bar_id = 1 + h;          // named barrier shared by the 4 warps of one half
    int acc = 0;
    uint32_t acc_phase = 0;
    float dscale_acc = 0.f;
    [[maybe_unused]] float sig_acc[3] = {0.f, 0.f, 0.f};
    [[maybe_unused]] const uint64_t store_policy = (ep.tune & kTuneStoreEvictFirst) ? l2_policy_evict_first() : 0;
    for (int tile = tile_first; tile < num_tiles; tile += tile_step) {
      int m_blk, n_blk;
      const int t_idx = tile / S;
      [[maybe_unused]] const int ks = tile - t_idx * S;
      tile_coords(t_idx, num_m, num_n, slab, slab_order, m_blk, n_blk, nullptr, kGroup);
      if constexpr (CG == 2) m_blk = m_blk * 2 + (int)cta_rank;
      const int gi = m_blk * BM + q * 32 + lane;          // this thread's row
      const int n_base = n_blk * BN + h * 128;            // first column of this warp's half
      // interior tiles (no ragged edge in this warp's 32 x 128 block) take the branch-free code
      const bool interior = (m_blk * BM + q * 32 + 32 <= M) && (n_base + 128 <= N);
      float* scratch = epi_scratch + ((acc * 2 + h) * 4) * 128;   // [4][128] for this (acc, half)
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + h * 128);
      const uint32_t tfull = smem_u32(&tfull_bar[acc]);
      const uint32_t tempty = smem_u32(&tempty_bar[acc]);

      if (CG == 2 && m_blk * BM >= M) {
        // odd number of 128-row blocks: the pair's second CTA holds no rows of the last tile row.  It still takes part
        // in the accumulator hand-shake (the leader's MMA wrote zeros into its tensor memory).
        mbar_wait(tfull, acc_phase);
        tc_fence_after();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) arrive_tempty<CG>(tempty);
      } else if constexpr (EPI == EPI_LSE) {
        mbar_wait(tfull, acc_phase);
        tc_fence_after();
        if (interior) lse_tile<true, CG>(ep, taddr, scratch, gi, n_base, m_blk, n_blk, h, q, lane, M, N, tempty, bar_id);
        else          lse_tile<false, CG>(ep, taddr, scratch, gi, n_base, m_blk, n_blk, h, q, lane, M, N, tempty, bar_id);
      } else if constexpr (EPI == EPI_DS) {
        // stage the per-column terms of this half in smem before touching TMEM:
        //   [0,128)   two-exp path: exponent offsets, w_col * 2^(v - cl2) = 2^(v - (cl2 - log2 w_col));
        //             w_col == 0 or a column past N -> offset +inf -> the term vanishes
        //   [128,256) one-exp path: factors 2^(nu - cl2_j), nu = cl2 of the half's first column
        //   [256,264) per-warp max / min of cl2 over the valid columns (range check of the one-exp path)
        const bool has_col = ep.w_col > 0.f;
        float nu = 0.f;
        {
          const int c = q * 32 + lane;
          const int gj = n_base + c;
          const bool cvalid = gj < N;
          const float lwc = has_col ? fast_log2(ep.w_col) : -INFINITY;
          const float cl2 = cvalid ? ep.col_lse[gj] * kLog2e : INFINITY;
          if (has_col) nu = ep.col_lse[min(n_base, N - 1)] * kLog2e;
          scratch[c] = cl2 - lwc;
          scratch[128 + c] = (cvalid && has_col) ? fast_exp2(nu - cl2) : 0.f;
          const float cmx = warp_max(cvalid ? cl2 : -INFINITY);
          const float cmn = -warp_max(cvalid ? -cl2 : -INFINITY);
          if (lane == 0) {
            scratch[256 + q] = cmx;
            scratch[260 + q] = cmn;
          }
        }
        const float g = ep.grad_out[0];
        const float rl2 = (gi < M) ? ep.row_lse[gi] * kLog2e : INFINITY;
        const float rmx = warp_max((gi < M) ? rl2 : -INFINITY);
        const float rmn = -warp_max((gi < M) ? -rl2 : -INFINITY);
        named_bar_sync(bar_id, 128);
        bool fast = (ep.tune & kTuneDsTwoExp) == 0;
        if (fast && has_col) {
          const float4 mx4 = *reinterpret_cast<const float4*>(scratch + 256);
          const float4 mn4 = *reinterpret_cast<const float4*>(scratch + 260);
          const float cmx = fmaxf(fmaxf(mx4.x, mx4.y), fmaxf(mx4.z, mx4.w));
          const float cmn = fminf(fminf(mn4.x, mn4.y), fminf(mn4.z, mn4.w));
          // every |rl2_i - cl2_j| of this warp's 32 x 128 block within range (NaN from inf - inf -> two-exp path)
          fast = (rmx - cmn <= kDsFastRange) && (cmx - rmn <= kDsFastRange);
        }
        mbar_wait(tfull, acc_phase);
        tc_fence_after();
        const uint32_t ostage = smem_u32(out_stage + e * 2048);
        const int row_blk = m_blk * BM + q * 32;
        if (fast) {
          const float nrl2 = -rl2;                                                 // OOB row: -inf -> e = 0
          const float gwr = g * ep.w_row;
          const float gwc_r = (has_col && gi < M) ? g * ep.w_col * fast_exp2(rl2 - nu) : 0.f;
          if (interior)
            dscale_acc += ds_tile<true, true>(ep, &tmC, taddr, scratch, ostage, gi, row_blk, n_base, nrl2, gwr, gwc_r,
                                              store_policy, lane, M, N);
          else
            dscale_acc += ds_tile<false, true>(ep, &tmC, taddr, scratch, ostage, gi, row_blk, n_base, nrl2, gwr, gwc_r,
                                               store_policy, lane, M, N);
        } else {
          const float lwr = (ep.w_row > 0.f) ? fast_log2(ep.w_row) : -INFINITY;
          const float nrl2 = (gi < M) ? -(rl2 - lwr) : -INFINITY;                  // -(rl2 - log2 w_row)
          if (interior)
            dscale_acc += ds_tile<true, false>(ep, &tmC, taddr, scratch, ostage, gi, row_blk, n_base, nrl2, 0.f, 0.f,
                                               store_policy, lane, M, N);
          else
            dscale_acc += ds_tile<false, false>(ep, &tmC, taddr, scratch, ostage, gi, row_blk, n_base, nrl2, 0.f, 0.f,
                                                store_policy, lane, M, N);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) arrive_tempty<CG>(tempty);
      } else if constexpr (EPI == EPI_SIG) {
        mbar_wait(tfull, acc_phase);
        tc_fence_after();
        const uint32_t ostage = smem_u32(out_stage + e * 2048);
        if (interior) sig_tile<true>(ep, &tmC, taddr, ostage, gi, m_blk * BM + q * 32, n_base, lane, M, N, sig_acc);
        else          sig_tile<false>(ep, &tmC, taddr, ostage, gi, m_blk * BM + q * 32, n_base, lane, M, N, sig_acc);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) arrive_tempty<CG>(tempty);
      } else if constexpr (EPI == EPI_BIAS) {
        mbar_wait(tfull, acc_phase);
        tc_fence_after();
        bias_tile(ep, &tmC, taddr, smem_u32(out_stage + e * 2048), m_blk * BM + q * 32, n_base, lane, N);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) arrive_tempty<CG>(tempty);
      } else {
        mbar_wait(tfull, acc_phase);
        tc_fence_after();
        store_tile(ep, taddr, gi, n_base, M, N, (S > 1) ? (long)ks * ep.split_stride : 0L);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) arrive_tempty<CG>(tempty);
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if constexpr (EPI == EPI_BIAS) {
      if (lane == 0) tma_store_wait_all();
    }
    if constexpr (EPI == EPI_SIG) {
      if (lane == 0) tma_store_wait_all();
      const float l0 = warp_sum(sig_acc[0]), l1 = warp_sum(sig_acc[1]), l2 = warp_sum(sig_acc[2]);
      if (lane == 0) {
        ep.loss_part[blockIdx.x * kEpiWarps + e] = l0;
        ep.dscale_part[blockIdx.x * kEpiWarps + e] = l1;
        ep.dbias_part[blockIdx.x * kEpiWarps + e] = l2;
      }
    }
    if constexpr (EPI == EPI_DS) {
      if (lane == 0) tma_store_wait_all();               // dS tiles fully written before the kernel ends
      dscale_acc = warp_sum(dscale_acc);
      if (lane == 0) ep.dscale_part[blockIdx.x * kEpiWarps + e] = dscale_acc;
    }
  }

  tc_fence_before();
  __syncthreads();
  if constexpr (kClustered) cluster_sync();        // no CTA exits while a peer may still multicast into it / read its smem
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<CG>(tmem_base, kTmemCols);
  }
}

// ---- host side ----------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
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

// bf16 row-major [outer, inner] matrix with row stride ld (elements); TMA box = [box_outer rows, 64 inner], 128B swizzle
static int make_tmap_bf16(CUtensorMap* tm, const void* base, int outer, int inner, long ld, int box_outer) {
  PFN_encodeTiled enc = get_encode();
  XTAG_REQUIRE(enc != nullptr, XTAG_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  XTAG_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0 && (ld * 2) % 16 == 0, XTAG_ERR_UNSUPPORTED,
               "TMA operand must be 16-byte aligned with a 16-byte multiple row stride (ld=%ld)", ld);
  cuuint64_t gdim[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  cuuint64_t gstride[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {64u, (cuuint32_t)box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  XTAG_REQUIRE(r == CUDA_SUCCESS, XTAG_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return XTAG_OK;
}

// acc[M,N] = sum_k A(m,k) B(n,k).  K-major operand: memory [M|N rows][K], row stride ld.
//                                   MN-major operand: memory [K rows][M|N], row stride ld.
// bf16 row-major [rows][cols] output written by 32 x 32 TMA tile stores (64-byte swizzle), clipped at the edges
static int make_store_tmap_bf16(CUtensorMap* tm, void* base, int rows, int cols, long ld) {
  PFN_encodeTiled enc = get_encode();
  XTAG_REQUIRE(enc != nullptr, XTAG_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {32u, 32u};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  XTAG_REQUIRE(r == CUDA_SUCCESS, XTAG_ERR_CUDA, "cuTensorMapEncodeTiled(store) failed with CUresult %d", (int)r);
  return XTAG_OK;
}

// per-device "already done" flags of one kernel instantiation (one process may drive several GPUs)
struct PerDevice {
  int v[64] = {};
  int& here() {
    int dev = 0;
    cudaGetDevice(&dev);
    return v[dev & 63];
  }
};

template <int EPI, bool A_MN, bool B_MN, int CL, int CG>
static int launch_tc_cl(const void* A, long lda, const void* B, long ldb, int M, int N, int K, const EpiParams& ep,
                        cudaStream_t st, int grid, int* grid_used) {
  constexpr int kSmemBytes = Geo<CG>::kSmemBytes;
  constexpr int kCluster = CL * CG;
  XTAG_CUDA(sync_spin_timeout());
  CUtensorMap tmA, tmB, tmC;
  memset(&tmC, 0, sizeof(tmC));
  if (EPI == EPI_DS || (EPI == EPI_SIG && ep.sig_store)) {
    int rcC = make_store_tmap_bf16(&tmC, ep.dS, M, N, ep.ldds);
    if (rcC) return rcC;
  }
  if (EPI == EPI_BIAS) {
    int rcC = make_store_tmap_bf16(&tmC, ep.C, M, N, ep.ldc);
    if (rcC) return rcC;
  }
  int rc = A_MN ? make_tmap_bf16(&tmA, A, K, M, lda, BK) : make_tmap_bf16(&tmA, A, M, K, lda, BM);
  if (rc) return rc;
  // K-major B: each CTA of a cluster fetches BN / CL rows of the shared tile
  rc = B_MN ? make_tmap_bf16(&tmB, B, K, N, ldb, BK) : make_tmap_bf16(&tmB, B, N, K, ldb, BN / kCluster);
  if (rc) return rc;
  static PerDevice attr_set;
  if (!attr_set.here()) {
    XTAG_CUDA(cudaFuncSetAttribute(tc_gemm_kernel<EPI, A_MN, B_MN, CL, CG>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   kSmemBytes));
    attr_set.here() = 1;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid, 1, 1);
  cfg.blockDim = dim3(kThreads, 1, 1);
  cfg.dynamicSmemBytes = kSmemBytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kCluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = kCluster > 1 ? 1 : 0;
  if (kCluster > 1) {
    // a persistent kernel needs every cluster co-resident: GPCs with a number of free SMs that is not a multiple of
    // the cluster size leave SMs unused, so the grid is sized by what the device can actually hold
    static PerDevice max_clusters;
    if (max_clusters.here() == 0) {
      int n = 0;
      XTAG_CUDA(cudaOccupancyMaxActiveClusters(&n, tc_gemm_kernel<EPI, A_MN, B_MN, CL, CG>, &cfg));
      XTAG_REQUIRE(n > 0, XTAG_ERR_CUDA, "no cluster of %d CTAs with %d bytes of shared memory fits on this device",
                   kCluster, kSmemBytes);
      max_clusters.here() = n;
    }
    if (grid > max_clusters.here() * kCluster) grid = max_clusters.here() * kCluster;
    cfg.gridDim = dim3((unsigned)grid, 1, 1);
  }
  if (grid_used) *grid_used = grid;
  {
    ProfScope prof(EPI, 2.0 * (double)M * (double)N * (double)K, st);
    XTAG_CUDA(cudaLaunchKernelEx(&cfg, tc_gemm_kernel<EPI, A_MN, B_MN, CL, CG>, tmA, tmB, tmC, M, N, K, ep));
  }
  XTAG_CHECK_LAUNCH();
  return XTAG_OK;
}

// Cluster size for a problem: the CL CTAs of a cluster take CL consecutive tiles of the static schedule, which are CL
// consecutive m tiles of one n tile exactly when the number of m tiles is a multiple of CL (tile_coords groups m first).
static int pick_cluster(int M, int N, int tune) {
  const int want = (tune & kTuneCluster4) ? 4 : (tune & kTuneCluster2) ? 2 : 1;
  const int num_m = (M + BM - 1) / BM, num_n = (N + BN - 1) / BN;
  int cl = want;
  while (cl > 1 && (num_m % cl != 0 || (long)num_m * num_n < 2L * cl)) cl >>= 1;
  return cl;
}

// CTA-pair kernels whenever the problem has at least two 128-row blocks (tune bit 24 turns them off)
static bool pick_pair(int M, int tune) { return !(tune & kTuneNoPair) && M > BM; }

template <int EPI, bool A_MN, bool B_MN>
static int launch_tc(const void* A, long lda, const void* B, long ldb, int M, int N, int K, const EpiParams& ep_in,
                     cudaStream_t st, int* grid_used = nullptr) {
  EpiParams ep = ep_in;
  ep.tune = tc_tune();
  // Plain GEMMs with few n tiles (dA / dB of the contrastive head: N = D = 1024 -> 4 n tiles, K = 32768): visit the n
  // tiles of one m tile back to back.  With the default groups of 8 m x 4 n = 32 tiles a wave of 74 CTA pairs cuts
  // through a group, and the tiles on the far side re-read their A row blocks -- the 2 GiB dS -- from HBM one K loop
  // later (ncu: 4.4 GB read per launch against 2.2 GB algorithmic).
  // (not with the multicast clusters of the single-CTA kernels, whose CTAs take consecutive m tiles of one n tile)
  if (EPI == EPI_STORE && (N + BN - 1) / BN <= 8 && ep.group_m == 0 &&
      (pick_pair(M, ep.tune) || pick_cluster(M, N, ep.tune) == 1))
    ep.group_m = 1;
  const int num_n = (N + BN - 1) / BN * ((EPI == EPI_STORE && ep.split_k > 1) ? ep.split_k : 1);   // x K slices
  if (pick_pair(M, ep.tune)) {
    const int num_pairs = ((M + 2 * BM - 1) / (2 * BM)) * num_n;
    int grid = num_sms() & ~1;
    if (grid > 2 * num_pairs) grid = 2 * num_pairs;
    return launch_tc_cl<EPI, A_MN, B_MN, 1, 2>(A, lda, B, ldb, M, N, K, ep, st, grid, grid_used);
  }
  const int num_tiles = ((M + BM - 1) / BM) * num_n;
  const int cl = pick_cluster(M, N, ep.tune);
  int grid = num_sms();
  if (grid > num_tiles) grid = num_tiles;
  grid -= grid % cl;
  if (cl == 4) return launch_tc_cl<EPI, A_MN, B_MN, 4, 1>(A, lda, B, ldb, M, N, K, ep, st, grid, grid_used);
  if (cl == 2) return launch_tc_cl<EPI, A_MN, B_MN, 2, 1>(A, lda, B, ldb, M, N, K, ep, st, grid, grid_used);
  return launch_tc_cl<EPI, A_MN, B_MN, 1, 1>(A, lda, B, ldb, M, N, K, ep, st, grid, grid_used);
}

int launch_lse_reduce(const float* parts, int P, int n, float in_mul, float out_mul, float* out, cudaStream_t st);
int launch_sum_into(const float* parts, int n, float* out, cudaStream_t st);

// ---- split-K for the plain GEMMs -----------------------------------------------------------------------------------
// A GEMM with few output tiles and a long K (the tag head's projection weight gradient: [3072, 512] over K = b*N =
// 201 728; the gradient GEMMs of small-batch contrastive steps) leaves most SMs idle with one work item per tile:
// its K loop is cut into S slices that run as S x tiles work items and write raw fp32 partial slabs; one small kernel
// sums the slabs (fixed order: deterministic) and casts.
static int pick_split(int M, int N, int K, int tune) {
  if (tune & kTuneNoSplitK) return 1;
  const bool pair = pick_pair(M, tune);
  // (the multicast clusters of the single-CTA kernels take CONSECUTIVE m tiles of one n tile: not with K slices)
  if (!pair && (tune & (kTuneCluster2 | kTuneCluster4))) return 1;
  const int units = pair ? num_sms() / 2 : num_sms();
  const int tiles = ((M + (pair ? 2 : 1) * BM - 1) / ((pair ? 2 : 1) * BM)) * ((N + BN - 1) / BN);
  const int num_k = (K + BK - 1) / BK;
  if (tiles * 2 > units || num_k < 16) return 1;
  int S = units / tiles;
  if (S > num_k / 8) S = num_k / 8;
  if (S > 32) S = 32;
  if (S < 2) return 1;
  const int per = (num_k + S - 1) / S;
  return (num_k + per - 1) / per;                    // every slice owns at least one k-block
}
static size_t split_ws_bytes(int M, int N, int K, int tune) {
  const int S = pick_split(M, N, K, tune);
  return S > 1 ? (size_t)S * (size_t)M * (size_t)N * 4 : 0;
}

__global__ void __launch_bounds__(256) split_reduce_kernel(const float* __restrict__ slabs, int S, size_t stride, int M,
                                                           int N, void* __restrict__ out, long ldc, int out_bf16) {
  const size_t n = (size_t)M * N;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    float acc = 0.f;
    for (int k = 0; k < S; ++k) acc += slabs[(size_t)k * stride + i];
    const size_t r = i / N, c = i - r * N;
    if (out_bf16) reinterpret_cast<__nv_bfloat16*>(out)[r * ldc + c] = __float2bfloat16_rn(acc);
    else          reinterpret_cast<float*>(out)[r * ldc + c] = acc;
  }
}
}  // namespace xtag
extern "C" int xtag_lse_reduce2_log2(const float* parts0, int P0, int n0, float* out0, const float* parts1, int P1,
                                     int n1, float* out1, void* stream);
namespace xtag {

static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

// plain GEMM C = alpha * A * B^T with optional split-K through the caller's workspace
template <bool A_MN, bool B_MN>
static int gemm_store(const void* A, long lda, const void* B, long ldb, int M, int N, int K, const EpiParams& es,
                      void* split_ws, size_t split_bytes, cudaStream_t st) {
  const int S = pick_split(M, N, K, tc_tune());
  if (S > 1 && split_ws && split_bytes >= (size_t)S * M * N * 4) {
    EpiParams ep = es;
    ep.C = split_ws; ep.ldc = N; ep.c_is_bf16 = 0;
    ep.split_k = S; ep.split_stride = (long)M * N;
    int rc = launch_tc<EPI_STORE, A_MN, B_MN>(A, lda, B, ldb, M, N, K, ep, st);
    if (rc) return rc;
    size_t blocks = ((size_t)M * N + 255) / 256;
    if (blocks > (size_t)num_sms() * 8) blocks = (size_t)num_sms() * 8;
    split_reduce_kernel<<<(int)blocks, 256, 0, st>>>((const float*)split_ws, S, (size_t)M * N, M, N, es.C, es.ldc,
                                                     es.c_is_bf16);
    XTAG_CHECK_LAUNCH();
    return XTAG_OK;
  }
  return launch_tc<EPI_STORE, A_MN, B_MN>(A, lda, B, ldb, M, N, K, es, st);
}

size_t tc_fwd_ws(int M, int N) {
  const size_t num_m = (M + BM - 1) / BM, num_n = (N + BN - 1) / BN;
  return align256(2 * num_n * (size_t)M * 4) + align256(num_m * (size_t)N * 4) + 256;
}

int tc_clip_fwd(const void* A, const void* Bm, int M, int N, int D, const float* scale, int label_offset,
                float* row_lse, float* col_lse, float* diag, void* ws, size_t ws_bytes, cudaStream_t st) {
  XTAG_REQUIRE(ws && ws_bytes >= tc_fwd_ws(M, N), XTAG_ERR_WORKSPACE, "clip_fwd(tc): workspace %zu < %zu", ws_bytes,
               tc_fwd_ws(M, N));
  const int num_m = (M + BM - 1) / BM, num_n = (N + BN - 1) / BN;
  EpiParams ep = {};
  ep.scale_p = scale;
  ep.label_offset = label_offset;
  ep.row_part = (float*)ws;
  ep.col_part = (float*)((uint8_t*)ws + align256(2 * (size_t)num_n * M * 4));
  ep.col_ld = N;
  ep.diag = diag;
  int rc = launch_tc<EPI_LSE, false, false>(A, D, Bm, D, M, N, D, ep, st);
  if (rc) return rc;
  return xtag_lse_reduce2_log2(ep.row_part, 2 * num_n, M, row_lse, ep.col_part, num_m, N, col_lse, st);
}

// One column block of a forward whose reductions are deferred (chunk-pipelined multi-GPU gather): the kernel only
// writes its log2-domain partials -- row_part [2*ceil(N/256)][M] (this block's slice of the caller's row-partial
// buffer), col_part [ceil(M/128)][col_ld] at the block's column offset -- and the caller reduces all blocks at once.
void tc_fwd_block_parts(int M, int N, int* row_parts, int* col_parts) {
  *row_parts = 2 * ((N + BN - 1) / BN);
  *col_parts = (M + BM - 1) / BM;
}
int tc_clip_fwd_block(const void* A, const void* Bm, int M, int N, int D, const float* scale, int label_offset,
                      float* row_part, float* col_part, int col_ld, float* diag, cudaStream_t st) {
  EpiParams ep = {};
  ep.scale_p = scale;
  ep.label_offset = label_offset;
  ep.row_part = row_part;
  ep.col_part = col_part;
  ep.col_ld = col_ld;
  ep.diag = diag;
  return launch_tc<EPI_LSE, false, false>(A, D, Bm, D, M, N, D, ep, st);
}

// Streamed forward: one persistent launch over the whole [N, D] gather buffer whose column blocks (blk_cols rows of Bm
// each) arrive one after the other; block order[k] is visited k-th and (wait[k]) only after ready_flags[order[k]] ==
// *epoch.  Same partial layout as tc_clip_fwd: row_part [2*num_n][M], col_part [num_m][col_ld].
int tc_clip_fwd_stream(const void* A, const void* Bm, int M, int N, int D, const float* scale, int label_offset,
                       const int* order, const int* wait, int nblk, int blk_cols, const int* ready_flags,
                       const int* epoch, float* row_part, float* col_part, int col_ld, float* diag, cudaStream_t st) {
  XTAG_REQUIRE(nblk >= 1 && nblk <= 16 && blk_cols % BN == 0 && (long)nblk * blk_cols == (long)N, XTAG_ERR_UNSUPPORTED,
               "clip_fwd_stream: needs <= 16 column blocks of a multiple of %d columns (nblk=%d, blk_cols=%d, N=%d)", BN,
               nblk, blk_cols, N);
  EpiParams ep = {};
  ep.scale_p = scale;
  ep.label_offset = label_offset;
  ep.row_part = row_part;
  ep.col_part = col_part;
  ep.col_ld = col_ld;
  ep.diag = diag;
  ep.ready_flags = ready_flags;
  ep.epoch_p = epoch;
  ep.blk_tiles = blk_cols / BN;
  bool seen[16] = {};
  for (int k = 0; k < nblk; ++k) {
    XTAG_REQUIRE(order[k] >= 0 && order[k] < nblk && !seen[order[k]], XTAG_ERR_INVALID,
                 "clip_fwd_stream: order must be a permutation of the column blocks");
    seen[order[k]] = true;
    ep.blk_order[k] = order[k];
    ep.blk_wait[k] = wait[k] ? 1 : 0;
  }
  return launch_tc<EPI_LSE, false, false>(A, D, Bm, D, M, N, D, ep, st);
}

// Backward workspace: dS [M][Np] bf16 (Np = N padded to 8 so rows stay 16-byte multiples) + d(logit_scale) partials.
// The two gradient GEMMs read dS and the features in place through MN-major UMMA descriptors.
struct BwdLayout {
  size_t Np, off_ds, off_part, off_split, split_bytes, total;
};
static BwdLayout bwd_layout(int M, int N, int D) {
  BwdLayout L;
  L.Np = ((size_t)N + 7) & ~(size_t)7;
  size_t o = 0;
  L.off_ds = o;   o += align256((size_t)M * L.Np * 2);
  L.off_part = o; o += align256((size_t)3 * 256 * kEpiWarps * 4);    // per-warp partials (the sigmoid loss has three)
  // split-K slabs of the gradient GEMMs (small problems only: large ones have enough tiles).  Sized for the
  // default tuning so that the workspace size does not depend on a runtime knob.
  const size_t sa = split_ws_bytes(M, D, (int)L.Np, 0), sb = split_ws_bytes(N, D, M, 0);
  L.split_bytes = sa > sb ? sa : sb;
  L.off_split = o; o += align256(L.split_bytes);
  L.total = o + 256;
  return L;
}
size_t tc_bwd_ws(int M, int N, int D) { return bwd_layout(M, N, D).total; }

int tc_clip_bwd(const void* A, const void* Bm, int M, int N, int D, const float* scale, int label_offset,
                const float* row_lse, const float* col_lse, float w_row, float w_col, float w_diag,
                const float* grad_out, void* dA, void* dB, int grad_dtype, float* dscale,
                void* ws, size_t ws_bytes, int flags, cudaStream_t st) {
  const BwdLayout L = bwd_layout(M, N, D);
  XTAG_REQUIRE(ws && ws_bytes >= L.total, XTAG_ERR_WORKSPACE, "clip_bwd(tc): workspace %zu < %zu", ws_bytes, L.total);
  uint8_t* w = (uint8_t*)ws;
  __nv_bfloat16* dS = (__nv_bfloat16*)(w + L.off_ds);
  float* part = (float*)(w + L.off_part);
  int rc = XTAG_OK;
  if (!(flags & XTAG_BWD_REUSE_DS)) {
    // padding columns of dS (N not a multiple of 8) enter the K sum of dA: keep them zero
    if (L.Np != (size_t)N) XTAG_CUDA(cudaMemsetAsync(dS, 0, (size_t)M * L.Np * 2, st));
    EpiParams ep = {};
    ep.scale_p = scale;
    ep.label_offset = label_offset;
    ep.row_lse = row_lse;
    ep.col_lse = col_lse;
    ep.w_row = w_row; ep.w_col = w_col; ep.w_diag = w_diag;
    ep.grad_out = grad_out;
    ep.dS = dS; ep.ldds = (int)L.Np;
    ep.dscale_part = part;
    int grid = 0;
    rc = launch_tc<EPI_DS, false, false>(A, D, Bm, D, M, N, D, ep, st, &grid);
    if (rc) return rc;
    if (dscale) {
      rc = launch_sum_into(part, grid * kEpiWarps, dscale, st);
      if (rc) return rc;
    }
  }
  EpiParams es = {};
  es.ldc = D; es.c_is_bf16 = (grad_dtype == XTAG_BF16); es.alpha = 1.f; es.alpha_p = scale;
  if (dA) {
    // dA[i,d] = s * sum_j dS[i,j] Bm[j,d]: A operand dS is K-major (j contiguous); B operand (n=d, k=j) is the
    // feature matrix itself, [K=j rows][N=d contiguous] = MN-major
    es.C = dA;
    rc = gemm_store<false, true>(dS, (long)L.Np, Bm, (long)D, M, D, N, es, w + L.off_split, L.split_bytes, st);
    if (rc) return rc;
  }
  if (dB) {
    // dB[j,d] = s * sum_i dS[i,j] A[i,d]: A operand (m=j, k=i) is dS read as [K=i rows][M=j contiguous];
    // B operand (n=d, k=i) is A read as [K=i rows][N=d contiguous]: both MN-major, nothing is transposed
    es.C = dB;
    rc = gemm_store<true, true>(dS, (long)L.Np, A, (long)D, N, D, M, es, w + L.off_split, L.split_bytes, st);
    if (rc) return rc;
  }
  return XTAG_OK;
}

// Sigmoid loss forward: loss, d loss / d logit_scale and d loss / d logit_bias for a unit upstream gradient in out3, and
// (stage_ds) the logit gradient dS staged in ws exactly where tc_clip_bwd(XTAG_BWD_REUSE_DS) expects it.
int tc_siglip_fwd(const void* A, const void* Bm, int M, int N, int D, const float* scale, const float* bias,
                  int label_offset, float w, float* out3, void* ws, size_t ws_bytes, int stage_ds, cudaStream_t st) {
  const BwdLayout L = bwd_layout(M, N, D);
  XTAG_REQUIRE(ws && ws_bytes >= L.total, XTAG_ERR_WORKSPACE, "siglip_fwd: workspace %zu < %zu", ws_bytes, L.total);
  uint8_t* wsb = (uint8_t*)ws;
  __nv_bfloat16* dS = (__nv_bfloat16*)(wsb + L.off_ds);
  float* part = (float*)(wsb + L.off_part);
  if (stage_ds && L.Np != (size_t)N) XTAG_CUDA(cudaMemsetAsync(dS, 0, (size_t)M * L.Np * 2, st));
  XTAG_CUDA(cudaMemsetAsync(out3, 0, 3 * sizeof(float), st));
  EpiParams ep = {};
  ep.scale_p = scale;
  ep.bias_p = bias;
  ep.label_offset = label_offset;
  ep.w_sig = w;
  ep.sig_store = stage_ds ? 1 : 0;
  ep.dS = dS; ep.ldds = (int)L.Np;
  ep.loss_part = part;
  ep.dscale_part = part + 256 * kEpiWarps;
  ep.dbias_part = part + 2 * 256 * kEpiWarps;
  int grid = 0;
  int rc = launch_tc<EPI_SIG, false, false>(A, D, Bm, D, M, N, D, ep, st, &grid);
  if (rc) return rc;
  for (int k = 0; k < 3; ++k) {
    rc = launch_sum_into(part + k * 256 * kEpiWarps, grid * kEpiWarps, out3 + k, st);
    if (rc) return rc;
  }
  return XTAG_OK;
}

}  // namespace xtag

using namespace xtag;

// C[M,N] = alpha * sum_k A(m,k) B(n,k).  a_mn == 0: A is [M][K] row-major, a_mn == 1: A is [K][M] row-major;
// same for B with N.  Contiguous operands (leading dimension = inner extent).
extern "C" size_t xtag_tc_gemm_ws_bytes(int M, int N, int K) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  return split_ws_bytes(M, N, K, 0);
}

extern "C" int xtag_tc_gemm(const void* A, const void* B, void* C, int c_dtype, int M, int N, int K, float alpha,
                            int a_mn, int b_mn, void* stream) {
  return xtag_tc_gemm_ex(A, B, C, c_dtype, M, N, K, alpha, a_mn, b_mn, nullptr, 0, stream);
}

extern "C" int xtag_tc_gemm_ex(const void* A, const void* B, void* C, int c_dtype, int M, int N, int K, float alpha,
                               int a_mn, int b_mn, void* ws, size_t ws_bytes, void* stream) {
  XTAG_REQUIRE(A && B && C && M > 0 && N > 0 && K > 0, XTAG_ERR_INVALID, "tc_gemm: bad arguments");
  XTAG_REQUIRE((a_mn ? M : K) % 8 == 0 && (b_mn ? N : K) % 8 == 0, XTAG_ERR_UNSUPPORTED,
               "tc_gemm: the contiguous extent of each operand must be a multiple of 8 (16-byte TMA rows)");
  XTAG_REQUIRE(c_dtype == XTAG_F32 || c_dtype == XTAG_BF16, XTAG_ERR_INVALID, "tc_gemm: bad C dtype");
  int rc = xtag_device_check();
  if (rc) return rc;
  EpiParams es = {};
  es.C = C; es.ldc = N; es.c_is_bf16 = (c_dtype == XTAG_BF16); es.alpha = alpha; es.alpha_p = nullptr;
  cudaStream_t st = (cudaStream_t)stream;
  const long lda = a_mn ? M : K, ldb = b_mn ? N : K;
  if (!a_mn && !b_mn) return gemm_store<false, false>(A, lda, B, ldb, M, N, K, es, ws, ws_bytes, st);
  if (!a_mn && b_mn) return gemm_store<false, true>(A, lda, B, ldb, M, N, K, es, ws, ws_bytes, st);
  if (a_mn && b_mn) return gemm_store<true, true>(A, lda, B, ldb, M, N, K, es, ws, ws_bytes, st);
  set_error("tc_gemm: A MN-major with B K-major is not instantiated");
  return XTAG_ERR_UNSUPPORTED;
}

// Dense layer on the tcgen05 kernels: C[M, N] (bf16, row stride ldc) = A[M, K] (bf16, row stride lda) * W[N, K]^T + bias.
// W is a torch nn.Linear weight as it lies in memory ([out, in] = K-major B operand); several Linear layers that read
// the same input are fused by concatenating their weights along `out` (the tag head's K|V projections of both layers:
// reference tagging_heads/bert.py:208-209).
extern "C" int xtag_tc_linear_bf16(const void* A, long lda, const void* W, const float* bias, void* C, long ldc, int M,
                                   int N, int K, void* stream) {
  XTAG_REQUIRE(A && W && C && M > 0 && N > 0 && K > 0, XTAG_ERR_INVALID, "tc_linear: bad arguments");
  XTAG_REQUIRE(K % 8 == 0 && lda % 8 == 0 && ldc % 8 == 0 && lda >= K && ldc >= N, XTAG_ERR_UNSUPPORTED,
               "tc_linear: K and the row strides must be multiples of 8 elements (16-byte TMA rows)");
  XTAG_REQUIRE((reinterpret_cast<uintptr_t>(C) & 15) == 0, XTAG_ERR_UNSUPPORTED, "tc_linear: C must be 16-byte aligned");
  int rc = xtag_device_check();
  if (rc) return rc;
  EpiParams es = {};
  es.C = C; es.ldc = (int)ldc; es.c_is_bf16 = 1; es.alpha = 1.f; es.bias = bias;
  return launch_tc<EPI_BIAS, false, false>(A, lda, W, (long)K, M, N, K, es, (cudaStream_t)stream);
}

extern "C" int xtag_tc_gemm_nt(const void* A, const void* B, void* C, int c_dtype, int M, int N, int K, float alpha,
                               void* stream) {
  return xtag_tc_gemm(A, B, C, c_dtype, M, N, K, alpha, 0, 0, stream);
}

// Host-side view of the device tile schedule (tests/test_tile_schedule.py checks on the CPU that every tile is visited
// exactly once, that the CTAs of a cluster share their n tile, and that a streamed forward walks the column blocks in
// the requested order).  M, N in elements; slab in n tiles (0 = one slab); order = nullptr or a permutation of the
// N / (slab * 256) column blocks.
extern "C" int xtag_debug_tile_coords(int M, int N, int slab, const int* order, int tile, int* m_blk, int* n_blk,
                                      int* slab_idx) {
  const int num_m = (M + BM - 1) / BM, num_n = (N + BN - 1) / BN;
  if (M <= 0 || N <= 0 || tile < 0 || tile >= num_m * num_n || !m_blk || !n_blk) return XTAG_ERR_INVALID;
  const int sl = (slab > 0 && slab < num_n) ? slab : num_n;
  int s = 0;
  tile_coords(tile, num_m, num_n, sl, order, *m_blk, *n_blk, &s);
  if (slab_idx) *slab_idx = s;
  return XTAG_OK;
}

// Same for the CTA-pair kernels and the plain-GEMM variants: `rows_per_tile` = 128 (single CTA) or 256 (CTA pair: the
// returned m_blk counts 256-row blocks), group_m = m tiles per schedule group (0 = default, 1 = n-fastest order of the
// gradient GEMMs), split_k = K slices per tile (work item = tile * split_k + slice).
extern "C" int xtag_debug_work_item(int M, int N, int rows_per_tile, int slab, int group_m, int split_k, int item,
                                    int* m_blk, int* n_blk, int* k_slice) {
  if (M <= 0 || N <= 0 || (rows_per_tile != BM && rows_per_tile != 2 * BM) || !m_blk || !n_blk) return XTAG_ERR_INVALID;
  const int S = split_k > 1 ? split_k : 1;
  const int num_m = (M + rows_per_tile - 1) / rows_per_tile, num_n = (N + BN - 1) / BN;
  if (item < 0 || item >= num_m * num_n * S) return XTAG_ERR_INVALID;
  const int sl = (slab > 0 && slab < num_n) ? slab : num_n;
  const int grp = group_m > 0 ? group_m : kGroupM / (rows_per_tile / BM);
  tile_coords(item / S, num_m, num_n, sl, nullptr, *m_blk, *n_blk, nullptr, grp);
  if (k_slice) *k_slice = item % S;
  return XTAG_OK;
}

// cluster size the launcher would pick for an [M, N] problem under tuning bits `tune`
extern "C" int xtag_debug_pick_cluster(int M, int N, int tune) { return pick_cluster(M, N, tune); }
