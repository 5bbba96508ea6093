/*
 * xtag_b200.h -- C ABI of libxtag_b200.so: the B200 (sm_100a) kernels behind XTag-CLIP's
 * data-parallel hot path (open_clip contrastive head + XTag cross-attention tag head).
 *
 * The reference (EJLEE5826/XTag-CLIP) is pure Python/PyTorch and has NO FFI layer of its own
 * (SURVEY.md section 8b): the boundary it exposes is the Python API.  Each entry point below
 * therefore names the reference *Python* interface whose arithmetic it replaces (file:line under
 * /root/reference); the Python mirror that keeps those signatures lives in xtag_clip_b200/ and
 * binds this library with ctypes (INTEGRATION.md shows the stub a maintainer adds).
 *
 * Conventions (all entry points):
 *   - every pointer is a DEVICE pointer owned by the caller unless the name ends in _host;
 *     kernels never allocate: scratch is sized by *_ws_bytes() and passed in
 *   - stream is a cudaStream_t passed as void*; calls are asynchronous and stream-ordered,
 *     re-entrant, no global mutable state besides a thread-local last-error string
 *   - matrices are dense row-major; dtype codes below; rows must be 16-byte aligned
 *   - return 0 on success; XTAG_ERR_* (<0) otherwise, message via xtag_last_error()
 *   - NO CPU fallback: without a CUDA device of compute capability 10.x every compute call
 *     returns XTAG_ERR_CUDA.
 */
#ifndef XTAG_B200_H_
#define XTAG_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define XTAG_ABI_VERSION 1

#define XTAG_F32  0
#define XTAG_BF16 1

#define XTAG_OK               0
#define XTAG_ERR_INVALID     -1   /* bad argument (null pointer, negative size, bad dtype)   */
#define XTAG_ERR_UNSUPPORTED -2   /* shape / alignment / dtype combination not implemented  */
#define XTAG_ERR_CUDA        -3   /* CUDA runtime / driver error, or no sm_100 device        */
#define XTAG_ERR_WORKSPACE   -4   /* workspace too small                                     */

/* implementation selector for the contrastive kernels */
#define XTAG_IMPL_AUTO 0          /* bf16 + aligned shapes -> tcgen05, otherwise SIMT fp32    */
#define XTAG_IMPL_SIMT 1          /* fp32 FFMA path (any dtype, any shape): the "fp32 mode"   */
#define XTAG_IMPL_TC   2          /* tcgen05/TMEM/TMA path (bf16 inputs, D % 8 == 0)          */

int         xtag_version(void);
const char* xtag_last_error(void);
/* 0 when the current device is compute capability 10.x, XTAG_ERR_CUDA otherwise */
int         xtag_device_check(void);
/* number of kernels launched by this library in this process (bench.py's gpu_launches) */
uint64_t    xtag_launch_count(void);
/* Diagnostics for bench.py's roofline: while enabled, every tcgen05 launch is bracketed by CUDA events on
 * its stream.  xtag_prof_read synchronises them and returns up to `cap` records: tag (0 = K1 forward,
 * 1 = K2 dS producer, 2 = plain GEMM), duration in ms, algorithmic FLOPs (2*M*N*K).  enable(0/1) clears. */
/* Runtime tuning bits of the kernels (no reference counterpart; A/B measurements and diagnostics -- DESIGN.md
 * section 4 lists what each measured on B200):
 *   bits [0,8)   L2 prefetch distance of the TMA producer in 64-wide k-blocks (0 = off; 0xff = de-duplicated
 *                next-tile prefetch: one CTA per shared operand tile asks L2 a whole tile ahead -- not yet measured)
 *   bit 8        dS tile stores carry an L2 evict_first policy
 *   bit 9        the streamed A operand (staged dS) of the gradient GEMMs is loaded evict_first
 *   bit 10       force the two-exponential dS epilogue (default: one exponential per element when the block's
 *                row/column log-sum-exps are within 2^60 of each other, exact two-exp path otherwise)
 *   bit 11       K4 backward as ONE single-pass kernel (K and V streamed once, dK/dV by TMA tile stores) instead of
 *                the query-major dQ kernel + key-major dK/dV kernel pair                     [default: set]
 *   bits 12, 13  DIAGNOSTICS, wrong results: skip the dS staging + store / skip only the dS TMA store
 *   bits 14, 15  thread-block clusters of 2 / 4 CTAs along M, the shared B tile TMA-multicast to the cluster
 *   bits [16,24) n-slab width (in 256-column tiles) of the tile schedule (0 = one slab)
 *   bit 24       do NOT use the CTA-pair kernels (tcgen05.mma.cta_group::2, 256 x 256 tiles per pair of SMs, 6-stage
 *                operand ring); default: pair kernels whenever the problem has more than 128 rows
 *   bits 25, 26  L2 evict_last on the B / A operand tile loads
 *   bit 27       plain GEMMs: never split the K loop
 * Initial value: environment variable XTAG_TC_TUNE (0x200800 if unset).  set returns the previous value. */
int         xtag_set_tune(int bits);
int         xtag_get_tune(void);
/* Budget of the bounded device-side waits (mbarrier phases; the ready flags of xtag_clip_fwd_stream, where a rank
 * legitimately waits for slow peers).  A wait that exceeds it prints a diagnostic and traps, which fails the launch
 * instead of hanging the GPU.  Default 1 800 000 ms (longer than the 10 min default of the NCCL process group the
 * exchange replaces); initial value from XTAG_SPIN_TIMEOUT_MS.  Applies to launches made after the call; do not call
 * it inside a stream capture. */
int         xtag_set_spin_timeout_ms(long long ms);
int         xtag_prof_enable(int on);
int         xtag_prof_read(int* tags_host, float* ms_host, double* flops_host, int cap);

/* ---------------------------------------------------------------------------------------------
 * K3  fused L2-normalise + cast.   Replaces F.normalize(features, dim=-1) in
 *     CLIP.encode_image / encode_text   (src/open_clip/model.py:311-313, 332-333).
 *     y = x / max(||x||_2, eps); inv_norm[r] = 1/max(||x_r||, eps) is saved for the backward.
 *     yT (optional, may be NULL): the same result transposed, [dim, rows] in y_dtype.
 *     bwd: gx = inv_norm * (gy - y (y . gy)) when ||x|| >= eps, gy * inv_norm otherwise
 *     (the sub-gradient torch's clamp_min takes).
 * ------------------------------------------------------------------------------------------- */
int xtag_l2norm_fwd(const void* x, int x_dtype, void* y, int y_dtype, void* yT, float* inv_norm,
                    int rows, int dim, float eps, void* stream);
int xtag_l2norm_bwd(const void* gy, int gy_dtype, const void* y, int y_dtype, const float* inv_norm,
                    void* gx, int gx_dtype, int rows, int dim, float eps, void* stream);

/* ---------------------------------------------------------------------------------------------
 * K1  fused contrastive forward.   Replaces ClipLoss.get_logits + get_ground_truth +
 *     F.cross_entropy x2   (src/open_clip/loss.py:91-139) for one row block of the logits:
 *         S = scale * A Bm^T            A [M, D] (the rank's rows), Bm [N, D] (all columns)
 *     without ever writing S.  Outputs (fp32, natural-log units):
 *         row_lse[i] = LSE_j S_ij            (image->text CE of the local rows)
 *         col_lse[j] = LSE_{i<M} S_ij        (PARTIAL over this rank's rows; combine across
 *                                             ranks with xtag_lse_combine)
 *         diag[i]    = S_{i, i+label_offset} (the label logit; label_offset = b*rank,
 *                                             loss.py:95-96).  label_offset = -1: this column block holds no
 *                                             labels (chunk-pipelined gather), diag is left untouched.
 * ------------------------------------------------------------------------------------------- */
size_t xtag_clip_fwd_ws_bytes(int M, int N, int D, int dtype, int impl);
int xtag_clip_fwd(const void* A, const void* Bm, int dtype, int M, int N, int D,
                  const float* scale /* device scalar: logit_scale.exp() */, int label_offset,
                  float* row_lse, float* col_lse, float* diag,
                  void* ws, size_t ws_bytes, int impl, void* stream);

/* K1 with deferred reductions, for a forward that consumes its columns block by block (chunk-pipelined feature
 * gather): each xtag_clip_fwd_block launches only the fused GEMM of one column block
 * Bm_blk [N, D] and leaves log2-domain partials behind,
 *     row_part [row_parts][M]            row_parts = 2*ceil(N/256)   (xtag_clip_fwd_block_parts)
 *     col_part [col_parts][col_ld] + j0  col_parts = ceil(M/128), j0 = the block's first global column
 * The caller stacks the row partials of all blocks in one [sum row_parts][M] buffer, lets all blocks share one
 * [col_parts][col_ld = all columns] buffer, and finishes the step with two xtag_lse_reduce_log2 calls
 * (out[j] = ln sum_p 2^parts[p*n + j]).  tcgen05 path only (bf16, D % 8 == 0). */
int xtag_clip_fwd_block_parts(int M, int N, int* row_parts, int* col_parts);
int xtag_clip_fwd_block(const void* A, const void* Bm_blk, int dtype, int M, int N, int D,
                        const float* scale, int label_offset,
                        float* row_part, float* col_part, int col_ld, float* diag, void* stream);
int xtag_lse_reduce_log2(const float* parts, int P, int n, float* out, void* stream);
/* two such reductions (the row and the column partials of one forward) in ONE launch */
int xtag_lse_reduce2_log2(const float* parts0, int P0, int n0, float* out0,
                          const float* parts1, int P1, int n1, float* out1, void* stream);

/* K1 fused with the feature exchange: ONE persistent launch over the whole gather buffer Bm_all [N, D] whose nblk
 * column blocks (blk_cols rows each, a multiple of 256) are filled concurrently by copy-engine pulls from the peers.
 * The kernel visits the blocks in order_host[] and, for a block with wait_host[k] != 0, starts loading its tiles only
 * once ready_flags[block] == *epoch (a 4-byte copy the exchange enqueues right behind the block's data on the same copy
 * stream), so the transfer overlaps the math tile by tile with no launch per block.  Partials as xtag_clip_fwd_block:
 * row_part [2*ceil(N/256)][M], col_part [ceil(M/128)][col_ld]; finish with two xtag_lse_reduce_log2 calls.
 * label_offset is the global column of row 0's label (b*rank).  tcgen05 path only. */
int xtag_clip_fwd_stream(const void* A, const void* Bm_all, int dtype, int M, int N, int D,
                         const float* scale, int label_offset,
                         const int* order_host, const int* wait_host, int nblk, int blk_cols,
                         const int* ready_flags, const int* epoch,
                         float* row_part, float* col_part, int col_ld, float* diag, void* stream);

/* out[j] = log sum_w exp(parts[w*N + j]): merges the per-rank partial column LSEs after the
 * all-gather (the one exchange step of the sharded loss). */
int xtag_lse_combine(const float* parts, int W, int N, float* out, void* stream);

/* Peer-memory variants for the NVLink exchange steps (pointers in `*_dev` are DEVICE arrays of W device pointers,
 * one per rank, into peer-mapped symmetric buffers):
 *   xtag_lse_combine_ptrs : out[j] = log sum_w exp(parts[w][j])            (column-LSE exchange, forward)
 *   xtag_sum_ptrs_bf16    : out[i] = sum_w parts[w][i], n bf16 elements    (reduce step of the pull-based
 *                                                                           reduce-scatter of the text gradient) */
int xtag_lse_combine_ptrs(const float* const* parts_dev, int W, int N, float* out, void* stream);
/* xtag_lse_combine_ptrs + xtag_clip_loss (+ the epoch bump of the flag-gated forward) in one launch: col_out [N] =
 * combined column LSEs; loss_out[0] = this rank's loss over its M rows / label columns [label_offset, label_offset+M);
 * epoch (device int, may be NULL) is incremented once, for the next forward.  scratch: device buffer of
 * xtag_lse_combine_loss_scratch_bytes() bytes, zero-initialised ONCE by the caller and then owned by these launches
 * (per-block partials + a ticket that the last block resets). */
size_t xtag_lse_combine_loss_scratch_bytes(void);
int xtag_lse_combine_ptrs_loss(const float* const* parts_dev, int W, int N, float* col_out, const float* row_lse,
                               const float* diag, int M, int label_offset, float* loss_out, int* epoch,
                               void* scratch, void* stream);
int xtag_sum_ptrs_bf16(const void* const* parts_dev, int W, size_t n, void* out, void* stream);

/* loss = 0.5 * [ mean_i(row_lse_i - diag_i) + mean_i(col_lse[label_offset+i] - diag_i) ]
 * (loss.py:134-137 with the two cross-entropies written out).  loss_out[0] is overwritten. */
int xtag_clip_loss(const float* row_lse, const float* diag, const float* col_lse,
                   int M, int label_offset, float* loss_out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * K2  fused contrastive backward (autograd of loss.py:116-137).  Recomputes S tile by tile and
 *     forms   dS_ij = g * ( w_row * exp(S_ij - row_lse_i) + w_col * exp(S_ij - col_lse_j)
 *                           - w_diag * [j == i + label_offset] )
 *     with g = *grad_out (device scalar), then
 *         dA = scale * dS   Bm      [M, D]   (skipped when dA == NULL)
 *         dB = scale * dS^T A       [N, D]   (skipped when dB == NULL)
 *         dscale[0] += sum_ij dS_ij * S_ij / scale
 *     The weights select the reference's gradient mode (DESIGN.md "gradient modes").
 *     dS is staged once in the workspace (bf16 on the tcgen05 path, fp32 on the SIMT path).
 * ------------------------------------------------------------------------------------------- */
size_t xtag_clip_bwd_ws_bytes(int M, int N, int D, int dtype, int impl);
int xtag_clip_bwd(const void* A, const void* Bm, int dtype, int M, int N, int D,
                  const float* scale /* device scalar */, int label_offset,
                  const float* row_lse, const float* col_lse,
                  float w_row, float w_col, float w_diag, const float* grad_out,
                  void* dA, void* dB, int grad_dtype, float* dscale,
                  void* ws, size_t ws_bytes, int impl, int flags, void* stream);
/* flags for xtag_clip_bwd */
#define XTAG_BWD_REUSE_DS 1       /* ws already holds dS from a preceding call with the same operands and weights:
                                     skip the dS producer and d(logit_scale) (lets a caller launch the dB GEMM, start
                                     its reduce-scatter, and overlap it with the dA GEMM) */

/* Dense layer on the tcgen05 kernels (the tag head's K|V projections):
 *   C[M, N] (bf16, row stride ldc) = A[M, K] (bf16, row stride lda) * W[N, K]^T + bias[N] (fp32, may be NULL)
 * W is an nn.Linear weight as it lies in memory; Linear layers reading the same input are fused by concatenating their
 * weights along N (reference tagging_heads/bert.py:208-209: key and value of both layers = one [D -> 4*768] GEMM).
 * C leaves through TMA tile stores.  K, lda, ldc multiples of 8. */
int xtag_tc_linear_bf16(const void* A, long lda, const void* W, const float* bias, void* C, long ldc,
                        int M, int N, int K, void* stream);

/* plain tcgen05 GEMM used by K2 and exported for bring-up tests:
 *   C[M,N] = alpha * A[M,K] * B[N,K]^T   (A, B bf16 row-major "K-major"; C f32 or bf16; alpha by value) */
int xtag_tc_gemm_nt(const void* A, const void* B, void* C, int c_dtype,
                    int M, int N, int K, float alpha, void* stream);
/* general operand layouts: C[M,N] = alpha * sum_k A(m,k) B(n,k); a_mn = 0: A stored [M][K], 1: stored [K][M]
 * ("MN-major", read in place through an MN-major UMMA descriptor); b_mn likewise with N.  (a_mn, b_mn) in
 * {(0,0), (0,1), (1,1)}: the three layouts K2 uses (S = A Bm^T, dA = dS Bm, dB = dS^T A). */
int xtag_tc_gemm(const void* A, const void* B, void* C, int c_dtype,
                 int M, int N, int K, float alpha, int a_mn, int b_mn, void* stream);
/* Same with a workspace of xtag_tc_gemm_ws_bytes(M, N, K) bytes (0 = none needed): outputs with few 128/256 x 256
 * tiles and a long K are computed split-K -- S slices of the K loop run as S x tiles work items into fp32 partial slabs
 * in `ws`, one small kernel sums them in a fixed order.  Without a workspace the K loop is never split. */
size_t xtag_tc_gemm_ws_bytes(int M, int N, int K);
int xtag_tc_gemm_ex(const void* A, const void* B, void* C, int c_dtype, int M, int N, int K, float alpha,
                    int a_mn, int b_mn, void* ws, size_t ws_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * K4  tag-head cross-attention core.  Replaces the eager attention in
 *     BertSelfAttention.forward, cross branch (src/open_clip/tagging_heads/bert.py:219-274):
 *         P = softmax(q k^T * softmax_scale) ; P = dropout(P) ; ctx = P v     per (sample, head)
 *     q [b, Lq, heads*dh] (row stride ldq elements), k/v [b, Lk, heads*dh] (row strides ldk/ldv;
 *     batch strides = L * ld), o [b, Lq, heads*dh] contiguous, lse [b, heads, Lq] fp32.
 *     The all-ones encoder mask of tag_forward (model.py:339-341) is the additive constant 0.
 *     dropout_p == 0 is eval mode; otherwise a Philox4x32-7 keep-mask (7 rounds, csrc/philox.cuh) keyed by (seed, offset).
 * ------------------------------------------------------------------------------------------- */
int xtag_xattn_fwd(const void* q, const void* k, const void* v, int dtype,
                   void* o, float* lse,
                   int b, int Lq, int Lk, int heads, int dh,
                   int ldq, int ldk, int ldv,
                   float softmax_scale, float dropout_p, uint64_t seed, uint64_t offset,
                   void* stream);
/* dq [b,Lq,heads*dh], dk / dv [b,Lk,heads*dh] contiguous.  ws: scratch of at least b*heads*Lq*4 bytes
 * (rowsum(dO*O), shared by the two backward kernels). */
int xtag_xattn_bwd(const void* q, const void* k, const void* v, const void* o, const void* d_o,
                   const float* lse, int dtype,
                   void* dq, void* dk, void* dv,
                   int b, int Lq, int Lk, int heads, int dh,
                   int ldq, int ldk, int ldv,
                   float softmax_scale, float dropout_p, uint64_t seed, uint64_t offset,
                   void* ws, size_t ws_bytes, void* stream);
/* Same with dK / dV written as column slices of wider buffers (row strides lddk / lddv in elements), so that the
 * gradients of several attention layers that share one fused K|V projection buffer land in ONE [b*Lk, ...] gradient
 * buffer and feed one projection-backward GEMM.  bf16 single-pass kernel only (XTAG_ERR_UNSUPPORTED otherwise). */
int xtag_xattn_bwd_ld(const void* q, const void* k, const void* v, const void* o, const void* d_o, const float* lse,
                      int dtype, void* dq, void* dk, void* dv, int b, int Lq, int Lk, int heads, int dh, int ldq,
                      int ldk, int ldv, int lddk, int lddv, float softmax_scale, float dropout_p, uint64_t seed,
                      uint64_t offset, void* ws, size_t ws_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * K5  AsymmetricLoss forward + d/dx in one pass.  Replaces AsymmetricLoss.forward
 *     (src/open_clip/tagging_heads/asymmetric_loss.py:16-50); focal weight carries no gradient.
 *     loss_out[0] = -sum(loss); dx[n] (fp32) = d loss / d x for grad_out = 1 (may be NULL).
 *     idx6 (optional): the per-category top-1 tag indices of CLIP.prepare_control_words
 *     (src/open_clip/model.py:354-374), int32 [rows, 6]; requires cols == 44.
 * ------------------------------------------------------------------------------------------- */
int xtag_asl_fwd(const void* x, int x_dtype, const float* y, int rows, int cols,
                 float gamma_neg, float gamma_pos, float clip, float eps,
                 float* loss_out, float* dx, int32_t* idx6, void* stream);

/* Host-side views of the tcgen05 kernels' static tile schedule, for CPU tests (no device needed):
 * tile -> (m tile, n tile, slab) exactly as the device code computes it, and the cluster size the launcher picks. */
int xtag_debug_tile_coords(int M, int N, int slab, const int* order_host, int tile,
                           int* m_blk, int* n_blk, int* slab_idx);
int xtag_debug_pick_cluster(int M, int N, int tune);
/* work item -> (m block, n tile, K slice) of the CTA-pair kernels (rows_per_tile = 256: m_blk counts 256-row blocks)
 * or the single-CTA kernels (128), with the schedule's group size (0 = default; 1 = the n-fastest order of plain GEMMs
 * with few n tiles) and split-K (item = tile * split_k + slice) */
int xtag_debug_work_item(int M, int N, int rows_per_tile, int slab, int group_m, int split_k, int item,
                         int* m_blk, int* n_blk, int* k_slice);

/* ---- symmetric cross-entropy on a materialised square score matrix: DQNCOSLoss of the TQN fusion head ---------
 * (reference src/open_clip/tagging_heads/asymmetric_loss.py:54-65;  SURVEY.md section 8f rank 2)
 *   loss = 0.5 * [ mean_i(LSE_j X_ij - X_ii) + mean_j(LSE_i X_ij - X_jj) ],  X [n, n] fp32 or bf16, row stride ld
 *   dX   = grad_out * ( softmax_row(X) + softmax_col(X) - 2 I ) / (2 n)
 * fwd reads X once (row_lse, col_lse, diag [n] fp32 natural log are kept for bwd; ws = xtag_symm_ce_ws_bytes(n));
 * bwd reads X once and writes dX (same dtype as X, row stride lddx).  grad_out is a DEVICE scalar. */
size_t xtag_symm_ce_ws_bytes(int n);
int xtag_symm_ce_fwd(const void* x, int dtype, int n, long ld, float* row_lse, float* col_lse, float* diag,
                     float* loss_out, void* ws, size_t ws_bytes, void* stream);
int xtag_symm_ce_bwd(const void* x, int dtype, int n, long ld, const float* row_lse, const float* col_lse,
                     const float* grad_out, void* dx, long lddx, void* stream);

/* ---- K6: dense-output block of the tag head's BERT layers, fused ----------------------------------------------
 * (reference src/open_clip/tagging_heads/bert.py:281-292 BertSelfOutput, :359-370 BertOutput)
 *   y = LayerNorm( dropout(x) + resid[r % resid_rows] ) * gamma + beta          x, z, y bf16 [rows, H]; H in {256..1024}
 * fwd: one pass; keeps z (the pre-normalisation sum, bf16), mean and rstd [rows] fp32 for the backward.
 * bwd: one pass over dy (fp32 or bf16) and z -> dx, dresid (bf16 [rows, H]; the caller sums dresid over the samples
 *      when resid was broadcast), dgamma | dbeta as ONE [2, H] fp32 buffer (dbeta == dgamma + H).
 * The dropout keep-mask is Philox4x32-7 keyed by (seed, offset, element index): statistically equivalent to
 * torch.nn.Dropout, not bit-identical; dropout_p == 0 is eval mode. */
size_t xtag_ln_res_bwd_ws_bytes(int rows, int H);
int xtag_ln_res_fwd(const void* x, const void* resid, int resid_dtype, int resid_rows, const float* gamma,
                    const float* beta, void* z, void* y, float* mean, float* rstd, int rows, int H, float eps,
                    float dropout_p, uint64_t seed, uint64_t offset, void* stream);
int xtag_ln_res_bwd(const void* dy, int dy_dtype, const void* z, const float* mean, const float* rstd,
                    const float* gamma, void* dx, void* dresid, float* dgamma, float* dbeta, int rows, int H,
                    float dropout_p, uint64_t seed, uint64_t offset, void* ws, size_t ws_bytes, void* stream);

/* ---- sigmoid (SigLIP) loss on the contrastive head's mainloop ----------------------------------------------------
 * (reference src/open_clip/loss.py:314-448;  SURVEY.md section 8f rank 4)
 *   z_ij = scale * <A_i, Bm_j> + bias;  label_ij = +1 for j == i + label_offset, -1 otherwise (label_offset == -1: all
 *   -1, the reference's negative_only blocks);  loss = weight * sum_ij softplus(-label_ij * z_ij)
 * One pass: out3 = { loss, d loss / d scale, d loss / d bias } for a unit upstream gradient, and (stage_ds != 0) the
 * logit gradient dS staged in ws -- workspace of xtag_clip_bwd_ws_bytes(M, N, D, XTAG_BF16, XTAG_IMPL_TC) bytes, laid
 * out as xtag_clip_bwd expects it, so the feature gradients are
 *   xtag_clip_bwd(A, Bm, ..., scale = device scalar (upstream gradient * logit scale), ..., flags = XTAG_BWD_REUSE_DS)
 * i.e. two GEMMs: 6 M N D executed FLOPs per step in total, no recompute.  bf16, D % 8 == 0. */
int xtag_siglip_fwd(const void* A, const void* Bm, int dtype, int M, int N, int D, const float* scale, const float* bias,
                    int label_offset, float weight, float* out3, void* ws, size_t ws_bytes, int stage_ds, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* XTAG_B200_H_ */
