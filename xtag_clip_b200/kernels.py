"""Tensor-level wrappers over the C ABI (include/xtag_b200.h).

`CudaKernels` is the only kernel provider the product code uses.  It passes raw device
pointers of torch tensors and the current CUDA stream to libxtag_b200.so; torch is used for
allocation and stream plumbing only.  Every method raises if its inputs are not CUDA tensors:
there is no CPU path.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import XTAG_BF16, XTAG_F32, IMPL_AUTO, check

_DT = {torch.float32: XTAG_F32, torch.bfloat16: XTAG_BF16}


def _dt(t: torch.Tensor) -> int:
    try:
        return _DT[t.dtype]
    except KeyError:
        raise TypeError(f"xtag kernels take float32 or bfloat16 tensors, got {t.dtype}") from None


def _p(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _cuda(*ts: torch.Tensor) -> None:
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("xtag_clip_b200 runs on CUDA (sm_100a) tensors only; there is no CPU fallback "
                               f"(got a tensor on {t.device})")


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


class CudaKernels:
    """One method per C entry point.  Shapes/semantics: see include/xtag_b200.h."""

    name = "cuda"

    def __init__(self, impl: int = IMPL_AUTO):
        self.lib = _lib.load()
        self.impl = impl

    # ---- K3 ----------------------------------------------------------------------------------
    def l2norm_fwd(self, x: torch.Tensor, out_dtype: torch.dtype, eps: float, want_transposed: bool = False,
                   out: Optional[torch.Tensor] = None):
        """out: write the normalised rows into this contiguous [rows, dim] tensor (e.g. an input slot of the loss)."""
        _cuda(x, out)
        x2 = x.reshape(-1, x.shape[-1]).contiguous()
        rows, dim = x2.shape
        if out is not None:
            assert out.is_contiguous() and out.numel() == rows * dim, "out must be a contiguous [rows, dim] tensor"
            out_dtype = out.dtype
            y = out.view(rows, dim)
        else:
            y = torch.empty((rows, dim), dtype=out_dtype, device=x.device)
        yT = torch.empty((dim, rows), dtype=out_dtype, device=x.device) if want_transposed else None
        inv = torch.empty((rows,), dtype=torch.float32, device=x.device)
        if rows:
            check(self.lib.xtag_l2norm_fwd(_p(x2), _dt(x2), _p(y), _DT[out_dtype], _p(yT), _p(inv), rows, dim,
                                           float(eps), _stream()), "xtag_l2norm_fwd")
        return y.reshape(x.shape), inv, yT

    def l2norm_bwd(self, gy: torch.Tensor, y: torch.Tensor, inv: torch.Tensor, gx_dtype: torch.dtype, eps: float):
        _cuda(gy, y, inv)
        y2 = y.reshape(-1, y.shape[-1]).contiguous()
        g2 = gy.reshape(-1, y.shape[-1]).to(y2.dtype).contiguous()
        rows, dim = y2.shape
        gx = torch.empty((rows, dim), dtype=gx_dtype, device=y.device)
        if rows:
            check(self.lib.xtag_l2norm_bwd(_p(g2), _dt(g2), _p(y2), _dt(y2), _p(inv), _p(gx), _DT[gx_dtype], rows, dim,
                                           float(eps), _stream()), "xtag_l2norm_bwd")
        return gx.reshape(y.shape)

    # ---- K1 ----------------------------------------------------------------------------------
    def clip_fwd(self, A: torch.Tensor, Bm: torch.Tensor, scale: torch.Tensor, label_offset: int,
                 col_out: Optional[torch.Tensor] = None, diag_out: Optional[torch.Tensor] = None
                 ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """-> (row_lse [M], col_lse_partial [N], diag [M]), fp32, natural-log units.  label_offset = -1: the column
        block holds no labels (diag untouched).  col_out / diag_out: write into these (contiguous) buffers."""
        _cuda(A, Bm, scale)
        assert A.dtype == Bm.dtype and A.dim() == 2 and Bm.dim() == 2 and A.shape[1] == Bm.shape[1]
        A, Bm = A.contiguous(), Bm.contiguous()
        M, D = A.shape
        N = Bm.shape[0]
        dev = A.device
        row_lse = torch.empty(M, dtype=torch.float32, device=dev)
        col_lse = col_out if col_out is not None else torch.empty(N, dtype=torch.float32, device=dev)
        diag = diag_out if diag_out is not None else torch.empty(M, dtype=torch.float32, device=dev)
        assert col_lse.is_contiguous() and col_lse.numel() == N and diag.is_contiguous() and diag.numel() == M
        nbytes = int(self.lib.xtag_clip_fwd_ws_bytes(M, N, D, _dt(A), self.impl))
        if nbytes == 0:
            check(-1, "xtag_clip_fwd_ws_bytes")
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        check(self.lib.xtag_clip_fwd(_p(A), _p(Bm), _dt(A), M, N, D, _p(scale), int(label_offset),
                                     _p(row_lse), _p(col_lse), _p(diag), _p(ws), nbytes, self.impl, _stream()),
              "xtag_clip_fwd")
        return row_lse, col_lse, diag

    # ---- K1, deferred reductions (one GEMM launch per column block, two reductions per step) ---------------
    def supports_fwd_blocks(self, A: torch.Tensor) -> bool:
        return A.is_cuda and A.dtype == torch.bfloat16 and A.shape[1] % 8 == 0 and self.impl in (IMPL_AUTO, _lib.IMPL_TC)

    def clip_fwd_blocks_begin(self, A: torch.Tensor, block_cols, col_out: Optional[torch.Tensor] = None):
        """Plan a forward over column blocks of `block_cols` columns each (in launch order; their position inside
        the [sum(block_cols)] column range is given per launch).  Returns the state `clip_fwd_block` /
        `clip_fwd_blocks_end` take."""
        _cuda(A)
        M = A.shape[0]
        import ctypes
        rp, cp = ctypes.c_int(0), ctypes.c_int(0)
        offs, tot = [], 0
        for n in block_cols:
            check(self.lib.xtag_clip_fwd_block_parts(M, int(n), ctypes.byref(rp), ctypes.byref(cp)),
                  "xtag_clip_fwd_block_parts")
            offs.append(tot)
            tot += rp.value
        Ncols = int(sum(block_cols))
        dev = A.device
        return dict(M=M, N=Ncols, row_off=offs, P=tot, num_m=cp.value,
                    row_part=torch.empty((tot, M), dtype=torch.float32, device=dev),
                    col_part=torch.empty((cp.value, Ncols), dtype=torch.float32, device=dev),
                    diag=torch.empty(M, dtype=torch.float32, device=dev), col_out=col_out, k=0)

    def clip_fwd_block(self, st, A: torch.Tensor, Bm_blk: torch.Tensor, scale: torch.Tensor, label_offset: int,
                       col_lo: int):
        """Launch K1 for the next planned block: Bm_blk holds global columns [col_lo, col_lo + rows(Bm_blk))."""
        _cuda(A, Bm_blk, scale)
        A, Bm_blk = A.contiguous(), Bm_blk.contiguous()
        M, D = A.shape
        n = Bm_blk.shape[0]
        k = st["k"]
        st["k"] = k + 1
        row_part = st["row_part"][st["row_off"][k]:]
        col_part = st["col_part"][:, col_lo:]
        check(self.lib.xtag_clip_fwd_block(_p(A), _p(Bm_blk), _dt(A), M, n, D, _p(scale), int(label_offset),
                                           row_part.data_ptr(), col_part.data_ptr(), st["N"], _p(st["diag"]),
                                           _stream()), "xtag_clip_fwd_block")

    def clip_fwd_blocks_end(self, st):
        """-> (row_lse [M], col_lse_partial [N] (written into col_out when given), diag [M])"""
        dev = st["row_part"].device
        row_lse = torch.empty(st["M"], dtype=torch.float32, device=dev)
        col = st["col_out"] if st["col_out"] is not None else torch.empty(st["N"], dtype=torch.float32, device=dev)
        assert col.is_contiguous() and col.numel() == st["N"]
        check(self.lib.xtag_lse_reduce2_log2(_p(st["row_part"]), st["P"], st["M"], _p(row_lse),
                                             _p(st["col_part"]), st["num_m"], st["N"], _p(col), _stream()),
              "xtag_lse_reduce2_log2")
        return row_lse, col, st["diag"]

    # ---- K1 fused with the exchange: one persistent launch gated by per-block ready flags -----------------------
    def supports_fwd_stream(self, A: torch.Tensor, blk_cols: int, nblk: int) -> bool:
        return self.supports_fwd_blocks(A) and blk_cols % 256 == 0 and 1 <= nblk <= 16

    def clip_fwd_stream(self, A: torch.Tensor, Bm_all: torch.Tensor, scale: torch.Tensor, label_offset: int,
                        order, wait, blk_cols: int, ready_flags: Optional[torch.Tensor],
                        epoch: Optional[torch.Tensor], col_out: Optional[torch.Tensor] = None):
        """-> (row_lse [M], col_lse_partial [N] (in col_out when given), diag [M]).  Bm_all [N, D] is the gather buffer;
        block order[k] (blk_cols rows) is visited k-th, after ready_flags[order[k]] == epoch when wait[k]."""
        import ctypes
        _cuda(A, Bm_all, scale, ready_flags, epoch)
        assert A.is_contiguous() and Bm_all.is_contiguous()
        M, D = A.shape
        N = Bm_all.shape[0]
        nblk = len(order)
        assert nblk * blk_cols == N and len(wait) == nblk
        dev = A.device
        rp, cp = ctypes.c_int(0), ctypes.c_int(0)
        check(self.lib.xtag_clip_fwd_block_parts(M, N, ctypes.byref(rp), ctypes.byref(cp)), "xtag_clip_fwd_block_parts")
        row_part = torch.empty((rp.value, M), dtype=torch.float32, device=dev)
        col_part = torch.empty((cp.value, N), dtype=torch.float32, device=dev)
        diag = torch.empty(M, dtype=torch.float32, device=dev)
        row_lse = torch.empty(M, dtype=torch.float32, device=dev)
        col = col_out if col_out is not None else torch.empty(N, dtype=torch.float32, device=dev)
        assert col.is_contiguous() and col.numel() == N
        o_arr = (ctypes.c_int * nblk)(*[int(x) for x in order])
        w_arr = (ctypes.c_int * nblk)(*[int(bool(x)) for x in wait])
        check(self.lib.xtag_clip_fwd_stream(_p(A), _p(Bm_all), _dt(A), M, N, D, _p(scale), int(label_offset),
                                            o_arr, w_arr, nblk, int(blk_cols), _p(ready_flags), _p(epoch),
                                            _p(row_part), _p(col_part), N, _p(diag), _stream()), "xtag_clip_fwd_stream")
        check(self.lib.xtag_lse_reduce2_log2(_p(row_part), rp.value, M, _p(row_lse), _p(col_part), cp.value, N, _p(col),
                                             _stream()), "xtag_lse_reduce2_log2")
        return row_lse, col, diag

    def lse_reduce_log2(self, parts: torch.Tensor) -> torch.Tensor:
        """parts [P, n] fp32 in the log2 domain -> out[j] = ln sum_p 2^parts[p, j]  (natural log)"""
        _cuda(parts)
        assert parts.dim() == 2 and parts.dtype == torch.float32
        parts = parts.contiguous()
        P, n = parts.shape
        out = torch.empty(n, dtype=torch.float32, device=parts.device)
        check(self.lib.xtag_lse_reduce_log2(_p(parts), P, n, _p(out), _stream()), "xtag_lse_reduce_log2")
        return out

    def lse_combine(self, parts: torch.Tensor) -> torch.Tensor:
        _cuda(parts)
        parts = parts.contiguous()
        W, N = parts.shape
        out = torch.empty(N, dtype=torch.float32, device=parts.device)
        check(self.lib.xtag_lse_combine(_p(parts), W, N, _p(out), _stream()), "xtag_lse_combine")
        return out

    def lse_combine_ptrs(self, ptrs_dev: torch.Tensor, W: int, N: int) -> torch.Tensor:
        """ptrs_dev: int64 CUDA tensor of W peer-mapped fp32 buffer addresses."""
        _cuda(ptrs_dev)
        out = torch.empty(N, dtype=torch.float32, device=ptrs_dev.device)
        check(self.lib.xtag_lse_combine_ptrs(_p(ptrs_dev), W, N, _p(out), _stream()), "xtag_lse_combine_ptrs")
        return out

    def lse_combine_ptrs_loss(self, ptrs_dev: torch.Tensor, W: int, N: int, row_lse: torch.Tensor, diag: torch.Tensor,
                              label_offset: int, epoch: Optional[torch.Tensor] = None,
                              scratch: Optional[torch.Tensor] = None):
        """-> (col_lse [N], loss 0-d): combine of the W peer buffers, this rank's loss and (optionally) the bump of the
        exchange epoch in one launch.  scratch: a persistent zero-initialised uint8 buffer of
        `combine_loss_scratch_bytes()` (one per exchange object); a fresh one is made when omitted."""
        _cuda(ptrs_dev, row_lse, diag, epoch, scratch)
        dev = ptrs_dev.device
        if scratch is None:
            scratch = torch.zeros(self.combine_loss_scratch_bytes(), dtype=torch.uint8, device=dev)
        out = torch.empty(N, dtype=torch.float32, device=dev)
        loss = torch.empty((), dtype=torch.float32, device=dev)
        check(self.lib.xtag_lse_combine_ptrs_loss(_p(ptrs_dev), W, N, _p(out), _p(row_lse), _p(diag), row_lse.numel(),
                                                  int(label_offset), _p(loss), _p(epoch), _p(scratch), _stream()),
              "xtag_lse_combine_ptrs_loss")
        return out, loss

    def combine_loss_scratch_bytes(self) -> int:
        return int(self.lib.xtag_lse_combine_loss_scratch_bytes())

    def sum_ptrs_bf16(self, ptrs_dev: torch.Tensor, W: int, shape) -> torch.Tensor:
        _cuda(ptrs_dev)
        out = torch.empty(shape, dtype=torch.bfloat16, device=ptrs_dev.device)
        check(self.lib.xtag_sum_ptrs_bf16(_p(ptrs_dev), W, out.numel(), _p(out), _stream()), "xtag_sum_ptrs_bf16")
        return out

    def clip_loss(self, row_lse: torch.Tensor, diag: torch.Tensor, col_lse: torch.Tensor, label_offset: int):
        _cuda(row_lse, diag, col_lse)
        out = torch.empty((), dtype=torch.float32, device=row_lse.device)
        check(self.lib.xtag_clip_loss(_p(row_lse), _p(diag), _p(col_lse), row_lse.numel(), int(label_offset),
                                      _p(out), _stream()), "xtag_clip_loss")
        return out

    # ---- K2 ----------------------------------------------------------------------------------
    def clip_bwd(self, A, Bm, scale, label_offset, row_lse, col_lse, w_row, w_col, w_diag, grad_out,
                 need_dA: bool, need_dB: bool, grad_dtype: torch.dtype, ws: Optional[torch.Tensor] = None,
                 reuse_ds: bool = False, return_ws: bool = False, dB_out: Optional[torch.Tensor] = None):
        """-> (dA [M,D] | None, dB [N,D] | None, dscale 0-d fp32[, ws]).  reuse_ds: `ws` is the workspace of a
        preceding call with the same operands/weights, its staged dS is reused (dscale is then 0)."""
        _cuda(A, Bm, scale, row_lse, col_lse, grad_out)
        A, Bm = A.contiguous(), Bm.contiguous()
        M, D = A.shape
        N = Bm.shape[0]
        dev = A.device
        dA = torch.empty((M, D), dtype=grad_dtype, device=dev) if need_dA else None
        dB = None
        if need_dB:
            dB = dB_out if dB_out is not None else torch.empty((N, D), dtype=grad_dtype, device=dev)
            assert dB.is_contiguous() and dB.shape == (N, D) and dB.dtype == grad_dtype
        # (a reuse_ds call never writes dscale: no fill kernel for it)
        dscale = None if reuse_ds else torch.zeros((), dtype=torch.float32, device=dev)
        g = grad_out.detach().to(torch.float32).reshape(1).contiguous()
        nbytes = int(self.lib.xtag_clip_bwd_ws_bytes(M, N, D, _dt(A), self.impl))
        if nbytes == 0:
            check(-1, "xtag_clip_bwd_ws_bytes")
        if ws is None:
            assert not reuse_ds
            ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        assert ws.numel() >= nbytes
        check(self.lib.xtag_clip_bwd(_p(A), _p(Bm), _dt(A), M, N, D, _p(scale), int(label_offset),
                                     _p(row_lse), _p(col_lse), float(w_row), float(w_col), float(w_diag), _p(g),
                                     _p(dA), _p(dB), _DT[grad_dtype], _p(dscale), _p(ws), ws.numel(), self.impl,
                                     _lib.BWD_REUSE_DS if reuse_ds else 0, _stream()), "xtag_clip_bwd")
        if return_ws:
            return dA, dB, dscale, ws
        return dA, dB, dscale

    # ---- sigmoid (SigLIP) loss on the same mainloop ------------------------------------------------------
    def siglip_fwd(self, A: torch.Tensor, Bm: torch.Tensor, scale: torch.Tensor, bias: torch.Tensor, label_offset: int,
                   weight: float, stage_ds: bool = True):
        """-> (out3 fp32 [3] = loss, d loss/d scale, d loss/d bias for a unit upstream gradient; ws | None).  ws holds
        the staged logit gradient for `clip_bwd(..., ws=ws, reuse_ds=True)`."""
        _cuda(A, Bm, scale, bias)
        assert A.dtype == torch.bfloat16 and Bm.dtype == torch.bfloat16 and A.shape[1] == Bm.shape[1]
        A, Bm = A.contiguous(), Bm.contiguous()
        M, D = A.shape
        N = Bm.shape[0]
        nbytes = int(self.lib.xtag_clip_bwd_ws_bytes(M, N, D, XTAG_BF16, _lib.IMPL_TC))
        if nbytes == 0:
            check(-1, "xtag_clip_bwd_ws_bytes")
        ws = torch.empty(nbytes, dtype=torch.uint8, device=A.device)
        out3 = torch.empty(3, dtype=torch.float32, device=A.device)
        check(self.lib.xtag_siglip_fwd(_p(A), _p(Bm), XTAG_BF16, M, N, D, _p(scale), _p(bias), int(label_offset),
                                       float(weight), _p(out3), _p(ws), nbytes, int(bool(stage_ds)), _stream()),
              "xtag_siglip_fwd")
        return out3, (ws if stage_ds else None)

    def tc_gemm_nt(self, A: torch.Tensor, B: torch.Tensor, out_dtype=torch.float32, alpha: float = 1.0):
        _cuda(A, B)
        A, B = A.contiguous(), B.contiguous()
        M, K = A.shape
        N = B.shape[0]
        C = torch.empty((M, N), dtype=out_dtype, device=A.device)
        check(self.lib.xtag_tc_gemm_nt(_p(A), _p(B), _p(C), _DT[out_dtype], M, N, K, float(alpha), _stream()),
              "xtag_tc_gemm_nt")
        return C

    def tc_linear(self, x2d: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor]) -> torch.Tensor:
        """y [M, N] bf16 = x2d [M, K] bf16 @ weight [N, K]^T bf16 + bias [N] fp32 on the tcgen05 dense-layer kernel."""
        _cuda(x2d, weight, bias)
        assert x2d.dtype == torch.bfloat16 and weight.dtype == torch.bfloat16 and x2d.dim() == 2 and weight.dim() == 2
        x2d, weight = x2d.contiguous(), weight.contiguous()
        M, Kd = x2d.shape
        N = weight.shape[0]
        assert weight.shape[1] == Kd and N % 8 == 0
        if bias is not None:
            bias = bias.detach().to(torch.float32).contiguous()
        y = torch.empty((M, N), dtype=torch.bfloat16, device=x2d.device)
        check(self.lib.xtag_tc_linear_bf16(_p(x2d), Kd, _p(weight), _p(bias), _p(y), N, M, N, Kd, _stream()),
              "xtag_tc_linear_bf16")
        return y

    def tc_gemm(self, A: torch.Tensor, B: torch.Tensor, a_mn: bool, b_mn: bool, out_dtype=torch.float32,
                alpha: float = 1.0):
        """C[M,N] = alpha * sum_k A(m,k) B(n,k); A is [M,K] (a_mn False) or [K,M] (True), B is [N,K] or [K,N]."""
        _cuda(A, B)
        A, B = A.contiguous(), B.contiguous()
        (K, M) = A.shape if a_mn else A.shape[::-1]
        N = B.shape[1] if b_mn else B.shape[0]
        C = torch.empty((M, N), dtype=out_dtype, device=A.device)
        nbytes = int(self.lib.xtag_tc_gemm_ws_bytes(M, N, K))          # split-K slabs (few output tiles, long K)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=A.device) if nbytes else None
        check(self.lib.xtag_tc_gemm_ex(_p(A), _p(B), _p(C), _DT[out_dtype], M, N, K, float(alpha), int(a_mn), int(b_mn),
                                       _p(ws), nbytes, _stream()), "xtag_tc_gemm_ex")
        return C

    # ---- K4 ----------------------------------------------------------------------------------
    def xattn_fwd(self, q, k, v, heads: int, softmax_scale: float, dropout_p: float, seed: int, offset: int):
        """q [b,Lq,H], k/v [b,Lk,H] (last dim contiguous; k and v may be strided views of one buffer)
        -> (ctx [b,Lq,H], lse [b,heads,Lq])"""
        _cuda(q, k, v)
        b, Lq, H = q.shape
        Lk = k.shape[1]
        dh = H // heads
        for t in (q, k, v):
            assert t.stride(-1) == 1 and t.stride(0) == t.shape[1] * t.stride(1), "rows must be uniformly strided"
        o = torch.empty((b, Lq, H), dtype=q.dtype, device=q.device)
        lse = torch.empty((b, heads, Lq), dtype=torch.float32, device=q.device)
        check(self.lib.xtag_xattn_fwd(_p(q), _p(k), _p(v), _dt(q), _p(o), _p(lse), b, Lq, Lk, heads, dh,
                                      q.stride(1), k.stride(1), v.stride(1), float(softmax_scale), float(dropout_p),
                                      int(seed), int(offset), _stream()), "xtag_xattn_fwd")
        return o, lse

    def xattn_bwd(self, q, k, v, o, do, lse, heads: int, softmax_scale: float, dropout_p: float, seed: int,
                  offset: int, dk_out: Optional[torch.Tensor] = None, dv_out: Optional[torch.Tensor] = None):
        """-> (dq, dk, dv).  dk_out / dv_out: write dK / dV into these [b, Lk, H] views (last dim contiguous, uniform
        row stride -- e.g. column slices of one fused K|V gradient buffer); bf16 tensor-core path only, otherwise the
        gradients are computed contiguously and copied in."""
        _cuda(q, k, v, o, do, lse, dk_out, dv_out)
        b, Lq, H = q.shape
        Lk = k.shape[1]
        dh = H // heads
        do = do.contiguous()
        dq = torch.empty((b, Lq, H), dtype=q.dtype, device=q.device)
        ws = torch.empty(b * heads * Lq, dtype=torch.float32, device=q.device)
        strided = dk_out is not None and dv_out is not None
        if strided:
            for t in (dk_out, dv_out):
                assert t.shape == (b, Lk, H) and t.dtype == q.dtype and t.stride(-1) == 1 and \
                    t.stride(0) == Lk * t.stride(1), "dk_out / dv_out: [b, Lk, H] views with uniformly strided rows"
            rc = self.lib.xtag_xattn_bwd_ld(_p(q), _p(k), _p(v), _p(o), _p(do), _p(lse), _dt(q), _p(dq), _p(dk_out),
                                            _p(dv_out), b, Lq, Lk, heads, dh, q.stride(1), k.stride(1), v.stride(1),
                                            dk_out.stride(1), dv_out.stride(1), float(softmax_scale), float(dropout_p),
                                            int(seed), int(offset), _p(ws), ws.numel() * 4, _stream())
            if rc == 0:
                return dq, dk_out, dv_out
            if rc != _lib.ERR_UNSUPPORTED:
                check(rc, "xtag_xattn_bwd_ld")
        dk = torch.empty((b, Lk, H), dtype=q.dtype, device=q.device)
        dv = torch.empty((b, Lk, H), dtype=q.dtype, device=q.device)
        check(self.lib.xtag_xattn_bwd(_p(q), _p(k), _p(v), _p(o), _p(do), _p(lse), _dt(q), _p(dq), _p(dk), _p(dv),
                                      b, Lq, Lk, heads, dh, q.stride(1), k.stride(1), v.stride(1),
                                      float(softmax_scale), float(dropout_p), int(seed), int(offset),
                                      _p(ws), ws.numel() * 4, _stream()), "xtag_xattn_bwd")
        if strided:
            dk_out.copy_(dk)
            dv_out.copy_(dv)
            return dq, dk_out, dv_out
        return dq, dk, dv

    # ---- symmetric CE on a materialised square matrix (DQNCOSLoss) ---------------------------------------
    def symm_ce_fwd(self, x: torch.Tensor):
        """x [n, n] (fp32 / bf16, last dim contiguous) -> (loss 0-d fp32, row_lse [n], col_lse [n]); X is read once."""
        _cuda(x)
        assert x.dim() == 2 and x.shape[0] == x.shape[1] and x.stride(1) == 1
        n = x.shape[0]
        dev = x.device
        row, col, diag = (torch.empty(n, dtype=torch.float32, device=dev) for _ in range(3))
        loss = torch.empty((), dtype=torch.float32, device=dev)
        nbytes = int(self.lib.xtag_symm_ce_ws_bytes(n))
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        check(self.lib.xtag_symm_ce_fwd(_p(x), _dt(x), n, x.stride(0), _p(row), _p(col), _p(diag), _p(loss), _p(ws), nbytes,
                                        _stream()), "xtag_symm_ce_fwd")
        return loss, row, col

    def symm_ce_bwd(self, x: torch.Tensor, row_lse: torch.Tensor, col_lse: torch.Tensor, grad_out: torch.Tensor):
        _cuda(x, row_lse, col_lse, grad_out)
        n = x.shape[0]
        g = grad_out.detach().to(torch.float32).reshape(1).contiguous()
        dx = torch.empty((n, n), dtype=x.dtype, device=x.device)
        check(self.lib.xtag_symm_ce_bwd(_p(x), _dt(x), n, x.stride(0), _p(row_lse), _p(col_lse), _p(g), _p(dx), n,
                                        _stream()), "xtag_symm_ce_bwd")
        return dx

    # ---- K6: dropout + residual + LayerNorm ------------------------------------------------------------
    def supports_ln_res(self, H: int) -> bool:
        return H % 256 == 0 and 256 <= H <= 1024

    def ln_res_fwd(self, x: torch.Tensor, resid: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float,
                   dropout_p: float, seed: int, offset: int):
        """x [rows, H] bf16, resid [resid_rows, H] (bf16 / fp32, rows % resid_rows == 0) -> (y bf16, z bf16, mean, rstd)"""
        _cuda(x, resid, gamma, beta)
        assert x.dtype == torch.bfloat16 and x.dim() == 2 and resid.dim() == 2 and x.shape[1] == resid.shape[1]
        x, resid = x.contiguous(), resid.contiguous()
        rows, H = x.shape
        assert rows % resid.shape[0] == 0
        dev = x.device
        z = torch.empty_like(x)
        y = torch.empty_like(x)
        mean = torch.empty(rows, dtype=torch.float32, device=dev)
        rstd = torch.empty(rows, dtype=torch.float32, device=dev)
        g32, b32 = gamma.detach().float().contiguous(), beta.detach().float().contiguous()
        check(self.lib.xtag_ln_res_fwd(_p(x), _p(resid), _dt(resid), resid.shape[0], _p(g32), _p(b32), _p(z), _p(y),
                                       _p(mean), _p(rstd), rows, H, float(eps), float(dropout_p), int(seed), int(offset),
                                       _stream()), "xtag_ln_res_fwd")
        return y, z, mean, rstd

    def ln_res_bwd(self, dy: torch.Tensor, z: torch.Tensor, mean: torch.Tensor, rstd: torch.Tensor, gamma: torch.Tensor,
                   dropout_p: float, seed: int, offset: int):
        """-> (dx bf16 [rows, H], dresid bf16 [rows, H], dgamma fp32 [H], dbeta fp32 [H])"""
        _cuda(dy, z, mean, rstd, gamma)
        rows, H = z.shape
        dy = dy.reshape(rows, H)
        if dy.dtype not in _DT:
            dy = dy.float()
        dy = dy.contiguous()
        dev = z.device
        dx = torch.empty_like(z)
        dres = torch.empty_like(z)
        dgb = torch.empty((2, H), dtype=torch.float32, device=dev)
        g32 = gamma.detach().float().contiguous()
        nbytes = int(self.lib.xtag_ln_res_bwd_ws_bytes(rows, H))
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        check(self.lib.xtag_ln_res_bwd(_p(dy), _dt(dy), _p(z), _p(mean), _p(rstd), _p(g32), _p(dx), _p(dres),
                                       dgb[0].data_ptr(), dgb[1].data_ptr(), rows, H, float(dropout_p), int(seed),
                                       int(offset), _p(ws), nbytes, _stream()), "xtag_ln_res_bwd")
        return dx, dres, dgb[0], dgb[1]

    # ---- K5 ----------------------------------------------------------------------------------
    def asl(self, x: torch.Tensor, y: torch.Tensor, gamma_neg, gamma_pos, clip, eps, want_dx: bool,
            want_idx: bool):
        """-> (loss 0-d fp32, dx fp32 | None, idx6 int32 [rows,6] | None)"""
        _cuda(x, y)
        x2 = x.reshape(-1, x.shape[-1]).contiguous()
        y2 = y.reshape(-1, x.shape[-1]).to(torch.float32).contiguous()
        rows, cols = x2.shape
        loss = torch.empty((), dtype=torch.float32, device=x.device)
        dx = torch.empty((rows, cols), dtype=torch.float32, device=x.device) if want_dx else None
        idx = torch.empty((rows, 6), dtype=torch.int32, device=x.device) if want_idx else None
        check(self.lib.xtag_asl_fwd(_p(x2), _dt(x2), _p(y2), rows, cols, float(gamma_neg), float(gamma_pos),
                                    float(clip or 0.0), float(eps), _p(loss), _p(dx), _p(idx), _stream()),
              "xtag_asl_fwd")
        return loss, dx, idx


_default: Optional[CudaKernels] = None


def default_kernels() -> CudaKernels:
    global _default
    if _default is None:
        _default = CudaKernels()
    return _default
