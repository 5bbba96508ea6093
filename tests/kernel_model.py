"""Test double for `xtag_clip_b200.kernels.CudaKernels` (TEST INFRASTRUCTURE, lives in tests/ only).

A plain torch fp64 statement of each C entry point's *contract* (include/xtag_b200.h).  Two uses:
  * CPU tests of the host logic (which collectives, label offsets, gradient weights, slicing) inject it
    through the `_kernels=` test hook, so world_size-2 gloo tests run without a GPU;
  * GPU tests compare individual kernel outputs against it.
It is never importable from the package and never used as a fallback.
"""
from __future__ import annotations

import math

import torch


def _f64(t):
    return t.detach().to(torch.float64)


class ModelKernels:
    name = "model"

    def __init__(self, fwd_blocks: bool = True):
        self.calls = []
        self.fwd_blocks = fwd_blocks       # False: exercise the per-block clip_fwd + lse_combine host path

    # K3
    def l2norm_fwd(self, x, out_dtype, eps, want_transposed=False):
        x64 = _f64(x)
        n = torch.linalg.vector_norm(x64, dim=-1, keepdim=True).clamp_min(eps)
        y = (x64 / n).to(out_dtype)
        inv = (1.0 / n).reshape(-1).to(torch.float32)
        yT = y.reshape(-1, y.shape[-1]).T.contiguous() if want_transposed else None
        return y, inv, yT

    def l2norm_bwd(self, gy, y, inv, gx_dtype, eps):
        g, yy, r = _f64(gy), _f64(y), _f64(inv).reshape(*y.shape[:-1], 1)
        # the ABI stores inv_norm in fp32: compare in fp32 like the kernel does
        clamped = inv.reshape(*y.shape[:-1], 1) >= torch.tensor(1.0 / eps, dtype=torch.float32)
        dot = (g * yy).sum(-1, keepdim=True)
        gx = torch.where(clamped, g * r, r * (g - yy * dot))
        return gx.to(gx_dtype)

    # K1
    def clip_fwd(self, A, Bm, scale, label_offset, col_out=None, diag_out=None):
        self.calls.append(("clip_fwd", tuple(A.shape), tuple(Bm.shape), int(label_offset)))
        S = float(scale.reshape(-1)[0]) * (_f64(A) @ _f64(Bm).T)
        M = A.shape[0]
        col = torch.logsumexp(S, 0).float()
        if col_out is not None:
            col_out.copy_(col)
            col = col_out
        diag = diag_out if diag_out is not None else torch.zeros(M, dtype=torch.float32)
        if label_offset >= 0:                      # -1: no labels in this column block, diag untouched
            diag.copy_(S[torch.arange(M), torch.arange(M) + label_offset].float())
        return torch.logsumexp(S, 1).float(), col, diag

    # K1 with deferred reductions: same contract, expressed with per-block natural-log LSEs
    def supports_fwd_blocks(self, A):
        return self.fwd_blocks

    def clip_fwd_blocks_begin(self, A, block_cols, col_out=None):
        M, N = A.shape[0], int(sum(block_cols))
        return dict(M=M, N=N, rows=[], col=torch.full((N,), float("nan"), dtype=torch.float32),
                    diag=torch.zeros(M, dtype=torch.float32), col_out=col_out, planned=list(block_cols), k=0)

    def clip_fwd_block(self, st, A, Bm_blk, scale, label_offset, col_lo):
        assert Bm_blk.shape[0] == st["planned"][st["k"]], "blocks must be launched in the planned order"
        st["k"] += 1
        row, col, _ = self.clip_fwd(A, Bm_blk, scale, label_offset, diag_out=st["diag"])
        st["rows"].append(row)
        st["col"][col_lo:col_lo + Bm_blk.shape[0]] = col

    def clip_fwd_blocks_end(self, st):
        assert st["k"] == len(st["planned"]) and not torch.isnan(st["col"]).any(), "a column block was never launched"
        row = torch.logsumexp(torch.stack(st["rows"]).double(), 0).float()
        col = st["col"]
        if st["col_out"] is not None:
            st["col_out"].copy_(col)
            col = st["col_out"]
        return row, col, st["diag"]

    # K1 fused with the exchange: the contract is the plain forward over the whole gather buffer; the model checks
    # the schedule arguments the host passes (own block first and not waited for, a permutation, one flag per block)
    def supports_fwd_stream(self, A, blk_cols, nblk):
        return self.fwd_blocks

    def clip_fwd_stream(self, A, Bm_all, scale, label_offset, order, wait, blk_cols, ready_flags, epoch, col_out=None):
        nblk = len(order)
        assert sorted(order) == list(range(nblk)) and nblk * blk_cols == Bm_all.shape[0]
        assert list(wait) == [False] + [True] * (nblk - 1), "only the rank's own block may be consumed without its flag"
        assert label_offset == order[0] * blk_cols, "the first block visited must be the one holding the labels"
        assert ready_flags.numel() == nblk and epoch.numel() == 1
        self.calls.append(("clip_fwd_stream", tuple(A.shape), tuple(Bm_all.shape), int(label_offset), tuple(order)))
        return self.clip_fwd(A, Bm_all, scale, label_offset, col_out=col_out)

    def lse_combine_ptrs(self, ptrs, W, N):
        return torch.logsumexp(torch.stack([_f64(p) for p in ptrs]), 0).float()

    def sum_ptrs_bf16(self, ptrs, W, shape):
        return sum(_f64(p) for p in ptrs).reshape(shape).to(torch.bfloat16)

    def lse_reduce_log2(self, parts):
        return (torch.logsumexp(_f64(parts) * math.log(2.0), 0)).float()

    def lse_combine(self, parts):
        return torch.logsumexp(_f64(parts), 0).float()

    def clip_loss(self, row_lse, diag, col_lse, label_offset):
        M = row_lse.numel()
        c = _f64(col_lse)[label_offset:label_offset + M]
        return (0.5 * ((_f64(row_lse) - _f64(diag)).mean() + (c - _f64(diag)).mean())).float()

    # K2
    def clip_bwd(self, A, Bm, scale, label_offset, row_lse, col_lse, w_row, w_col, w_diag, grad_out,
                 need_dA, need_dB, grad_dtype, ws=None, reuse_ds=False, return_ws=False, dB_out=None):
        self.calls.append(("clip_bwd", tuple(A.shape), tuple(Bm.shape), int(label_offset), w_row, w_col, w_diag))
        s = float(scale.reshape(-1)[0])
        g = float(grad_out.reshape(-1)[0])
        A64, B64 = _f64(A), _f64(Bm)
        if reuse_ds and torch.is_tensor(ws):
            # XTAG_BWD_REUSE_DS on a logit gradient staged by another producer (siglip_fwd): two GEMMs, alpha = scale
            dS = _f64(ws)
            dA = (s * dS @ B64).to(grad_dtype) if need_dA else None
            dB = (s * dS.T @ A64).to(grad_dtype) if need_dB else None
            return dA, dB, torch.zeros((), dtype=torch.float32)
        raw = A64 @ B64.T
        S = s * raw
        M, N = S.shape
        dS = w_row * torch.exp(S - _f64(row_lse)[:, None])
        if w_col != 0.0:
            dS = dS + w_col * torch.exp(S - _f64(col_lse)[None, :])
        eye = torch.zeros(M, N, dtype=torch.float64)
        eye[torch.arange(M), torch.arange(M) + label_offset] = 1.0
        dS = g * (dS - w_diag * eye)
        dA = (s * dS @ B64).to(grad_dtype) if need_dA else None
        dB = (s * dS.T @ A64).to(grad_dtype) if need_dB else None
        if dB is not None and dB_out is not None:
            dB_out.copy_(dB)
            dB = dB_out
        ds = torch.zeros((), dtype=torch.float32) if reuse_ds else (dS * raw).sum().float()
        if return_ws:
            return dA, dB, ds, "ws"
        return dA, dB, ds

    # sigmoid loss on the contrastive mainloop (xtag_siglip_fwd): out3 = (loss, dloss/dscale, dloss/dbias) for a unit
    # upstream gradient; "ws" = the staged logit gradient
    def siglip_fwd(self, A, Bm, scale, bias, label_offset, weight, stage_ds=True):
        self.calls.append(("siglip_fwd", tuple(A.shape), tuple(Bm.shape), int(label_offset)))
        s, b0 = float(scale.reshape(-1)[0]), float(bias.reshape(-1)[0])
        raw = _f64(A) @ _f64(Bm).T
        z = s * raw + b0
        M, N = z.shape
        lab = -torch.ones(M, N, dtype=torch.float64)
        if label_offset >= 0:
            lab[torch.arange(M), torch.arange(M) + label_offset] = 1.0
        t = -lab * z
        loss = weight * torch.nn.functional.softplus(t).sum()
        dS = -lab * weight * torch.sigmoid(t)
        out3 = torch.stack([loss, (dS * raw).sum(), dS.sum()]).float()
        return out3, (dS if stage_ds else None)

    # K4
    @staticmethod
    def _attn(q, k, v, heads, scale):
        b, Lq, H = q.shape
        Lk, dh = k.shape[1], H // heads
        qh = q.reshape(b, Lq, heads, dh).permute(0, 2, 1, 3)
        kh = k.reshape(b, Lk, heads, dh).permute(0, 2, 1, 3)
        vh = v.reshape(b, Lk, heads, dh).permute(0, 2, 1, 3)
        s = qh @ kh.transpose(-1, -2) * scale
        lse = torch.logsumexp(s, -1)
        o = (torch.softmax(s, -1) @ vh).permute(0, 2, 1, 3).reshape(b, Lq, H)
        return o, lse

    def xattn_fwd(self, q, k, v, heads, softmax_scale, dropout_p, seed, offset):
        assert dropout_p == 0.0, "the contract model covers eval mode"
        self.calls.append(("xattn_fwd", tuple(q.shape), tuple(k.shape), heads))
        o, lse = self._attn(_f64(q), _f64(k), _f64(v), heads, softmax_scale)
        return o.to(q.dtype), lse.float()

    def xattn_bwd(self, q, k, v, o, do, lse, heads, softmax_scale, dropout_p, seed, offset, dk_out=None, dv_out=None):
        if dk_out is not None:
            return self.xattn_bwd_into(q, k, v, o, do, lse, heads, softmax_scale, dropout_p, seed, offset, dk_out, dv_out)
        assert dropout_p == 0.0
        q64, k64, v64 = (_f64(t).requires_grad_(True) for t in (q, k, v))
        with torch.enable_grad():
            out, _ = self._attn(q64, k64, v64, heads, softmax_scale)
            out.backward(_f64(do))
        return q64.grad.to(q.dtype), k64.grad.to(q.dtype), v64.grad.to(q.dtype)

    def xattn_bwd_into(self, q, k, v, o, do, lse, heads, softmax_scale, dropout_p, seed, offset, dk_out, dv_out):
        dq, dk, dv = self.xattn_bwd(q, k, v, o, do, lse, heads, softmax_scale, dropout_p, seed, offset)
        dk_out.copy_(dk)
        dv_out.copy_(dv)
        return dq, dk_out, dv_out

    # dense layer with bias epilogue (xtag_tc_linear_bf16) and the plain GEMM (xtag_tc_gemm_ex)
    def tc_linear(self, x2d, weight, bias):
        self.calls.append(("tc_linear", tuple(x2d.shape), tuple(weight.shape)))
        y = _f64(x2d) @ _f64(weight).T
        if bias is not None:
            y = y + _f64(bias)
        return y.to(torch.bfloat16)

    def tc_gemm(self, A, B, a_mn, b_mn, out_dtype=torch.float32, alpha=1.0):
        self.calls.append(("tc_gemm", tuple(A.shape), tuple(B.shape), bool(a_mn), bool(b_mn)))
        A64 = _f64(A).T if a_mn else _f64(A)          # -> [M, K]
        B64 = _f64(B).T if b_mn else _f64(B)          # -> [N, K]
        return (alpha * (A64 @ B64.T)).to(out_dtype)

    # K6: y = LayerNorm(dropout(x) + resid) (eval mode in the contract model)
    def supports_ln_res(self, H):
        return True

    def ln_res_fwd(self, x, resid, gamma, beta, eps, dropout_p, seed, offset):
        assert dropout_p == 0.0, "the contract model covers eval mode"
        self.calls.append(("ln_res_fwd", tuple(x.shape), tuple(resid.shape)))
        rows = x.shape[0]
        z = _f64(x) + _f64(resid).repeat(rows // resid.shape[0], 1)
        mean = z.mean(-1)
        rstd = 1.0 / torch.sqrt(z.var(-1, unbiased=False) + eps)
        y = (z - mean[:, None]) * rstd[:, None] * _f64(gamma) + _f64(beta)
        return y.to(torch.bfloat16), z.to(torch.bfloat16), mean.float(), rstd.float()

    def ln_res_bwd(self, dy, z, mean, rstd, gamma, dropout_p, seed, offset):
        assert dropout_p == 0.0
        self.calls.append(("ln_res_bwd", tuple(z.shape)))
        H = z.shape[1]
        d = _f64(dy).reshape(-1, H)
        xh = (_f64(z) - _f64(mean)[:, None]) * _f64(rstd)[:, None]
        dg, db = (d * xh).sum(0), d.sum(0)
        dyg = d * _f64(gamma)
        dz = _f64(rstd)[:, None] * (dyg - dyg.mean(-1, keepdim=True) - xh * (dyg * xh).mean(-1, keepdim=True))
        dzb = dz.to(torch.bfloat16)
        return dzb, dzb.clone(), dg.float(), db.float()

    # K5
    def asl(self, x, y, gamma_neg, gamma_pos, clip, eps, want_dx, want_idx):
        x64 = _f64(x).reshape(-1, x.shape[-1]).requires_grad_(True)
        y64 = _f64(y).reshape(-1, x.shape[-1])
        with torch.enable_grad():
            p = torch.sigmoid(x64)
            pn = 1 - p
            if clip and clip > 0:
                pn = (pn + clip).clamp(max=1)
            loss = y64 * torch.log(p.clamp(min=eps)) + (1 - y64) * torch.log(pn.clamp(min=eps))
            if gamma_neg > 0 or gamma_pos > 0:
                with torch.no_grad():
                    pt = p * y64 + pn * (1 - y64)
                    w = torch.pow(1 - pt, gamma_pos * y64 + gamma_neg * (1 - y64))
                loss = loss * w
            total = -loss.sum()
            total.backward()
        idx = None
        if want_idx:
            s = torch.sigmoid(x64.detach())
            cols, pos = [], 0
            for size in (3, 4, 3, 4, 4, 4):
                sc = s[:, pos:pos + size] + s[:, 22 + pos:22 + pos + size]
                cols.append(sc.argmax(-1, keepdim=True) + pos)
                pos += size
            idx = torch.cat(cols, -1).to(torch.int32)
        return total.detach().float(), (x64.grad.float() if want_dx else None), idx
