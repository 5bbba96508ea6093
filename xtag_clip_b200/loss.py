"""Drop-in for the reference's contrastive head: ``ClipLoss``, ``gather_features``, ``create_loss``.

Mirrors /root/reference/src/open_clip/loss.py:21-139 and factory.py:433-469 (same names, argument
meaning, return values and error behaviour) with the arithmetic replaced by the fused sm_100a
kernels of libxtag_b200.so (K1 forward, K2 backward).  The B x B logit matrix is never
materialised; the only collectives are one feature all-gather, one all-gather of the [B] partial
column log-sum-exps, and (gather_with_grad) one reduce-scatter of the text-feature gradient.

Gradient modes, rank r owns rows/cols R_r = [r*b, (r+1)*b):
  world_size == 1                     dS = (P_row + P_col - 2*1)/(2B) on the full matrix
  local_loss, gather_with_grad        rank computes dS[R_r, :] with weights 1/(2b); dI_r is complete
                                      locally, dT partial [B, D] is reduce-scattered (SUM) -- exactly
                                      the reference's W x (single-process gradient)
  local_loss, not gather_with_grad    only the direct local operands carry gradient: dI_r from
                                      (P_row - 1)[R_r, :], dT_r from (P_col - 1)[:, R_r] (second call
                                      with the operands swapped)
  not local_loss                      the full B x B problem on every rank (replicated, as in the
                                      reference); gather_with_grad reduce-scatters, otherwise the
                                      local block is sliced out

Documented deviation: with local_loss + gather_with_grad the per-rank ``logit_scale.grad`` is this
rank's *row block* share  sum_{i in R_r, j} dS_ij S_ij / s, whereas the reference's per-rank value
mixes the row term of R_r with the column term of R_r.  Their sum over ranks -- what DDP's gradient
all-reduce produces for the shared ``logit_scale`` parameter -- is identical.
The upstream gradient of the loss is assumed equal on all ranks (it is for ``loss.backward()`` and
for GradScaler-scaled losses).
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Optional

import torch
import torch.nn as nn
from torch.nn import functional as F

try:
    import torch.distributed.nn
    from torch import distributed as dist

    has_distributed = True
except ImportError:  # pragma: no cover
    has_distributed = False
    dist = None


# ------------------------------------------------------------------------------------------------
# plumbing: collectives on the default (or given) process group
# ------------------------------------------------------------------------------------------------
class _Comm:
    def __init__(self, world_size: int, rank: int, group=None):
        self.world_size, self.rank, self.group = world_size, rank, group
        self._side = None
        self._symm = {}

    def symm_exchange(self, x: torch.Tensor):
        """Peer-memory exchange object for [b, D] bf16 features (None when symmetric memory is unusable: CPU
        tensors, other dtypes, or a failed rendezvous -- the NCCL P2P pipeline is used then)."""
        from . import symm
        if x.dtype != torch.bfloat16 or not symm.available(x.device):
            return None
        key = (tuple(x.shape), x.dtype)
        if key not in self._symm:
            try:
                self._symm[key] = symm.SymmExchange(self.world_size, self.rank, self.group, x.device, x.shape[0],
                                                    x.shape[1], x.dtype, torch.bfloat16)
            except Exception as e:  # pragma: no cover  (reported once; the NCCL path still works)
                import warnings
                warnings.warn(f"xtag_clip_b200: symmetric-memory exchange unavailable ({e}); using NCCL P2P")
                self._symm[key] = None
        return self._symm[key]

    # -- stream helpers (no-ops for CPU tensors, so the same schedule runs under gloo in the CPU tests) --------
    def side_stream(self, device):
        if device.type != "cuda":
            return None
        if self._side is None:
            self._side = torch.cuda.Stream(device=device)
        return self._side

    def chunk_groups(self):
        """Ranks are exchanged in groups of G consecutive ranks (G = 2 from 4 ranks on: 2b columns per kernel
        launch keep the 148 SMs in whole waves); returns (G, number of groups, own group)."""
        W = self.world_size
        G = 2 if (W >= 4 and W % 2 == 0) else 1
        return G, W // G, self.rank // G

    def pipelined_gather(self, x: torch.Tensor, out_all: torch.Tensor):
        """Chunk-pipelined all-gather of [b, D] blocks into out_all [W*b, D] (rank order).  Round 0 exchanges
        inside the rank's own group, round j sends to group g+j and receives group g-j (batched NCCL P2P on a side
        stream).  Returns [(row_lo, row_hi, event | None)] in arrival order: the caller consumes block k while
        block k+1 is still in flight.  Replaces the blocking torch.distributed.nn.all_gather of loss.py:51-52."""
        W, r = self.world_size, self.rank
        b = x.shape[0]
        G, NG, g0 = self.chunk_groups()
        side = self.side_stream(x.device)
        out_all[r * b:(r + 1) * b].copy_(x)
        plan = []
        if side is not None:
            side.wait_stream(torch.cuda.current_stream())
        ctx = torch.cuda.stream(side) if side is not None else _NullCtx()
        with ctx:
            for j in range(NG):
                send_grp, recv_grp = (g0 + j) % NG, (g0 - j) % NG
                ops = []
                for m in range(G):
                    ps, pr = send_grp * G + m, recv_grp * G + m
                    if ps != r:
                        ops.append(dist.P2POp(dist.isend, x, ps, group=self.group))
                    if pr != r:
                        ops.append(dist.P2POp(dist.irecv, out_all[pr * b:(pr + 1) * b], pr, group=self.group))
                if ops:
                    for req in dist.batch_isend_irecv(ops):
                        req.wait()
                ev = None
                if side is not None:
                    ev = torch.cuda.Event()
                    ev.record(side)
                plan.append((recv_grp * G * b, (recv_grp + 1) * G * b, ev))
        return plan

    def all_gather_cat(self, x: torch.Tensor) -> torch.Tensor:
        """[b, ...] on every rank -> [W*b, ...] in rank order (what torch.cat(all_gather) gives)."""
        x = x.contiguous()
        out = torch.empty((self.world_size * x.shape[0],) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
        dist.all_gather_into_tensor(out, x, group=self.group)
        return out

    def reduce_scatter_sum(self, x: torch.Tensor) -> torch.Tensor:
        """[W*b, ...] partials -> [b, ...] = sum over ranks of this rank's block."""
        x = x.contiguous()
        b = x.shape[0] // self.world_size
        out = torch.empty((b,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
        if x.is_cuda:
            dist.reduce_scatter_tensor(out, x, op=dist.ReduceOp.SUM, group=self.group)
        else:
            # gloo has no reduce_scatter: all_reduce and slice (CPU host-logic tests only)
            y = x.clone()
            dist.all_reduce(y, op=dist.ReduceOp.SUM, group=self.group)
            out.copy_(y[self.rank * b:(self.rank + 1) * b])
        return out


class _NullCtx:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


@dataclass
class _Cfg:
    local_loss: bool
    gather_with_grad: bool
    rank: int
    world_size: int
    kernels: object
    comm: Optional[_Comm]
    comm_dtype: Optional[torch.dtype]
    pipeline: bool = True
    symm: bool = True
    stream_fwd: bool = True
    pull_streams: int = 2
    exchange: str = "pull"
    will_backward: bool = True       # a backward of this forward is expected (grad mode on, some input requires grad)
    gather_buf: Optional[torch.Tensor] = None   # captured steps: the symmetric gather buffer (symm.make_graph_gather)


def _as_scale_tensor(logit_scale, device) -> torch.Tensor:
    if torch.is_tensor(logit_scale):
        return logit_scale.detach().to(device=device, dtype=torch.float32).reshape(1).contiguous()
    return torch.full((1,), float(logit_scale), dtype=torch.float32, device=device)


class _FusedClipLoss(torch.autograd.Function):
    """loss = ClipLoss.forward(...) of the reference (loss.py:128-139), all four distributed modes."""

    @staticmethod
    def forward(ctx, img: torch.Tensor, txt: torch.Tensor, logit_scale, cfg: _Cfg):
        K, W, r = cfg.kernels, cfg.world_size, cfg.rank
        scale = _as_scale_tensor(logit_scale, img.device)
        b = img.shape[0]
        img_all = txt_all = None
        loss = None
        if W == 1:
            off = 0
            row_lse, col_lse, diag = K.clip_fwd(img, txt, scale, 0)
        elif cfg.local_loss and cfg.gather_with_grad and cfg.pipeline:
            # performance mode: chunk-pipelined gather, K1 runs on column block k while block k+1 is in flight
            off = b * r
            B = b * W
            sx = cfg.comm.symm_exchange(txt) if cfg.symm else None
            if cfg.gather_buf is not None and sx is not None and tuple(cfg.gather_buf.shape) == (B, txt.shape[1]):
                txt_all = cfg.gather_buf
            else:
                txt_all = torch.empty((B, txt.shape[1]), dtype=txt.dtype, device=txt.device)
            streamed = (sx is not None and cfg.stream_fwd and hasattr(K, "supports_fwd_stream")
                        and K.supports_fwd_stream(img, b, W))
            if streamed:
                # K1 fused with the exchange: ONE persistent launch consumes the gather buffer block by block, gated by
                # the ready flags the copy stream writes behind each block
                sx.begin_step()
                # push exchange: the gather buffer is peer-writable and saved for the backward, which releases it with
                # a barrier.  A forward without a backward to come (no_grad / evaluation), or one issued while another
                # forward's backward is still outstanding, therefore takes the pull exchange (private gather buffer);
                # every rank sees the same call sequence and so takes the same branch.
                pushed = cfg.exchange == "push" and cfg.will_backward and not sx.bwd_pending
                if pushed:
                    txt_all, order, wait, flags = sx.gather_pushed(txt, max(cfg.pull_streams, 1))
                    sx.bwd_pending = True
                    ctx.pushed = sx
                else:
                    order, wait = sx.gather_streamed(txt, txt_all, cfg.pull_streams)
                    flags = sx.flags
                row_lse, _, diag = K.clip_fwd_stream(img, txt_all, scale, off, order, wait, b, flags, sx.epoch,
                                                     col_out=sx.col_buffer())
                plan = []
            elif sx is not None:
                sx.begin_step()
                plan = sx.gather_pipelined(txt, txt_all)          # copy-engine pulls from peer memory
                col_part = sx.col_buffer()
            else:
                plan = cfg.comm.pipelined_gather(txt, txt_all)    # batched NCCL P2P rounds
                col_part = torch.empty(B, dtype=torch.float32, device=txt.device)
            blocks = (not streamed) and hasattr(K, "supports_fwd_blocks") and K.supports_fwd_blocks(img)
            if streamed:
                pass
            elif blocks:
                # one K1 launch per arriving block, the row / column reductions of all blocks once at the end
                st = K.clip_fwd_blocks_begin(img, [hi - lo for lo, hi, _ in plan], col_out=col_part)
            else:
                diag = torch.empty(b, dtype=torch.float32, device=txt.device)
                row_parts = []
            for lo, hi, ev in plan:
                for e in (ev if isinstance(ev, (list, tuple)) else [ev]):
                    if e is not None:
                        torch.cuda.current_stream().wait_event(e)
                lab = off - lo if lo <= off < hi else -1          # only the rank's own block holds its labels
                if blocks:
                    K.clip_fwd_block(st, img, txt_all[lo:hi], scale, lab, lo)
                else:
                    rp, _, _ = K.clip_fwd(img, txt_all[lo:hi], scale, lab, col_out=col_part[lo:hi], diag_out=diag)
                    row_parts.append(rp)
            if streamed:
                pass
            elif blocks:
                row_lse, _, diag = K.clip_fwd_blocks_end(st)
            else:
                row_lse = K.lse_combine(torch.stack(row_parts)) if len(row_parts) > 1 else row_parts[0]
            if sx is not None:
                sx.end_gather(streamed)
                if hasattr(sx, "combine_cols_loss"):              # barrier + one kernel: combine, loss, epoch bump
                    col_lse, loss = sx.combine_cols_loss(K, row_lse, diag, off, streamed)
                else:
                    col_lse = sx.combine_cols(K)                  # barrier + one kernel over the W peer buffers
                ctx.symm = (sx, sx.step)
            else:
                side = cfg.comm.side_stream(txt.device)
                if side is not None:
                    torch.cuda.current_stream().wait_stream(side)     # sends of `txt` done before it can be freed
                parts = cfg.comm.all_gather_cat(col_part.reshape(1, -1))      # [W, B]
                col_lse = K.lse_combine(parts)
        elif cfg.local_loss:
            off = b * r
            txt_all = cfg.comm.all_gather_cat(txt)
            if not cfg.gather_with_grad:
                img_all = cfg.comm.all_gather_cat(img)       # needed by the swapped backward call
            row_lse, col_part, diag = K.clip_fwd(img, txt_all, scale, off)
            parts = cfg.comm.all_gather_cat(col_part.reshape(1, -1))      # [W, B]
            col_lse = K.lse_combine(parts)
        else:
            off = 0
            img_all = cfg.comm.all_gather_cat(img)
            txt_all = cfg.comm.all_gather_cat(txt)
            row_lse, col_lse, diag = K.clip_fwd(img_all, txt_all, scale, 0)
        if loss is None:
            loss = K.clip_loss(row_lse, diag, col_lse, off)
        if not hasattr(ctx, "symm"):
            ctx.symm = None
        if not hasattr(ctx, "pushed"):
            ctx.pushed = None
        ctx.cfg, ctx.off, ctx.b = cfg, off, b
        ctx.scale_meta = (logit_scale.shape, logit_scale.dtype) if torch.is_tensor(logit_scale) else None
        ctx.save_for_backward(img, txt, scale, row_lse, col_lse, img_all, txt_all)
        return loss

    @staticmethod
    def backward(ctx, g):
        img, txt, scale, row_lse, col_lse, img_all, txt_all = ctx.saved_tensors
        cfg, off, b = ctx.cfg, ctx.off, ctx.b
        K, W = cfg.kernels, cfg.world_size
        need_i, need_t = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        need_s = ctx.needs_input_grad[2]
        fdt = img.dtype
        d_img = d_txt = None
        if W == 1:
            B = b
            d_img, d_txt, ds = K.clip_bwd(img, txt, scale, 0, row_lse, col_lse, 0.5 / B, 0.5 / B, 1.0 / B, g,
                                          need_i, need_t, fdt)
        elif cfg.local_loss and cfg.gather_with_grad:
            cdt = cfg.comm_dtype or fdt
            w = (0.5 / b, 0.5 / b, 1.0 / b)
            side = cfg.comm.side_stream(img.device) if cfg.pipeline else None
            sx = None
            if ctx.symm is not None and ctx.symm[0].step == ctx.symm[1] and cdt == torch.bfloat16:
                sx = ctx.symm[0]             # same step as the forward: its slot parity is still current
            if need_t and need_i and sx is not None:
                # dS + dB GEMM (partial written straight into the symmetric buffer); peers pull their blocks with
                # the copy engines while the dA GEMM (reusing the staged dS) runs; one kernel sums the W blocks
                _, _, ds, ws = K.clip_bwd(img, txt_all, scale, off, row_lse, col_lse, *w, g, False, True, cdt,
                                          return_ws=True, dB_out=sx.dT_buffer())
                sx.reduce_scatter_begin()
                d_img, _, _ = K.clip_bwd(img, txt_all, scale, off, row_lse, col_lse, *w, g, True, False, fdt,
                                         ws=ws, reuse_ds=True)
                d_txt = sx.reduce_scatter_end(K).to(fdt)
            elif need_t and need_i and cfg.pipeline:
                # dS + dB GEMM first, reduce-scatter of the [B, D] partial on the side stream while the dA GEMM
                # (which reuses the staged dS) runs on the main stream
                _, d_txt_all, ds, ws = K.clip_bwd(img, txt_all, scale, off, row_lse, col_lse, *w, g, False, True, cdt,
                                                  return_ws=True)
                if side is not None:
                    side.wait_stream(torch.cuda.current_stream())
                    d_txt_all.record_stream(side)
                with (torch.cuda.stream(side) if side is not None else _NullCtx()):
                    d_txt = cfg.comm.reduce_scatter_sum(d_txt_all)
                d_img, _, _ = K.clip_bwd(img, txt_all, scale, off, row_lse, col_lse, *w, g, True, False, fdt,
                                         ws=ws, reuse_ds=True)
                if side is not None:
                    torch.cuda.current_stream().wait_stream(side)
                    d_txt.record_stream(torch.cuda.current_stream())
                d_txt = d_txt.to(fdt)
            else:
                d_img, d_txt_all, ds = K.clip_bwd(img, txt_all, scale, off, row_lse, col_lse, *w, g, need_i, need_t,
                                                  cdt if need_t else fdt)
                if need_i and d_img.dtype != fdt:
                    d_img = d_img.to(fdt)
                if need_t:
                    d_txt = cfg.comm.reduce_scatter_sum(d_txt_all).to(fdt)
            if ctx.pushed is not None:
                ctx.pushed.push_step_done()     # releases the peer-writable gather buffer of this step
        elif cfg.local_loss:
            ds = torch.zeros((), dtype=torch.float32, device=img.device)
            if need_i or need_s:
                d_img, _, ds1 = K.clip_bwd(img, txt_all, scale, off, row_lse, col_lse, 0.5 / b, 0.0, 0.5 / b, g,
                                           need_i, False, fdt)
                ds = ds + ds1
            if need_t or need_s:
                never = torch.full((img_all.shape[0],), float("inf"), dtype=torch.float32, device=img.device)
                d_txt, _, ds2 = K.clip_bwd(txt, img_all, scale, off, col_lse[off:off + b].contiguous(), never,
                                           0.5 / b, 0.0, 0.5 / b, g, need_t, False, fdt)
                ds = ds + ds2
        else:
            B = img_all.shape[0]
            cdt = (cfg.comm_dtype or fdt) if cfg.gather_with_grad else fdt
            d_ia, d_ta, ds = K.clip_bwd(img_all, txt_all, scale, 0, row_lse, col_lse, 0.5 / B, 0.5 / B, 1.0 / B, g,
                                        need_i, need_t, cdt)
            if cfg.gather_with_grad:
                d_img = cfg.comm.reduce_scatter_sum(d_ia).to(fdt) if need_i else None
                d_txt = cfg.comm.reduce_scatter_sum(d_ta).to(fdt) if need_t else None
            else:
                lo, hi = cfg.rank * b, (cfg.rank + 1) * b
                d_img = d_ia[lo:hi].contiguous() if need_i else None
                d_txt = d_ta[lo:hi].contiguous() if need_t else None
        d_scale = None
        if need_s and ctx.scale_meta is not None:
            shape, dtype = ctx.scale_meta
            d_scale = ds.to(dtype).reshape(shape)
        return d_img, d_txt, d_scale, None


class _ChunkedClipLoss(torch.autograd.Function):
    """Single-process loss over features given as CHUNKS of which only some carry gradient -- the gradient-accumulation
    path of the reference (src/others/train_other.py:140-197): cached no-grad features of the other micro-batches are
    concatenated with one live chunk and the loss is taken over the whole batch.  The forward is the full fused forward;
    the backward forms gradients only for the live row / column blocks (two rectangular K2 calls per live chunk with
    the saved full-batch log-sum-exps) instead of the full dI / dT the reference computes and then discards."""

    @staticmethod
    def forward(ctx, logit_scale, cfg, n, *chunks):
        K = cfg.kernels
        imgs, txts = chunks[:n], chunks[n:]
        img_all = torch.cat([c.detach() for c in imgs], dim=0).contiguous()
        txt_all = torch.cat([c.detach() for c in txts], dim=0).contiguous()
        scale = _as_scale_tensor(logit_scale, img_all.device)
        row_lse, col_lse, diag = K.clip_fwd(img_all, txt_all, scale, 0)
        loss = K.clip_loss(row_lse, diag, col_lse, 0)
        ctx.cfg, ctx.n = cfg, n
        ctx.sizes = [int(c.shape[0]) for c in imgs]
        ctx.scale_meta = (logit_scale.shape, logit_scale.dtype) if torch.is_tensor(logit_scale) else None
        ctx.save_for_backward(img_all, txt_all, scale, row_lse, col_lse)
        return loss

    @staticmethod
    def backward(ctx, g):
        img_all, txt_all, scale, row_lse, col_lse = ctx.saved_tensors
        K, n = ctx.cfg.kernels, ctx.n
        B = img_all.shape[0]
        fdt = img_all.dtype
        w_row, w_col, w_diag = 0.5 / B, 0.5 / B, 1.0 / B
        need = ctx.needs_input_grad
        grads = [None] * (3 + 2 * n)
        if need[0] and ctx.scale_meta is not None:
            # d(logit_scale) is a sum over the whole matrix: the dS producer alone, no gradient GEMM
            _, _, ds = K.clip_bwd(img_all, txt_all, scale, 0, row_lse, col_lse, w_row, w_col, w_diag, g,
                                  False, False, fdt)
            shape, dtype = ctx.scale_meta
            grads[0] = ds.to(dtype).reshape(shape)
        lo = 0
        for k, size in enumerate(ctx.sizes):
            hi = lo + size
            if need[3 + k]:           # image chunk k: rows [lo, hi) of S
                grads[3 + k], _, _ = K.clip_bwd(img_all[lo:hi], txt_all, scale, lo, row_lse[lo:hi].contiguous(), col_lse,
                                                w_row, w_col, w_diag, g, True, False, fdt)
            if need[3 + n + k]:       # text chunk k: columns [lo, hi) of S = rows of S^T (row / column roles swap)
                grads[3 + n + k], _, _ = K.clip_bwd(txt_all[lo:hi], img_all, scale, lo, col_lse[lo:hi].contiguous(),
                                                    row_lse, w_col, w_row, w_diag, g, True, False, fdt)
            lo = hi
        return tuple(grads)


# ------------------------------------------------------------------------------------------------
# CUDA-graph replay of one loss step (launch-bound regimes: small batches, 8-GPU shards)
# ------------------------------------------------------------------------------------------------
class _GraphedStep:
    """Forward and backward of `_FusedClipLoss` captured once into two CUDA graphs (the pattern of
    torch.cuda.make_graphed_callables): all kernels, copy-engine pulls, barriers and stream fork/joins of a step are
    replayed with two graph launches, which removes the ~100 host-side launches per step that bound the 8-GPU shard
    (GPU work per step < 1 ms).  Inputs are copied into static buffers; outputs are cloned out of them."""

    def __init__(self, img, txt, logit_scale, cfg: "_Cfg"):
        dev = img.device
        self.cfg = cfg
        self.s_img = img.detach().clone().requires_grad_(img.requires_grad)
        self.s_txt = None
        if (cfg.world_size > 1 and cfg.local_loss and cfg.gather_with_grad and cfg.pipeline and cfg.symm
                and cfg.stream_fwd and cfg.exchange == "pull"):
            # the captured step's gather buffer lives in symmetric memory; its own block is the text input slot
            sx = cfg.comm.symm_exchange(txt)
            if sx is not None and hasattr(sx, "make_graph_gather") and hasattr(cfg.kernels, "supports_fwd_stream") \
                    and cfg.kernels.supports_fwd_stream(img, img.shape[0], cfg.world_size):
                cfg.gather_buf = sx.make_graph_gather()
                self.s_txt = sx.own_block(cfg.gather_buf).detach()
                self.s_txt.copy_(txt.detach())
                self.s_txt.requires_grad_(txt.requires_grad)
        if self.s_txt is None:
            self.s_txt = txt.detach().clone().requires_grad_(txt.requires_grad)
        self.scale_is_tensor = torch.is_tensor(logit_scale)
        if self.scale_is_tensor:
            self.s_scale = logit_scale.detach().clone().to(dev).requires_grad_(logit_scale.requires_grad)
        else:
            self.s_scale = float(logit_scale)
        self.s_g = torch.ones((), dtype=torch.float32, device=dev)
        self.inputs = [t for t in (self.s_img, self.s_txt, self.s_scale if self.scale_is_tensor else None)
                       if t is not None and t.requires_grad]
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):                                 # eager warm-up: lazy initialisations happen here
                loss = _FusedClipLoss.apply(self.s_img, self.s_txt, self.s_scale, cfg)
                if self.inputs:
                    torch.autograd.grad(loss, self.inputs, self.s_g)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize(dev)
        from . import _lib
        self._lib = _lib
        self.pool = torch.cuda.graph_pool_handle()
        self.fwd = torch.cuda.CUDAGraph()
        n0 = _lib.launch_count()
        with torch.cuda.graph(self.fwd, pool=self.pool):
            self.s_loss = _FusedClipLoss.apply(self.s_img, self.s_txt, self.s_scale, cfg)
        self.n_fwd = _lib.launch_count() - n0          # library kernels recorded into the forward graph
        self.bwd = None
        self.s_grads = ()
        self.n_bwd = 0
        if self.inputs:
            self.bwd = torch.cuda.CUDAGraph()
            n0 = _lib.launch_count()
            with torch.cuda.graph(self.bwd, pool=self.pool):
                self.s_grads = torch.autograd.grad(self.s_loss, self.inputs, self.s_g)
            self.n_bwd = _lib.launch_count() - n0
        self.pending = False            # a replayed forward whose backward has not run yet

    def input_slots(self):
        """The static input tensors of the captured step.  A producer that writes its features straight into them
        (e.g. `l2_normalize(x, out=slot)`) and passes them to the loss skips the per-step input copies."""
        return self.s_img.detach(), self.s_txt.detach()

    def run_forward(self, img, txt, logit_scale):
        # the saved activations of a captured step live in ONE set of static buffers: a second forward before the
        # first one's backward would silently overwrite them (the caller checks `pending` and runs eagerly instead)
        assert not self.pending
        if img.data_ptr() != self.s_img.data_ptr():
            self.s_img.detach().copy_(img)
        if txt.data_ptr() != self.s_txt.data_ptr():
            self.s_txt.detach().copy_(txt)
        if self.scale_is_tensor:
            self.s_scale.detach().copy_(logit_scale.detach())
        self.fwd.replay()
        self._lib.note_replayed(self.n_fwd)
        self.pending = self.bwd is not None
        return self.s_loss.detach().clone()

    def run_backward(self, g):
        self.s_g.copy_(g)
        self.bwd.replay()
        self._lib.note_replayed(self.n_bwd)
        self.pending = False
        return [t.clone() for t in self.s_grads]


class _GraphedClipLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, img, txt, logit_scale, step: "_GraphedStep"):
        ctx.step = step
        ctx.scale_is_tensor = torch.is_tensor(logit_scale)
        return step.run_forward(img, txt, logit_scale)

    @staticmethod
    def backward(ctx, g):
        step = ctx.step
        grads = iter(step.run_backward(g))
        d_img = next(grads) if step.s_img.requires_grad else None
        d_txt = next(grads) if step.s_txt.requires_grad else None
        d_s = next(grads) if (step.scale_is_tensor and step.s_scale.requires_grad) else None
        return d_img, d_txt, d_s, None


# ------------------------------------------------------------------------------------------------
# reference API
# ------------------------------------------------------------------------------------------------
def gather_features(
        image_features,
        text_features,
        local_loss=False,
        gather_with_grad=False,
        rank=0,
        world_size=1,
        use_horovod=False
):
    """Same contract as the reference's ``gather_features`` (loss.py:21-65): returns
    ``(all_image_features, all_text_features)``, the per-rank blocks concatenated in rank order,
    autograd-connected exactly as the reference (see module docstring).  This is the materialising
    helper kept for API compatibility (``get_logits`` and subclasses use it); the fused
    ``ClipLoss.forward`` does its own gather."""
    assert has_distributed, 'torch.distributed did not import correctly, please use a PyTorch version with support.'
    if use_horovod:
        raise NotImplementedError("Horovod is not available in xtag_clip_b200 (NCCL/torch.distributed only)")
    if gather_with_grad:
        all_image_features = torch.cat(torch.distributed.nn.all_gather(image_features), dim=0)
        all_text_features = torch.cat(torch.distributed.nn.all_gather(text_features), dim=0)
    else:
        comm = _Comm(world_size, rank)
        with torch.no_grad():
            gi = list(comm.all_gather_cat(image_features).chunk(world_size, dim=0))
            gt = list(comm.all_gather_cat(text_features).chunk(world_size, dim=0))
        if not local_loss:
            # ensure grads for local rank when all_* features don't have a gradient
            gi[rank] = image_features
            gt[rank] = text_features
        all_image_features = torch.cat(gi, dim=0)
        all_text_features = torch.cat(gt, dim=0)
    return all_image_features, all_text_features


class ClipLoss(nn.Module):
    """``ClipLoss(local_loss=False, gather_with_grad=False, cache_labels=False, rank=0, world_size=1,
    use_horovod=False)`` -- reference signature (loss.py:68-89).  ``forward(image_features,
    text_features, logit_scale, output_dict=False)`` returns a 0-d fp32 tensor or
    ``{"contrastive_loss": tensor}`` (loss.py:128-139)."""

    def __init__(
            self,
            local_loss=False,
            gather_with_grad=False,
            cache_labels=False,
            rank=0,
            world_size=1,
            use_horovod=False,
            *,
            group=None,
            comm_dtype: Optional[torch.dtype] = None,
            pipeline: bool = True,
            symmetric_memory: bool = True,
            stream_forward: bool = True,
            pull_streams: int = 2,
            exchange: Optional[str] = None,
            cuda_graph: bool = False,
            compute_dtype: Optional[torch.dtype] = None,
            _kernels=None,
    ):
        super().__init__()
        if use_horovod:
            raise NotImplementedError("use_horovod=True: Horovod is not supported by xtag_clip_b200; "
                                      "use torch.distributed (NCCL)")
        self.local_loss = local_loss
        self.gather_with_grad = gather_with_grad
        self.cache_labels = cache_labels
        self.rank = rank
        self.world_size = world_size
        self.use_horovod = use_horovod
        self._group = group
        self._comm_dtype = comm_dtype
        self._pipeline = pipeline
        self._symm = symmetric_memory
        self._stream_fwd = stream_forward
        self._pull_streams = pull_streams
        # "pull" (validated) or "push" (experimental: no start-of-step barrier); default from XTAG_EXCHANGE
        self._exchange = exchange or os.environ.get("XTAG_EXCHANGE", "pull")
        if self._exchange not in ("pull", "push"):
            raise ValueError(f"exchange must be 'pull' or 'push', got {self._exchange!r}")
        self._cuda_graph = cuda_graph
        # None: bf16 inputs run the tcgen05 path, fp32 inputs the exact fp32 path -- except under torch.autocast(bf16),
        # where fp32 features are cast to bf16 first, as the reference's own matmul is by autocast (train.py amp_bf16)
        if compute_dtype not in (None, torch.bfloat16, torch.float32):
            raise ValueError(f"compute_dtype must be None, torch.bfloat16 or torch.float32, got {compute_dtype}")
        self._compute_dtype = compute_dtype
        self._graphs = {}
        self._kernels = _kernels
        self._comm = None

        # cache state
        self.prev_num_logits = 0
        self.labels = {}

    # -- helpers kept for API compatibility (CoCaLoss / DistillClipLoss call them, loss.py:170, 202-208)
    def get_ground_truth(self, device, num_logits) -> torch.Tensor:
        if self.prev_num_logits != num_logits or device not in self.labels:
            labels = torch.arange(num_logits, device=device, dtype=torch.long)
            if self.world_size > 1 and self.local_loss:
                labels = labels + num_logits * self.rank
            if self.cache_labels:
                self.labels[device] = labels
                self.prev_num_logits = num_logits
        else:
            labels = self.labels[device]
        return labels

    def get_logits(self, image_features, text_features, logit_scale):
        """Materialising slow path with the reference's exact expressions (loss.py:104-126).
        Not used by ``forward``."""
        if self.world_size > 1:
            all_image_features, all_text_features = gather_features(
                image_features, text_features,
                local_loss=self.local_loss, gather_with_grad=self.gather_with_grad,
                rank=self.rank, world_size=self.world_size, use_horovod=self.use_horovod)
            if self.local_loss:
                logits_per_image = logit_scale * image_features @ all_text_features.T
                logits_per_text = logit_scale * text_features @ all_image_features.T
            else:
                logits_per_image = logit_scale * all_image_features @ all_text_features.T
                logits_per_text = logits_per_image.T
        else:
            logits_per_image = logit_scale * image_features @ text_features.T
            logits_per_text = logit_scale * text_features @ image_features.T
        return logits_per_image, logits_per_text

    def _cfg(self, will_backward: bool = True, exchange: Optional[str] = None) -> _Cfg:
        k = self._kernels
        if k is None:
            from .kernels import default_kernels
            k = default_kernels()
        comm = None
        if self.world_size > 1:
            assert has_distributed, 'torch.distributed did not import correctly, please use a PyTorch version with support.'
            if self._comm is None:
                self._comm = _Comm(self.world_size, self.rank, self._group)     # keeps its side stream
            comm = self._comm
        return _Cfg(self.local_loss, self.gather_with_grad, self.rank, self.world_size, k, comm, self._comm_dtype,
                    self._pipeline, self._symm, self._stream_fwd, self._pull_streams, exchange or self._exchange,
                    will_backward)

    def forward(self, image_features, text_features, logit_scale, output_dict=False):
        if image_features.dim() != 2 or image_features.shape != text_features.shape:
            raise ValueError(f"image_features {tuple(image_features.shape)} and text_features "
                             f"{tuple(text_features.shape)} must both be [batch, dim]")
        # One compute dtype for both operands: bf16 (tcgen05 path) or fp32 (exact path).  The reference's train loop calls
        # the loss inside torch.autocast(bf16) on fp32 features (F.normalize is promoted to fp32 by autocast) and its
        # B x B matmul then runs in bf16: under bf16 autocast -- or with compute_dtype=torch.bfloat16 -- fp32 features
        # are cast to bf16 here (a differentiable cast: the feature gradients come back in the input dtype).
        cd = self._compute_dtype
        if cd is None:
            both_bf16 = image_features.dtype == torch.bfloat16 and text_features.dtype == torch.bfloat16
            amp_bf16 = (image_features.is_cuda and torch.is_autocast_enabled("cuda")
                        and torch.get_autocast_dtype("cuda") == torch.bfloat16)
            cd = torch.bfloat16 if (both_bf16 or amp_bf16) else torch.float32
        img = image_features if image_features.dtype == cd else image_features.to(cd)
        txt = text_features if text_features.dtype == cd else text_features.to(cd)
        img, txt = img.contiguous(), txt.contiguous()
        will_bwd = torch.is_grad_enabled() and (img.requires_grad or txt.requires_grad or
                                                (torch.is_tensor(logit_scale) and logit_scale.requires_grad))
        if self._cuda_graph and img.is_cuda and torch.is_grad_enabled():
            total_loss = self._forward_graphed(img, txt, logit_scale, will_bwd)
        else:
            total_loss = _FusedClipLoss.apply(img, txt, logit_scale, self._cfg(will_bwd))
        return {"contrastive_loss": total_loss} if output_dict else total_loss

    def graph_input_slots(self, batch: int, dim: int, dtype: torch.dtype = torch.bfloat16):
        """`cuda_graph=True`: the static (image, text) input tensors of the captured step for [batch, dim] features, or
        None before that step has been captured.  Features written straight into them (e.g. by
        `l2_normalize(x, out=slot)`) and passed to `forward` are consumed in place: no per-step input copy."""
        for key, step in self._graphs.items():
            if step and key[0] == (batch, dim) and key[1] == dtype:
                return step.input_slots()
        return None

    def forward_chunks(self, image_chunks, text_chunks, logit_scale, output_dict=False):
        """Gradient-accumulation form of `forward` (src/others/train_other.py:183-191 concatenates cached no-grad
        features with one live chunk): `image_chunks` / `text_chunks` are equally long lists of [b_k, D] tensors whose
        concatenation is the batch; chunks with `requires_grad=False` get no gradient work.  Equivalent to
        `forward(torch.cat(image_chunks), torch.cat(text_chunks), logit_scale)`."""
        if len(image_chunks) != len(text_chunks) or not image_chunks:
            raise ValueError("image_chunks and text_chunks must be equally long, non-empty lists")
        for a, c in zip(image_chunks, text_chunks):
            if a.dim() != 2 or a.shape != c.shape or a.shape[1] != image_chunks[0].shape[1]:
                raise ValueError("every image/text chunk pair must be [b_k, D] with one common D")
        live = sum(int(t.requires_grad) for t in list(image_chunks) + list(text_chunks))
        dts = {t.dtype for t in list(image_chunks) + list(text_chunks)}
        if self.world_size > 1 or live == 2 * len(image_chunks) or len(dts) != 1 or not torch.is_grad_enabled():
            # nothing to skip (or a sharded loss, whose gather needs whole tensors): the ordinary path
            return self.forward(torch.cat(list(image_chunks), dim=0), torch.cat(list(text_chunks), dim=0), logit_scale,
                                output_dict)
        chunks = [c if c.dtype != torch.float16 else c.float() for c in list(image_chunks) + list(text_chunks)]
        total_loss = _ChunkedClipLoss.apply(logit_scale, self._cfg(), len(image_chunks), *chunks)
        return {"contrastive_loss": total_loss} if output_dict else total_loss

    def _forward_graphed(self, img, txt, logit_scale, will_bwd: bool = True):
        """`cuda_graph=True`: one captured step per (shape, dtype, requires_grad) signature.  All ranks must take
        the same path (the captured step contains the cross-rank barriers), so the ranks agree on whether the capture
        succeeded before anyone uses it."""
        st = torch.is_tensor(logit_scale)
        key = (tuple(img.shape), img.dtype, img.requires_grad, txt.requires_grad, st,
               bool(st and logit_scale.requires_grad), tuple(logit_scale.shape) if st else None)
        step = self._graphs.get(key)
        if step is None:
            err = None
            try:
                step = _GraphedStep(img, txt, logit_scale, self._cfg(will_bwd))
            except Exception as e:      # capture refused (e.g. an op that is illegal under capture): stay eager
                err, step = e, False
            if self.world_size > 1:
                ok = torch.tensor([0 if step is False else 1], device=img.device, dtype=torch.int32)
                dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self._group)
                if int(ok) == 0 and step is not False:
                    step = False        # another rank failed: nobody replays a step with cross-rank barriers alone
            if step is False:
                import warnings
                why = f"{type(err).__name__}: {err}" if err is not None else "capture failed on another rank"
                warnings.warn(f"xtag_clip_b200: CUDA-graph capture failed ({why}); running eagerly")
            self._graphs[key] = step
        if step is False:
            return _FusedClipLoss.apply(img, txt, logit_scale, self._cfg(will_bwd))
        if step.pending:
            # a second forward of this signature before the first one's backward: the captured step has ONE set of
            # saved activations, so this call runs eagerly (pull exchange: it must not touch the peer-writable gather
            # buffer the pending backward still reads)
            return _FusedClipLoss.apply(img, txt, logit_scale, self._cfg(will_bwd, exchange="pull"))
        return _GraphedClipLoss.apply(img, txt, logit_scale, step)

    @property
    def last_path(self):
        """Diagnostics for benchmarks: which launch / exchange path the module is set up to use."""
        graphs = [bool(v) for v in self._graphs.values()]
        ex = None
        if self._comm is not None:
            sx = [v for v in self._comm._symm.values() if v is not None]
            ex = ("symmetric-memory " + ("push" if (sx and sx[0]._pushed) else "pull")) if sx else "nccl"
        return dict(cuda_graph=bool(graphs) and all(graphs), exchange=ex)


def create_loss(args):
    """Reference factory contract (factory.py:433-469) for the losses in scope: the fused ``ClipLoss`` configured from
    ``args.local_loss / gather_with_grad / rank / world_size / horovod``, or the fused ``SigLipLoss`` for
    ``args.siglip``.  Distill / CoCa losses are outside the hot path."""
    if getattr(args, "distill", False) or "coca" in str(getattr(args, "model", "")).lower():
        raise NotImplementedError("xtag_clip_b200.create_loss covers the ClipLoss / SigLipLoss paths only "
                                  "(--distill and coca models are out of scope)")
    if getattr(args, "siglip", False):
        assert not args.horovod, "Horovod not currently supported for SigLip"
        from .siglip import SigLipLoss
        return SigLipLoss(rank=args.rank, world_size=args.world_size, dist_impl=getattr(args, "loss_dist_impl", None))
    return ClipLoss(
        local_loss=args.local_loss,
        gather_with_grad=args.gather_with_grad,
        cache_labels=True,
        rank=args.rank,
        world_size=args.world_size,
        use_horovod=args.horovod,
    )
