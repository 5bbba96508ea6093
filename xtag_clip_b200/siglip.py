"""Drop-in for the reference's ``SigLipLoss`` (src/open_clip/loss.py:314-448), the sigmoid sibling of ``ClipLoss`` on
the same tcgen05 mainloop.

    SigLipLoss(cache_labels=False, rank=0, world_size=1, dist_impl=None)
    forward(image_features, text_features, logit_scale, logit_bias, output_dict=False)

The sigmoid loss has no row / column normaliser, so ONE pass over the logit tiles (``xtag_siglip_fwd``) yields the
loss, d/d logit_scale, d/d logit_bias AND the complete logit gradient dS, staged as bf16; the backward is the two
gradient GEMMs of the contrastive head on that dS (``xtag_clip_bwd`` with ``XTAG_BWD_REUSE_DS``): 6 B^2 D executed
FLOPs per step, the algorithmic count, and the logits never reach HBM.

Sharded (world_size > 1): rank r computes its b rows against ALL B text columns with the positives on its own block
-- the same terms the reference sums over its neighbour exchanges (every ``dist_impl`` variant adds the own block with
positives and each other rank's block ``negative_only``); the text features are all-gathered and the text-gradient
partials reduce-scattered back, which is what the reference's autograd-aware exchanges do in aggregate.
bf16 features (or fp32 features under ``torch.autocast(bf16)`` / ``compute_dtype=torch.bfloat16``), D % 8 == 0.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from .loss import _Comm, _as_scale_tensor


class _FusedSigLip(torch.autograd.Function):
    @staticmethod
    def forward(ctx, img, txt, logit_scale, logit_bias, K, comm, rank, world, will_backward):
        b = img.shape[0]
        scale = _as_scale_tensor(logit_scale, img.device)
        bias = _as_scale_tensor(logit_bias, img.device)
        if world > 1:
            txt_all = comm.all_gather_cat(txt)
            off = rank * b
        else:
            txt_all, off = txt, 0
        out3, ws = K.siglip_fwd(img, txt_all, scale, bias, off, 1.0 / b, will_backward)
        ctx.meta = (K, comm, world, b, off,
                    (logit_scale.shape, logit_scale.dtype) if torch.is_tensor(logit_scale) else None,
                    (logit_bias.shape, logit_bias.dtype) if torch.is_tensor(logit_bias) else None)
        ctx.ws = ws
        ctx.save_for_backward(img, txt_all, scale, out3)
        return out3[0].clone()

    @staticmethod
    def backward(ctx, g):
        img, txt_all, scale, out3 = ctx.saved_tensors
        K, comm, world, b, off, s_meta, b_meta = ctx.meta
        need_i, need_t = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        g32 = g.detach().float().reshape(1)
        d_img = d_txt = None
        if need_i or need_t:
            if ctx.ws is None:
                raise RuntimeError("SigLipLoss: the forward ran without staging the logit gradient (no_grad)")
            gs = (g32 * scale).contiguous()                  # GEMM alpha = upstream gradient * logit scale
            dummy = out3                                     # (row / column LSEs are unused with REUSE_DS)
            d_img, d_txt_all, _ = K.clip_bwd(img, txt_all, gs, max(off, 0), dummy, dummy, 0.0, 0.0, 0.0, g32, need_i,
                                             need_t, img.dtype, ws=ctx.ws, reuse_ds=True)
            if need_t:
                d_txt = comm.reduce_scatter_sum(d_txt_all) if world > 1 else d_txt_all
        d_s = d_b = None
        if ctx.needs_input_grad[2] and s_meta is not None:
            d_s = (g32 * out3[1]).to(s_meta[1]).reshape(s_meta[0])
        if ctx.needs_input_grad[3] and b_meta is not None:
            d_b = (g32 * out3[2]).to(b_meta[1]).reshape(b_meta[0])
        return d_img, d_txt, d_s, d_b, None, None, None, None, None


class SigLipLoss(nn.Module):
    """Reference signature (loss.py:325-341).  ``dist_impl`` is accepted for compatibility: every variant of the
    reference sums the same terms, which the sharded kernel computes in one launch per rank."""

    def __init__(self, cache_labels: bool = False, rank: int = 0, world_size: int = 1, dist_impl: Optional[str] = None,
                 *, group=None, compute_dtype: Optional[torch.dtype] = None, _kernels=None):
        super().__init__()
        self.cache_labels = cache_labels
        self.rank = rank
        self.world_size = world_size
        self.dist_impl = dist_impl or 'bidir'
        assert self.dist_impl in ('bidir', 'shift', 'reduce', 'gather')
        self._group = group
        self._compute_dtype = compute_dtype
        self._kernels = _kernels
        self._comm = None
        self.prev_num_logits = 0
        self.labels = {}

    def get_ground_truth(self, device, dtype, num_logits, negative_only=False) -> torch.Tensor:
        labels = -torch.ones((num_logits, num_logits), device=device, dtype=dtype)
        if not negative_only:
            labels = 2 * torch.eye(num_logits, device=device, dtype=dtype) + labels
        return labels

    def get_logits(self, image_features, text_features, logit_scale, logit_bias=None):
        """Materialising helper with the reference's expression (loss.py:345-349); not used by ``forward``."""
        logits = logit_scale * image_features @ text_features.T
        if logit_bias is not None:
            logits += logit_bias
        return logits

    def forward(self, image_features, text_features, logit_scale, logit_bias, output_dict=False):
        if image_features.dim() != 2 or image_features.shape != text_features.shape:
            raise ValueError(f"image_features {tuple(image_features.shape)} and text_features "
                             f"{tuple(text_features.shape)} must both be [batch, dim]")
        if not image_features.is_cuda and self._kernels is None:      # (_kernels: the CPU contract model of the tests)
            raise RuntimeError("xtag_clip_b200.SigLipLoss runs on CUDA (sm_100a) tensors only; there is no CPU fallback")
        cd = self._compute_dtype
        if cd is None:
            both = image_features.dtype == torch.bfloat16 and text_features.dtype == torch.bfloat16
            amp = image_features.is_cuda and torch.is_autocast_enabled("cuda") and \
                torch.get_autocast_dtype("cuda") == torch.bfloat16
            cd = torch.bfloat16 if (both or amp) else None
        if cd != torch.bfloat16 or image_features.shape[1] % 8 != 0:
            raise NotImplementedError("SigLipLoss: the fused path needs bf16 features (or bf16 autocast / "
                                      "compute_dtype=torch.bfloat16) with dim % 8 == 0")
        img = image_features.to(cd).contiguous()
        txt = text_features.to(cd).contiguous()
        k = self._kernels
        if k is None:
            from .kernels import default_kernels
            k = default_kernels()
        if self.world_size > 1 and self._comm is None:
            self._comm = _Comm(self.world_size, self.rank, self._group)
        will_bwd = torch.is_grad_enabled() and any(torch.is_tensor(t) and t.requires_grad
                                                   for t in (img, txt, logit_scale, logit_bias))
        loss = _FusedSigLip.apply(img, txt, logit_scale, logit_bias, k, self._comm, self.rank, self.world_size, will_bwd)
        return {"contrastive_loss": loss} if output_dict else loss
