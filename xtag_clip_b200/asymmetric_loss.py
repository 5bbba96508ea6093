"""Drop-in for ``AsymmetricLoss`` (reference src/open_clip/tagging_heads/asymmetric_loss.py:6-50):
same constructor, same ``forward(x, y) -> -loss.sum()``, computed by the fused K5 kernel (forward value
and d/dx in one pass; the focal weight carries no gradient, as in the reference which disables grad
around it)."""
from __future__ import annotations

import torch
import torch.nn as nn


class _Asl(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, y, gamma_neg, gamma_pos, clip, eps, kernels):
        if kernels is None:
            from .kernels import default_kernels
            kernels = default_kernels()
        xx = x.float() if x.dtype == torch.float16 else x
        loss, dx, _ = kernels.asl(xx, y, gamma_neg, gamma_pos, clip, eps, want_dx=True, want_idx=False)
        ctx.save_for_backward(dx)
        ctx.meta = (x.shape, x.dtype)
        return loss

    @staticmethod
    def backward(ctx, g):
        (dx,) = ctx.saved_tensors
        shape, dtype = ctx.meta
        return (dx * g).reshape(shape).to(dtype), None, None, None, None, None, None


class AsymmetricLoss(nn.Module):
    def __init__(self, gamma_neg=4, gamma_pos=1, clip=0.05, eps=1e-8, disable_torch_grad_focal_loss=True, *,
                 _kernels=None):
        super().__init__()
        if not disable_torch_grad_focal_loss:
            raise NotImplementedError("only disable_torch_grad_focal_loss=True (the reference's setting) is fused")
        self.gamma_neg = gamma_neg
        self.gamma_pos = gamma_pos
        self.clip = clip
        self.disable_torch_grad_focal_loss = disable_torch_grad_focal_loss
        self.eps = eps
        self._kernels = _kernels

    def forward(self, x, y):
        """x: input logits [b, 44]; y: multi-label binarised targets, same shape."""
        y = y.to(x.device)
        return _Asl.apply(x, y, self.gamma_neg, self.gamma_pos, self.clip, self.eps, self._kernels)
