"""Oracle (test infrastructure): CPU restatement of the reference's sigmoid (SigLIP) loss.

Follows /root/reference/src/open_clip/loss.py:
  * :339-343  get_ground_truth  labels = -1 everywhere, +1 on the diagonal (all -1 when negative_only)
  * :345-349  get_logits        logit_scale * image_features @ text_features.T (+ logit_bias)
  * :351-360  _loss             -logsigmoid(labels * logits).sum() / image_features.shape[0]
  * :362-446  forward           own block with positives + every other rank's text block negative_only; the four
                                dist_impl variants ('bidir', 'shift', 'reduce', 'gather') exchange the text blocks
                                differently but sum the same terms, and every exchange is autograd-aware (a rank's
                                text gradient is the sum over the ranks that used its block)
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch
import torch.nn.functional as F


def siglip_block_loss(image_features, text_features, logit_scale, logit_bias, negative_only: bool = False):
    """SigLipLoss._loss (loss.py:351-360)."""
    logits = logit_scale * image_features @ text_features.T
    if logit_bias is not None:
        logits = logits + logit_bias
    n = image_features.shape[0]
    labels = -torch.ones((n, n), dtype=image_features.dtype)
    if not negative_only:
        labels = 2 * torch.eye(n, dtype=image_features.dtype) + labels
    return -F.logsigmoid(labels * logits).sum() / n


def siglip_loss_world(I_list: Sequence[torch.Tensor], T_list: Sequence[torch.Tensor], logit_scale, logit_bias
                      ) -> Tuple[List[torch.Tensor], List[torch.Tensor], List[torch.Tensor], List[torch.Tensor],
                                 List[torch.Tensor]]:
    """Emulates ``loss_r = SigLipLoss(rank=r, world_size=W)(I_r, T_r, s, b)`` + ``loss_r.backward()`` on every rank in
    one process -> (losses[r], dI[r], dT[r], dscale[r], dbias[r]); dT[k] sums the contributions of every rank that
    used rank k's text block (the exchanges' backward)."""
    W = len(I_list)
    dt = I_list[0].dtype
    losses, dI, dT, dS, dB = [], [], [torch.zeros_like(t) for t in T_list], [], []
    for r in range(W):
        s = torch.as_tensor(logit_scale, dtype=dt).clone().requires_grad_(True)
        b = torch.as_tensor(logit_bias, dtype=dt).clone().requires_grad_(True)
        I = I_list[r].clone().requires_grad_(True)
        Ts = [t.clone().requires_grad_(True) for t in T_list]
        loss = siglip_block_loss(I, Ts[r], s, b)
        for k in range(W):
            if k != r:
                loss = loss + siglip_block_loss(I, Ts[k], s, b, negative_only=True)
        loss.backward()
        losses.append(loss.detach())
        dI.append(I.grad)
        dS.append(s.grad)
        dB.append(b.grad)
        for k in range(W):
            if Ts[k].grad is not None:
                dT[k] = dT[k] + Ts[k].grad
    return losses, dI, dT, dS, dB
