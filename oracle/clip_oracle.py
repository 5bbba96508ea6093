"""Oracle (test infrastructure): CPU restatement of the open_clip contrastive head.

Follows /root/reference:
  * src/open_clip/model.py:311-313, 332-333   F.normalize on the feature epilogue
  * src/open_clip/loss.py:21-65                gather_features (rank order, which chunks carry grad)
  * src/open_clip/loss.py:91-102               get_ground_truth (arange + num_logits*rank)
  * src/open_clip/loss.py:104-126              get_logits  ((s*I) @ T.T ; local / global / W=1)
  * src/open_clip/loss.py:128-139              forward: (CE(Li) + CE(Lt)) / 2

Everything is plain torch on CPU; callers choose the dtype of the inputs (fp64 for the
"truth", fp32 to mimic the reference's fp32 run).  The multi-rank functions emulate W ranks
inside ONE process: every rank's view of the gathered tensors is rebuilt with exactly the
grad-carrying / detached chunks the reference produces, and the autograd-aware all_gather's
backward (reduce-scatter SUM, torch.distributed.nn.functional._AllGather.backward) is
restated as "sum over ranks of d loss_r / d chunk_k".
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F


def l2_normalize(x: torch.Tensor, eps: float = 1e-12) -> torch.Tensor:
    """x / max(||x||_2, eps) along the last dim (model.py:313 -> torch F.normalize)."""
    # same decomposition as torch.nn.functional.normalize: norm -> clamp_min(eps) -> divide.
    # (vector_norm has a zero sub-gradient at x == 0, so a zero row gets gx = gy / eps.)
    n = torch.linalg.vector_norm(x, ord=2, dim=-1, keepdim=True).clamp_min(eps)
    return x / n


def clip_logits(image_features, text_features, logit_scale):
    """W=1 branch of ClipLoss.get_logits (loss.py:122-124).

    Note the precedence the reference has: ``logit_scale * image_features @ text_features.T``
    is ``(s * I) @ T.T`` -- the scale is applied to the features first.
    """
    logits_per_image = (logit_scale * image_features) @ text_features.T
    logits_per_text = (logit_scale * text_features) @ image_features.T
    return logits_per_image, logits_per_text


def _ce_mean(logits: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
    """F.cross_entropy(reduction='mean') written out: mean_i( LSE_j logits_ij - logits_i,label_i )."""
    lse = torch.logsumexp(logits, dim=1)
    picked = logits.gather(1, labels[:, None])[:, 0]
    return (lse - picked).mean()


def clip_loss_single(image_features, text_features, logit_scale) -> torch.Tensor:
    """ClipLoss(world_size=1).forward (loss.py:128-139)."""
    li, lt = clip_logits(image_features, text_features, logit_scale)
    labels = torch.arange(li.shape[0], dtype=torch.long)
    return (_ce_mean(li, labels) + _ce_mean(lt, labels)) / 2


def clip_loss_closed_form(image_features, text_features, logit_scale):
    """Closed-form loss AND gradients of the W=1 loss, no autograd.

    S = s I T^T ; loss = 1/2 [ mean_i(LSE_j S_ij - S_ii) + mean_j(LSE_i S_ij - S_jj) ]
    dS = (softmax_row(S) + softmax_col(S) - 2*eye) / (2B)
    dI = s dS T ; dT = s dS^T I ; ds = sum(dS * S) / s
    Used to cross-check the autograd oracle and as the statement the CUDA kernels implement.
    """
    I, T, s = image_features, text_features, logit_scale
    B = I.shape[0]
    S = s * (I @ T.T)
    row_lse = torch.logsumexp(S, dim=1)
    col_lse = torch.logsumexp(S, dim=0)
    diag = torch.diagonal(S)
    loss = 0.5 * ((row_lse - diag).mean() + (col_lse - diag).mean())
    dS = (torch.exp(S - row_lse[:, None]) + torch.exp(S - col_lse[None, :])
          - 2 * torch.eye(B, dtype=S.dtype)) / (2 * B)
    dI = s * (dS @ T)
    dT = s * (dS.T @ I)
    ds = (dS * S).sum() / s
    return loss, dI, dT, ds, row_lse, col_lse


def _rank_loss(r: int, I_chunks: Sequence[torch.Tensor], T_chunks: Sequence[torch.Tensor],
               scale: torch.Tensor, local_loss: bool) -> torch.Tensor:
    """loss on rank r given that rank's view of the gathered chunks (loss.py:104-139)."""
    W = len(I_chunks)
    all_I = torch.cat(list(I_chunks), dim=0)      # rank order (loss.py:51-52, 62-63)
    all_T = torch.cat(list(T_chunks), dim=0)
    if W > 1:
        if local_loss:
            li = (scale * I_chunks[r]) @ all_T.T   # loss.py:117
            lt = (scale * T_chunks[r]) @ all_I.T   # loss.py:118
        else:
            li = (scale * all_I) @ all_T.T         # loss.py:120
            lt = li.T                              # loss.py:121
    else:
        li = (scale * I_chunks[0]) @ T_chunks[0].T
        lt = (scale * T_chunks[0]) @ I_chunks[0].T
    n = li.shape[0]
    labels = torch.arange(n, dtype=torch.long)
    if W > 1 and local_loss:
        labels = labels + n * r                    # loss.py:95-96
    return (_ce_mean(li, labels) + _ce_mean(lt, labels)) / 2


def clip_loss_world(
    I_list: Sequence[torch.Tensor],
    T_list: Sequence[torch.Tensor],
    logit_scale: float | torch.Tensor,
    local_loss: bool,
    gather_with_grad: bool,
    grad_outputs: Optional[Sequence[float]] = None,
) -> Tuple[List[torch.Tensor], List[torch.Tensor], List[torch.Tensor], List[torch.Tensor]]:
    """Emulates ``loss_r = ClipLoss(local_loss, gather_with_grad, rank=r, world_size=W)(I_r, T_r, s)``
    followed by ``loss_r.backward()`` on every rank r, in one process.

    Returns (losses[r], dI[r], dT[r], dscale[r]): what rank r would hold in
    ``I_r.grad``, ``T_r.grad`` and ``logit_scale.grad`` (before DDP's 1/W averaging).

    Which chunks carry gradient on rank r (loss.py:21-65):
      gather_with_grad=True   every chunk k (torch.distributed.nn.all_gather, :51-52); its
                              backward reduce-scatters: rank k receives sum_r dloss_r/dchunk_k.
                              With local_loss the local operand of the matmul is the *direct*
                              local tensor (:117-118) and ALSO appears inside the gathered cat.
      gather_with_grad=False  dist.all_gather output is constant; if not local_loss the local
                              chunk is re-inserted (:58-61); if local_loss nothing is
                              re-inserted and only the direct local operand carries grad.
    """
    W = len(I_list)
    dt = I_list[0].dtype
    go = [1.0] * W if grad_outputs is None else list(grad_outputs)
    losses, dI, dT, dS = [], [torch.zeros_like(x) for x in I_list], \
        [torch.zeros_like(x) for x in T_list], []
    for r in range(W):
        s = torch.as_tensor(logit_scale, dtype=dt).clone().requires_grad_(True)
        if W == 1:
            Ic = [I_list[0].clone().requires_grad_(True)]
            Tc = [T_list[0].clone().requires_grad_(True)]
            loss = _rank_loss(0, Ic, Tc, s, local_loss)
            loss.backward(torch.as_tensor(go[0], dtype=dt))
            return [loss.detach()], [Ic[0].grad], [Tc[0].grad], [s.grad]
        # leaves as seen from rank r
        leaf_I = [x.clone().requires_grad_(True) for x in I_list]
        leaf_T = [x.clone().requires_grad_(True) for x in T_list]
        if gather_with_grad:
            Ic, Tc = list(leaf_I), list(leaf_T)
        else:
            Ic = [x.detach() for x in leaf_I]
            Tc = [x.detach() for x in leaf_T]
            if not local_loss:
                Ic[r], Tc[r] = leaf_I[r], leaf_T[r]
        if local_loss:
            # direct local operands (loss.py:117-118) always carry grad
            all_I = torch.cat(Ic, dim=0)
            all_T = torch.cat(Tc, dim=0)
            li = (s * leaf_I[r]) @ all_T.T
            lt = (s * leaf_T[r]) @ all_I.T
            n = li.shape[0]
            labels = torch.arange(n, dtype=torch.long) + n * r
            loss = (_ce_mean(li, labels) + _ce_mean(lt, labels)) / 2
        else:
            loss = _rank_loss(r, Ic, Tc, s, local_loss)
        loss.backward(torch.as_tensor(go[r], dtype=dt))
        losses.append(loss.detach())
        dS.append(s.grad)
        for k in range(W):
            if leaf_I[k].grad is not None:
                # gather_with_grad: reduce-scatter SUM onto the owner k;
                # otherwise only k == r ever has a grad.
                dI[k] = dI[k] + leaf_I[k].grad
            if leaf_T[k].grad is not None:
                dT[k] = dT[k] + leaf_T[k].grad
    return losses, dI, dT, dS


def clip_loss_local_rank(I_loc, T_loc, all_I, all_T, logit_scale, rank: int) -> torch.Tensor:
    """What ONE rank of a ``local_loss`` world computes (loss.py:116-118 logits, :95-96 labels, :134-137 loss):
    its b rows of both logit matrices against all B gathered columns.  The W rank shards partition the rows of the
    two B x B matrices exactly, so W such steps are the whole-batch work of the reference's algorithm (bench.py times
    this as the bounded CPU sample of config 5)."""
    li = (logit_scale * I_loc) @ all_T.T
    lt = (logit_scale * T_loc) @ all_I.T
    n = li.shape[0]
    labels = torch.arange(n, dtype=torch.long) + n * rank
    return (_ce_mean(li, labels) + _ce_mean(lt, labels)) / 2
