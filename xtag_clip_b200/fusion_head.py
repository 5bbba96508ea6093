"""Drop-in for the TQN fusion head and its loss -- the second cross-attention on the path when the reference runs
with `--use-fusion`.  HOST COMPOSITION over kernels that already exist:

  * `FusionHead` mirrors /root/reference/src/open_clip/CAR_heads/TQN_model.py:13-78 (`TQN_Model`) and the decoder of
    CAR_heads/transformer_decoder.py:10-48, 146-240 with the reference's parameter names, so a reference
    `fusion_model.state_dict()` loads unchanged (the per-layer `self_attn` / `norm1` tensors the reference owns but never
    uses are kept as inert parameters for that reason).  The attention core runs on K4 (`xtag_xattn_fwd/bwd`, flash-style,
    no [B, heads, Q, P] score tensor in HBM) in chunks of <= 64 queries; projections / LayerNorm / MLP are torch
    library calls, as in the tag head.  Layer 0's query projection does not depend on the sample (every sample
    attends with the same B query vectors) and is computed once on [Q, E].
  * `fusion_scores` is model.py:552-561: memory = [mean token | tokens] of one modality, queries = the per-sample mean
    tokens of the other, output squeezed to the B x B matrix `i2t_cls` / `t2i_cls`.
  * `DQNCOSLoss` mirrors tagging_heads/asymmetric_loss.py:54-65: (CE(X, arange) + CE(X^T, arange)) / 2 on a
    materialised B x B matrix, on its own kernels (`csrc/symm_ce.cu`): the forward reads X once (row LSEs per 32-row
    strip, column partials finished by the shared LSE reduction), the backward is one pass that writes the closed form
    (softmax_row + softmax_col - 2 I) / (2B).

Status: host logic and parity are covered on CPU against fixtures the reference produced (tests/golden/fusion.npz);
the GPU tests of this module (tests/test_fusion_head.py, `-m gpu`) pass on a B200.
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from .tag_head import cross_attention

Q_CHUNK = 64          # the exact fp32 K4 kernels hold the whole [Lq, Lk] score block of a (sample, head) in shared memory
_LOG2E = 1.4426950408889634


def _kernels(k=None):
    if k is not None:
        return k
    from .kernels import default_kernels
    return default_kernels()


class _DecoderLayer(nn.Module):
    """TransformerDecoderWoSelfAttenLayer (transformer_decoder.py:146-166): parameter holder."""

    def __init__(self, d_model: int, nhead: int, dim_feedforward: int, dropout: float):
        super().__init__()
        self.self_attn = nn.MultiheadAttention(d_model, nhead, dropout=dropout)          # never used by the reference
        self.multihead_attn = nn.MultiheadAttention(d_model, nhead, dropout=dropout)
        self.linear1 = nn.Linear(d_model, dim_feedforward)
        self.linear2 = nn.Linear(dim_feedforward, d_model)
        self.norm1 = nn.LayerNorm(d_model)                                               # never used by the reference
        self.norm2 = nn.LayerNorm(d_model)
        self.norm3 = nn.LayerNorm(d_model)


class _Decoder(nn.Module):
    def __init__(self, d_model, nhead, dim_feedforward, dropout, num_layers, norm):
        super().__init__()
        self.layers = nn.ModuleList([_DecoderLayer(d_model, nhead, dim_feedforward, dropout) for _ in range(num_layers)])
        self.norm = norm                      # the SAME module as FusionHead.decoder_norm (TQN_model.py:28-30)


class FusionHead(nn.Module):
    """`TQN_Model(cfg=None)` -> `FusionHead()`; `forward(image_features [B, P, E], text_features [Q, E]) -> [B, Q,
    class_num]` (TQN_model.py:62-78)."""

    def __init__(self, d_model: int = 512, class_num: int = 1, num_layers: int = 4, nhead: int = 4,
                 dim_feedforward: int = 1024, dropout: float = 0.1, *, _kernels=None):
        super().__init__()
        self.d_model, self.nhead, self.p_drop = d_model, nhead, dropout
        self.logit_scale = nn.Parameter(torch.ones([]) * math.log(1 / 0.07))
        self.decoder_norm = nn.LayerNorm(d_model)
        self.decoder = _Decoder(d_model, nhead, dim_feedforward, dropout, num_layers, self.decoder_norm)
        self.dropout_feas = nn.Dropout(dropout)
        self.mlp_head = nn.Sequential(
            nn.Linear(d_model, 1024), nn.ReLU(inplace=True), nn.Dropout(dropout),
            nn.Linear(1024, 512), nn.ReLU(inplace=True), nn.Dropout(dropout),
            nn.Linear(512, 256), nn.ReLU(inplace=True), nn.Dropout(dropout),
            nn.Linear(256, class_num))
        self._k = _kernels
        self._step = 0
        self.apply(self._init_weights)

    @staticmethod
    def _init_weights(module):                                  # TQN_model.py:48-60
        if isinstance(module, nn.Linear):
            module.weight.data.normal_(mean=0.0, std=0.02)
        elif isinstance(module, nn.MultiheadAttention):
            module.in_proj_weight.data.normal_(mean=0.0, std=0.02)
            module.out_proj.weight.data.normal_(mean=0.0, std=0.02)

    @classmethod
    def from_reference(cls, fusion_model: nn.Module, **kw) -> "FusionHead":
        sd = fusion_model.state_dict()
        d = sd["decoder_norm.weight"].shape[0]
        layers = 1 + max(int(k.split(".")[2]) for k in sd if k.startswith("decoder.layers."))
        head = cls(d, sd["mlp_head.9.weight"].shape[0], layers, **kw)
        head.load_state_dict(sd, strict=True)
        return head.to(device=sd["decoder_norm.weight"].device, dtype=sd["decoder_norm.weight"].dtype)

    def _attend(self, q, kv, seed, offset):
        """q [B, Q, E], kv [B, P, 2E] (K | V) -> ctx [B, Q, E]; queries in chunks of Q_CHUNK (the chunks share K/V)."""
        E = self.d_model
        k, v = kv[..., :E], kv[..., E:]
        drop = self.p_drop if self.training else 0.0
        if q.is_cuda and q.dtype == torch.bfloat16 and (E // self.nhead) in (64, 128, 192, 256):
            # tensor-core K4 takes a query set of any length in ONE launch (64-row chunks on the grid; the backward is
            # the chunked dQ kernel + the dK / dV kernel that loops over the query chunks: no per-chunk partial
            # gradients to add up)
            return cross_attention(q, k, v, self.nhead, drop, seed, offset * 64, _kernels=self._k)
        out = []
        for ci, c in enumerate(range(0, q.shape[1], Q_CHUNK)):
            out.append(cross_attention(q[:, c:c + Q_CHUNK], k, v, self.nhead, drop, seed, offset * 64 + ci,
                                       _kernels=self._k))
        return out[0] if len(out) == 1 else torch.cat(out, dim=1)

    def forward(self, image_features, text_features, pos=None, return_atten=False, inside_repeat=True,
                seed: Optional[int] = None):
        if return_atten:
            raise NotImplementedError("FusionHead never materialises the attention map (return_atten=True)")
        if pos is not None:
            raise NotImplementedError("positional terms are not used on this path (model.py:559-560 passes none)")
        if image_features.dim() != 3 or image_features.shape[-1] != self.d_model:
            raise ValueError(f"image_features must be [B, P, {self.d_model}], got {tuple(image_features.shape)}")
        B, P, E = image_features.shape
        train = self.training
        if seed is None:
            seed = int(torch.initial_seed() & 0x7FFFFFFFFFFFFFFF)
        self._step += 1
        mem = self.decoder_norm(image_features)                               # TQN_model.py:69
        if inside_repeat:
            tq = self.decoder_norm(text_features)                             # [Q, E], identical for every sample
            tgt = None
        else:                                                                 # caller passed [Q, B, E] (seq-first)
            tq = None
            tgt = self.decoder_norm(text_features).transpose(0, 1)
        for li, layer in enumerate(self.decoder.layers):
            mha = layer.multihead_attn
            Wi, bi = mha.in_proj_weight, mha.in_proj_bias
            if tgt is None:
                q = F.linear(layer.norm2(tq), Wi[:E], bi[:E]).unsqueeze(0).expand(B, -1, -1)   # once, not per sample
                resid = tq.unsqueeze(0)
            else:
                q = F.linear(layer.norm2(tgt), Wi[:E], bi[:E])
                resid = tgt
            kv = F.linear(mem, Wi[E:], bi[E:])                                # fused K | V projection, [B, P, 2E]
            ctx = self._attend(q, kv, seed, self._step * len(self.decoder.layers) + li)
            a = F.dropout(mha.out_proj(ctx.to(kv.dtype)), self.p_drop, train)
            tgt = resid + a
            f = layer.linear2(F.dropout(F.relu(layer.linear1(layer.norm3(tgt))), self.p_drop, train))
            tgt = tgt + F.dropout(f, self.p_drop, train)
        out = self.dropout_feas(self.decoder.norm(tgt))                       # [B, Q, E]
        return self.mlp_head(out)


def fusion_scores(head: FusionHead, tokens: torch.Tensor, other_tokens: torch.Tensor) -> torch.Tensor:
    """model.py:552-561 for one direction: `fusion_scores(head, out_token, text_tokens)` is `i2t_cls`,
    `fusion_scores(head, text_tokens, out_token)` is `t2i_cls` (both [B, B])."""
    mem = torch.cat([tokens.mean(dim=1, keepdim=True), tokens], dim=1)
    return head(mem, other_tokens.mean(dim=1)).squeeze(-1)


class _SymmetricCE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, kernels):
        K = _kernels(kernels)
        ctx.fused = hasattr(K, "symm_ce_fwd") and x.dtype in (torch.float32, torch.bfloat16)
        if ctx.fused:
            # one pass over X forward (xtag_symm_ce_fwd), one pass backward (xtag_symm_ce_bwd)
            xc = x.detach() if x.stride(-1) == 1 else x.detach().contiguous()
            loss, row_lse, col_lse = K.symm_ce_fwd(xc)
            ctx.K = K
            ctx.save_for_backward(xc, row_lse, col_lse)
            return loss
        xs = x.detach().float() * _LOG2E                                      # log2-domain copy, [B, B]
        col_lse = K.lse_reduce_log2(xs)                                       # ln sum_i exp(x_ij)
        row_lse = K.lse_reduce_log2(xs.t().contiguous())                      # ln sum_j exp(x_ij)
        d = x.detach().float().diagonal()
        ctx.save_for_backward(x, row_lse, col_lse)
        return 0.5 * ((row_lse - d).mean() + (col_lse - d).mean())

    @staticmethod
    def backward(ctx, g):
        x, row_lse, col_lse = ctx.saved_tensors
        if ctx.fused:
            return ctx.K.symm_ce_bwd(x, row_lse, col_lse, g), None
        n = x.shape[0]
        xf = x.float()
        dx = torch.exp(xf - row_lse[:, None]) + torch.exp(xf - col_lse[None, :])
        dx.diagonal().sub_(2.0)
        return (dx * (g.float() / (2.0 * n))).to(x.dtype), None


class DQNCOSLoss(nn.Module):
    """`DQNCOSLoss()(input [B, B]) -> 0-d` (asymmetric_loss.py:54-65)."""

    def __init__(self, *, _kernels=None):
        super().__init__()
        self._k = _kernels

    def forward(self, input):
        if input.dim() != 2 or input.shape[0] != input.shape[1]:
            raise ValueError(f"DQNCOSLoss expects a square [B, B] matrix, got {tuple(input.shape)}")
        return _SymmetricCE.apply(input, self._k)
