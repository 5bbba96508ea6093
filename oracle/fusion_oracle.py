"""Oracle (test infrastructure): CPU restatement of the TQN fusion head and its loss.

Follows /root/reference:
  * src/open_clip/CAR_heads/TQN_model.py:13-78        TQN_Model (d_model 512, 4 heads, FF 1024, 4 pre-norm decoder
                                                      layers WITHOUT self-attention, shared decoder_norm, MLP head
                                                      512 -> 1024 -> 512 -> 256 -> class_num, ReLU)
  * src/open_clip/CAR_heads/transformer_decoder.py:10-48    TransformerDecoder.forward (layer loop + final norm)
  * src/open_clip/CAR_heads/transformer_decoder.py:146-240  TransformerDecoderWoSelfAttenLayer.forward_pre
        tgt2 = norm2(tgt); tgt2 = MHA(q=tgt2, k=memory, v=memory); tgt += tgt2
        tgt2 = norm3(tgt); tgt2 = linear2(relu(linear1(tgt2))); tgt += tgt2          (dropouts are identity in eval)
  * src/open_clip/model.py:552-561                    how CLIP.forward calls it: memory = [mean token | tokens] of one
                                                      modality, queries = the B global features of the other; the
                                                      [B, B, 1] output is squeezed to a B x B matrix
  * src/open_clip/tagging_heads/asymmetric_loss.py:54-65    DQNCOSLoss: (CE(X, arange) + CE(X^T, arange)) / 2
Third-party arithmetic: torch.nn.MultiheadAttention (packed in_proj, softmax(q k^T / sqrt(E/h)) v, out_proj) and
torch.nn.LayerNorm (eps 1e-5) -- restated here, pinned by tests/golden/fusion.npz which the reference produced.
"""
from __future__ import annotations

import math
from typing import Dict

import torch
import torch.nn.functional as F

FUSION_CFG = dict(heads=4, ff=1024, layers=4, mlp=(1024, 512, 256), ln_eps=1e-5, d_model=512, class_num=1)


def make_fusion_params(seed: int, d_model: int = 512, class_num: int = 1, layers: int = 4,
                       dtype: torch.dtype = torch.float64) -> Dict[str, torch.Tensor]:
    """Deterministic TQN_Model parameters keyed by the reference's state_dict names (including the self_attn / norm1
    tensors every layer owns but never uses, and `decoder.norm.*`, which is the same module as `decoder_norm.*`).
    Drawn in fp32 in a fixed key order from one generator so tests and make_golden.py rebuild identical tensors."""
    g = torch.Generator().manual_seed(seed)
    FF = FUSION_CFG["ff"]

    def w(*shape, std=0.05):
        return (torch.randn(*shape, generator=g, dtype=torch.float32) * std).to(dtype)

    def ln_w(n):
        return (1.0 + torch.randn(n, generator=g, dtype=torch.float32) * 0.05).to(dtype)

    E = d_model
    p: Dict[str, torch.Tensor] = {}
    p["logit_scale"] = torch.tensor(math.log(1 / 0.07), dtype=dtype)
    p["decoder_norm.weight"] = ln_w(E)
    p["decoder_norm.bias"] = w(E, std=0.02)
    for l in range(layers):
        k = f"decoder.layers.{l}."
        for attn in ("self_attn", "multihead_attn"):
            p[k + attn + ".in_proj_weight"] = w(3 * E, E)
            p[k + attn + ".in_proj_bias"] = w(3 * E, std=0.02)
            p[k + attn + ".out_proj.weight"] = w(E, E)
            p[k + attn + ".out_proj.bias"] = w(E, std=0.02)
        p[k + "linear1.weight"] = w(FF, E)
        p[k + "linear1.bias"] = w(FF, std=0.02)
        p[k + "linear2.weight"] = w(E, FF)
        p[k + "linear2.bias"] = w(E, std=0.02)
        for n in ("norm1", "norm2", "norm3"):
            p[k + n + ".weight"] = ln_w(E)
            p[k + n + ".bias"] = w(E, std=0.02)
    p["decoder.norm.weight"] = p["decoder_norm.weight"]
    p["decoder.norm.bias"] = p["decoder_norm.bias"]
    dims = (E,) + FUSION_CFG["mlp"] + (class_num,)
    for i, idx in enumerate((0, 3, 6, 9)):
        p[f"mlp_head.{idx}.weight"] = w(dims[i + 1], dims[i])
        p[f"mlp_head.{idx}.bias"] = w(dims[i + 1], std=0.02)
    return p


def _ln(x, w, b):
    return F.layer_norm(x, (x.shape[-1],), w, b, FUSION_CFG["ln_eps"])


def fusion_forward(memory_tokens: torch.Tensor, query_features: torch.Tensor, params: Dict[str, torch.Tensor],
                   layers: int = 4) -> torch.Tensor:
    """TQN_Model.forward(image_features=memory_tokens [B, P, E], text_features=query_features [Q, E]) in eval mode
    -> [B, Q, class_num].  Batch-first restatement of the reference's sequence-first computation."""
    B, P, E = memory_tokens.shape
    Q = query_features.shape[0]
    H = FUSION_CFG["heads"]
    dh = E // H
    nw, nb = params["decoder_norm.weight"], params["decoder_norm.bias"]
    mem = _ln(memory_tokens, nw, nb)                                   # TQN_model.py:69
    tgt = _ln(query_features, nw, nb).unsqueeze(0).expand(B, Q, E)     # :67-70 (repeat over the batch, then norm)
    for l in range(layers):
        k = f"decoder.layers.{l}."
        Wi, bi = params[k + "multihead_attn.in_proj_weight"], params[k + "multihead_attn.in_proj_bias"]
        Wo, bo = params[k + "multihead_attn.out_proj.weight"], params[k + "multihead_attn.out_proj.bias"]
        t2 = _ln(tgt, params[k + "norm2.weight"], params[k + "norm2.bias"])
        q = F.linear(t2, Wi[:E], bi[:E]).reshape(B, Q, H, dh).permute(0, 2, 1, 3)
        kk = F.linear(mem, Wi[E:2 * E], bi[E:2 * E]).reshape(B, P, H, dh).permute(0, 2, 1, 3)
        v = F.linear(mem, Wi[2 * E:], bi[2 * E:]).reshape(B, P, H, dh).permute(0, 2, 1, 3)
        att = torch.softmax(q @ kk.transpose(-1, -2) / math.sqrt(dh), dim=-1)
        ctx = (att @ v).permute(0, 2, 1, 3).reshape(B, Q, E)
        tgt = tgt + F.linear(ctx, Wo, bo)
        t2 = _ln(tgt, params[k + "norm3.weight"], params[k + "norm3.bias"])
        t2 = F.linear(F.relu(F.linear(t2, params[k + "linear1.weight"], params[k + "linear1.bias"])),
                      params[k + "linear2.weight"], params[k + "linear2.bias"])
        tgt = tgt + t2
    out = _ln(tgt, nw, nb)                                             # transformer_decoder.py:38-39 (decoder.norm)
    for idx in (0, 3, 6):
        out = F.relu(F.linear(out, params[f"mlp_head.{idx}.weight"], params[f"mlp_head.{idx}.bias"]))
    return F.linear(out, params["mlp_head.9.weight"], params["mlp_head.9.bias"])


def fusion_scores(tokens: torch.Tensor, other_tokens: torch.Tensor, params, layers: int = 4) -> torch.Tensor:
    """model.py:552-561: memory = [mean token | tokens] of one modality, queries = per-sample mean token of the other
    -> the B x B matrix `i2t_cls` (tokens = ViT tokens, other = text tokens) or `t2i_cls` (swapped)."""
    mem = torch.cat([tokens.mean(dim=1, keepdim=True), tokens], dim=1)
    return fusion_forward(mem, other_tokens.mean(dim=1), params, layers).squeeze(-1)


def dqn_cos_loss(x: torch.Tensor) -> torch.Tensor:
    """asymmetric_loss.py:54-65 with the two cross-entropies written out."""
    n = x.shape[0]
    d = x.diagonal()
    return 0.5 * ((torch.logsumexp(x, 1) - d).mean() + (torch.logsumexp(x, 0) - d).mean())
