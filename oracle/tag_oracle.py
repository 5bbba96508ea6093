"""Oracle (test infrastructure): CPU restatement of the XTag cross-attention tag head.

Follows /root/reference:
  * src/open_clip/model.py:270-288      tag head construction (2-layer BERT, self-attention and
                                        embeddings deleted, tag_labels Embedding(44,768), tag_fc)
  * src/open_clip/model.py:337-352      tag_forward
  * src/open_clip/model.py:354-383      prepare_control_words (index part only)
  * src/open_clip/tagging_heads/bert.py:189-278   BertSelfAttention.forward (cross branch)
  * src/open_clip/tagging_heads/bert.py:281-292   BertSelfOutput  (dense + residual + LayerNorm)
  * src/open_clip/tagging_heads/bert.py:344-370   BertIntermediate (erf GELU), BertOutput
  * src/open_clip/tagging_heads/bert.py:386-456   BertLayer.forward(mode='tagging')
  * src/open_clip/tagging_heads/bert.py:743-880   BertModel.forward (all-ones encoder mask -> +0)
  * src/open_clip/tagging_heads/asymmetric_loss.py:16-50   AsymmetricLoss.forward
  * src/open_clip/tagging_heads/tag_bert_config.json      hidden 768, 4 heads, 2 layers, FF 3072,
                                                          LN eps 1e-12, dropout 0.1 (eval: off)

Third-party arithmetic on this path: HF ``transformers`` (version unpinned by the reference;
5.5.0 installed) contributes only ``ACT2FN['gelu']`` (= exact erf GELU) and
``invert_attention_mask`` ((1 - mask) * finfo.min -> 0 for the all-ones mask used here).
"""
from __future__ import annotations

import math
from typing import Dict

import torch
import torch.nn.functional as F

TAG_CFG = dict(hidden=768, heads=4, layers=2, intermediate=3072, ln_eps=1e-12,
               num_tags=22, num_queries=44, attn_dropout=0.1, hidden_dropout=0.1,
               category_sizes=(3, 4, 3, 4, 4, 4))


def _layer_keys(l: int):
    p = f"tag_head.encoder.layer.{l}."
    return p


def make_tag_params(seed: int, embed_dim: int, gain: float = 1.0,
                    dtype: torch.dtype = torch.float64) -> Dict[str, torch.Tensor]:
    """Deterministic tag-head parameters, keyed by the reference's state_dict names
    (``tag_head.encoder.layer.{0,1}.crossattention.self.{query,key,value}.*`` ...).

    Weights ~ N(0, (gain*0.02)^2) like ``BertPreTrainedModel._init_weights`` (bert.py:631-641,
    initializer_range 0.02); biases and LayerNorm affine terms are perturbed too so that every
    term of the forward is exercised.  Values are drawn in fp32 in a fixed key order from one
    ``torch.Generator`` so that tests and ``make_golden.py`` rebuild identical tensors.
    """
    H, FF, Q = TAG_CFG["hidden"], TAG_CFG["intermediate"], TAG_CFG["num_queries"]
    g = torch.Generator().manual_seed(seed)
    std = 0.02 * gain

    def w(*shape):
        return (torch.randn(*shape, generator=g, dtype=torch.float32) * std).to(dtype)

    def b(n):
        return (torch.randn(n, generator=g, dtype=torch.float32) * 0.02).to(dtype)

    def ln_w(n):
        return (1.0 + torch.randn(n, generator=g, dtype=torch.float32) * 0.05).to(dtype)

    p: Dict[str, torch.Tensor] = {}
    p["tag_labels.weight"] = w(Q, H)
    for l in range(TAG_CFG["layers"]):
        k = _layer_keys(l)
        p[k + "crossattention.self.query.weight"] = w(H, H)
        p[k + "crossattention.self.query.bias"] = b(H)
        p[k + "crossattention.self.key.weight"] = w(H, embed_dim)
        p[k + "crossattention.self.key.bias"] = b(H)
        p[k + "crossattention.self.value.weight"] = w(H, embed_dim)
        p[k + "crossattention.self.value.bias"] = b(H)
        p[k + "crossattention.output.dense.weight"] = w(H, H)
        p[k + "crossattention.output.dense.bias"] = b(H)
        p[k + "crossattention.output.LayerNorm.weight"] = ln_w(H)
        p[k + "crossattention.output.LayerNorm.bias"] = b(H)
        p[k + "intermediate.dense.weight"] = w(FF, H)
        p[k + "intermediate.dense.bias"] = b(FF)
        p[k + "output.dense.weight"] = w(H, FF)
        p[k + "output.dense.bias"] = b(H)
        p[k + "output.LayerNorm.weight"] = ln_w(H)
        p[k + "output.LayerNorm.bias"] = b(H)
    p["tag_fc.weight"] = w(1, H)
    p["tag_fc.bias"] = b(1)
    return p


def cross_attention_core(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, heads: int) -> torch.Tensor:
    """softmax(q k^T / sqrt(dh) + 0) v per (sample, head)  (bert.py:219-274, eval mode).

    q [b, Lq, H], k/v [b, Lk, H] with H = heads*dh, head h occupying columns [h*dh, (h+1)*dh)
    (``transpose_for_scores``, bert.py:184-187).  Returns the merged context [b, Lq, H].
    """
    b, Lq, Hd = q.shape
    Lk = k.shape[1]
    dh = Hd // heads
    qh = q.reshape(b, Lq, heads, dh).permute(0, 2, 1, 3)
    kh = k.reshape(b, Lk, heads, dh).permute(0, 2, 1, 3)
    vh = v.reshape(b, Lk, heads, dh).permute(0, 2, 1, 3)
    scores = qh @ kh.transpose(-1, -2) / math.sqrt(dh)
    probs = torch.softmax(scores, dim=-1)
    ctx = probs @ vh
    return ctx.permute(0, 2, 1, 3).reshape(b, Lq, Hd)


def tag_head_forward(tokens: torch.Tensor, p: Dict[str, torch.Tensor]) -> torch.Tensor:
    """CLIP.tag_forward (model.py:337-352) in eval mode: tokens [b, N, D] -> tag_logits [b, 44]."""
    H, heads, eps = TAG_CFG["hidden"], TAG_CFG["heads"], TAG_CFG["ln_eps"]
    bsz = tokens.shape[0]
    h = p["tag_labels.weight"].unsqueeze(0).repeat(bsz, 1, 1)          # model.py:342
    for l in range(TAG_CFG["layers"]):
        k_ = _layer_keys(l)
        ca = k_ + "crossattention."
        q = F.linear(h, p[ca + "self.query.weight"], p[ca + "self.query.bias"])        # bert.py:199
        k = F.linear(tokens, p[ca + "self.key.weight"], p[ca + "self.key.bias"])       # bert.py:208
        v = F.linear(tokens, p[ca + "self.value.weight"], p[ca + "self.value.bias"])   # bert.py:209
        ctx = cross_attention_core(q, k, v, heads)                                      # bert.py:231-274
        a = F.linear(ctx, p[ca + "output.dense.weight"], p[ca + "output.dense.bias"])  # bert.py:289
        a = F.layer_norm(a + h, (H,), p[ca + "output.LayerNorm.weight"],
                         p[ca + "output.LayerNorm.bias"], eps)                          # bert.py:291
        f = F.linear(a, p[k_ + "intermediate.dense.weight"], p[k_ + "intermediate.dense.bias"])
        f = F.gelu(f)                                                                   # erf GELU, bert.py:355
        o = F.linear(f, p[k_ + "output.dense.weight"], p[k_ + "output.dense.bias"])    # bert.py:367
        h = F.layer_norm(o + a, (H,), p[k_ + "output.LayerNorm.weight"],
                         p[k_ + "output.LayerNorm.bias"], eps)                          # bert.py:369
    return F.linear(h, p["tag_fc.weight"], p["tag_fc.bias"]).squeeze(-1)                # model.py:351


def asymmetric_loss(x: torch.Tensor, y: torch.Tensor, gamma_neg: float = 4.0, gamma_pos: float = 1.0,
                    clip: float = 0.05, eps: float = 1e-8) -> torch.Tensor:
    """AsymmetricLoss.forward (asymmetric_loss.py:16-50); the focal weight carries no grad
    (the reference disables grad globally around it, :41-48).  Returns -sum (not mean)."""
    p = torch.sigmoid(x)
    p_neg = 1 - p
    if clip is not None and clip > 0:
        p_neg = (p_neg + clip).clamp(max=1)
    loss = y * torch.log(p.clamp(min=eps)) + (1 - y) * torch.log(p_neg.clamp(min=eps))
    if gamma_neg > 0 or gamma_pos > 0:
        with torch.no_grad():
            pt = p * y + p_neg * (1 - y)
            w = torch.pow(1 - pt, gamma_pos * y + gamma_neg * (1 - y))
        loss = loss * w
    return -loss.sum()


def control_word_indices(tag_logits: torch.Tensor) -> torch.Tensor:
    """Index part of CLIP.prepare_control_words (model.py:354-374): per category (sizes
    [3,4,3,4,4,4] over the 22 tags) the arg-max of sigmoid(l[:, j]) + sigmoid(l[:, 22+j]).
    Returns int64 [b, 6] indices into the 22-entry tag list."""
    n = TAG_CFG["num_tags"]
    s = torch.sigmoid(tag_logits)
    out, pos = [], 0
    for size in TAG_CFG["category_sizes"]:
        score = s[:, pos:pos + size] + s[:, n + pos:n + pos + size]
        out.append(torch.argsort(score, dim=-1, descending=True)[:, :1] + pos)
        pos += size
    return torch.cat(out, dim=-1)
