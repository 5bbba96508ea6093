"""Drop-in for the XTag cross-attention tag head (``CLIP.tag_forward``) and the feature epilogue.

Mirrors, with the same parameter names so a reference ``state_dict`` loads unchanged
(35 ``tag_*`` keys):
  * /root/reference/src/open_clip/model.py:270-288   construction (BertModel minus embeddings and
    self-attention, ``tag_labels`` Embedding(44, 768), ``tag_fc`` Linear(768, 1))
  * model.py:337-352                                   ``tag_forward(tag_embeds [b,N,D]) -> [b,44]``
  * tagging_heads/bert.py:189-278, 281-292, 344-370, 386-456   the tagging-mode layer
  * model.py:311-313, 332-333                          ``F.normalize`` epilogue -> ``l2_normalize``
  * model.py:354-383                                   ``prepare_control_words``

What runs where: the attention core (QK^T, softmax, dropout, PV) is the fused CUDA kernel K4; the
L2-normalise is K3; the dense projections / LayerNorm / GELU stay torch library calls in v1
("stays cuBLAS / torch in v1").  Layer-0 queries do not depend on the sample
(model.py:342 repeats the same ``tag_labels.weight``), so Q-projection of layer 0 is computed once
per call on [44, 768] instead of [b, 44, 768].
"""
from __future__ import annotations

import math
from typing import List, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

TAG_HIDDEN = 768
TAG_HEADS = 4
TAG_LAYERS = 2
TAG_FF = 3072
TAG_LN_EPS = 1e-12
TAG_QUERIES = 44
TAG_DROPOUT = 0.1
CATEGORY_SIZES = (3, 4, 3, 4, 4, 4)


def _kernels(k=None):
    if k is not None:
        return k
    from .kernels import default_kernels
    return default_kernels()


# ---- K3 -------------------------------------------------------------------------------------------
class _L2Normalize(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, eps, out_dtype, kernels, out):
        K = _kernels(kernels)
        if out is not None:
            y, inv, _ = K.l2norm_fwd(x, out.dtype, eps, out=out)
            ctx.mark_dirty(out)
            y = out
        else:
            y, inv, _ = K.l2norm_fwd(x, out_dtype or x.dtype, eps)
        ctx.save_for_backward(y, inv)
        ctx.meta = (K, eps, x.dtype)
        return y

    @staticmethod
    def backward(ctx, gy):
        y, inv = ctx.saved_tensors
        K, eps, in_dtype = ctx.meta
        return K.l2norm_bwd(gy, y, inv, in_dtype, eps), None, None, None, None


def l2_normalize(x: torch.Tensor, eps: float = 1e-12, out_dtype: Optional[torch.dtype] = None, *,
                 out: Optional[torch.Tensor] = None, _kernels=None):
    """``F.normalize(x, dim=-1)`` (model.py:313, 333) with an optional fused cast of the output.
    ``out``: write the result straight into this [rows, dim] tensor -- an input slot of the contrastive loss
    (``ClipLoss.graph_input_slots``: for a sharded, captured step the text slot IS the rank's block of the
    symmetric-memory gather buffer), so the features are written once, already in bf16, where K1 and the peers read
    them.  The returned tensor is ``out`` with autograd history."""
    # (a fresh alias of the slot: the in-place output must not inherit the autograd history of the previous step)
    return _L2Normalize.apply(x, eps, out_dtype, _kernels, out.detach() if out is not None else None)


# ---- K4 -------------------------------------------------------------------------------------------
class _CrossAttention(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, k, v, heads, dropout_p, seed, offset, kernels, sink):
        K = _kernels(kernels)
        scale = 1.0 / math.sqrt(q.shape[-1] // heads)
        o, lse = K.xattn_fwd(q, k, v, heads, scale, dropout_p, seed, offset)
        ctx.save_for_backward(q, k, v, o, lse)
        ctx.meta = (K, heads, scale, dropout_p, seed, offset, sink)
        return o

    @staticmethod
    def backward(ctx, do):
        q, k, v, o, lse = ctx.saved_tensors
        K, heads, scale, dropout_p, seed, offset, sink = ctx.meta
        dk_out = dv_out = None
        if sink is not None:
            # k and v are column slices of ONE fused projection buffer: their gradients go straight into the matching
            # slices of one gradient buffer, which then feeds a single projection-backward GEMM
            dk_out, dv_out = sink[0].slices(k, sink[1], sink[2])
        if dk_out is not None:
            dq, dk, dv = K.xattn_bwd(q, k, v, o, do, lse, heads, scale, dropout_p, seed, offset, dk_out=dk_out,
                                     dv_out=dv_out)
        else:
            dq, dk, dv = K.xattn_bwd(q, k, v, o, do, lse, heads, scale, dropout_p, seed, offset)
        return dq, dk, dv, None, None, None, None, None, None


def cross_attention(q, k, v, heads: int, dropout_p: float = 0.0, seed: int = 0, offset: int = 0, *, _kernels=None,
                    _sink=None):
    """softmax(q k^T / sqrt(dh)) v per (sample, head) -- bert.py:219-274.  q [b,Lq,H], k/v [b,Lk,H]."""
    if q.dtype == torch.float16:          # fp16 autocast: run the exact fp32 kernel
        q, k, v = q.float(), k.float(), v.float()
    return _CrossAttention.apply(q.contiguous(), k, v, heads, dropout_p, seed, offset, _kernels, _sink)


# ---- fused K|V projection of all layers ------------------------------------------------
class _KVGradBuffer:
    """One [b, N, width] gradient buffer per tag_forward call, allocated by the first attention backward that runs;
    every layer's K4 backward writes its dK / dV into its column slices."""

    def __init__(self, width: int):
        self.width, self.buf, self.written = width, None, set()

    def slices(self, like: torch.Tensor, col_k: int, col_v: int):
        b, n, h = like.shape
        if self.buf is None:
            self.buf = torch.empty((b, n, self.width), dtype=like.dtype, device=like.device)
        self.written.update((col_k, col_v))
        return self.buf[..., col_k:col_k + h], self.buf[..., col_v:col_v + h]


class _FusedKVProjection(torch.autograd.Function):
    """kv [b, N, 2*L*768] = tokens [b, N, D] @ cat(W_k0, W_v0, W_k1, W_v1)^T + cat(biases)  -- the key / value Linear
    layers of every tagging layer (bert.py:208-209) as ONE tcgen05 GEMM with a bias epilogue (`xtag_tc_linear_bf16`);
    the outputs are the per-layer column slices, which K4 reads in place through their row stride.
    Backward: the slices' gradients arrive inside one buffer (`_KVGradBuffer`), so d tokens = dkv @ W and
    dW = dkv^T @ tokens are two GEMMs on the package's own kernels with the operands read in place (MN-major
    descriptors); d bias is a column sum."""

    @staticmethod
    def forward(ctx, tokens, weight, bias, n_out, kernels, gbuf):
        K = _kernels(kernels)
        b, n, d = tokens.shape
        x2d = tokens.reshape(b * n, d)
        wb = weight.to(torch.bfloat16)
        kv = K.tc_linear(x2d, wb, bias).view(b, n, -1)
        h = kv.shape[-1] // n_out
        ctx.save_for_backward(x2d, wb)
        ctx.meta = (K, gbuf, tokens.shape, weight.dtype, bias.dtype, n_out, h)
        return tuple(kv[..., i * h:(i + 1) * h] for i in range(n_out))

    @staticmethod
    def backward(ctx, *grads):
        x2d, wb = ctx.saved_tensors
        K, gbuf, tok_shape, w_dtype, b_dtype, n_out, h = ctx.meta
        b, n, d = tok_shape
        buf = gbuf.buf
        if buf is None:
            buf = torch.empty((b, n, n_out * h), dtype=torch.bfloat16, device=x2d.device)
        for i, g in enumerate(grads):
            sl = buf[..., i * h:(i + 1) * h]
            if g is None:
                sl.zero_()
            elif g.data_ptr() != sl.data_ptr() or g.stride() != sl.stride():
                sl.copy_(g)                 # a gradient that did not come through the shared buffer
        dkv = buf.view(b * n, n_out * h)
        d_tok = d_w = d_b = None
        if ctx.needs_input_grad[0]:
            # d tokens [b*n, D] = dkv [b*n, K'] @ W [K', D]: W is the B operand stored [K'][D] (MN-major)
            d_tok = K.tc_gemm(dkv, wb, False, True, torch.bfloat16).view(b, n, d)
        if ctx.needs_input_grad[1]:
            # dW [K', D] = dkv^T @ tokens: both operands stored [b*n rows][...]: MN-major A and B, fp32 out
            d_w = K.tc_gemm(dkv, x2d, True, True, torch.float32).to(w_dtype)
        if ctx.needs_input_grad[2]:
            d_b = dkv.sum(dim=0, dtype=torch.float32).to(b_dtype)
        return d_tok, d_w, d_b, None, None, None


# ---- K6: dropout + residual + LayerNorm of the dense-output blocks ---------------------------------------------------
class _ResidualDropoutLayerNorm(torch.autograd.Function):
    """y = LayerNorm(dropout(x) + resid) (bert.py:281-292, 359-370) as one pass forward and one pass backward
    (`xtag_ln_res_fwd/bwd`); x [b, L, H] bf16 = the dense layer's output, resid [b, L, H] or [1, L, H]."""

    @staticmethod
    def forward(ctx, x, resid, gamma, beta, eps, p, seed, offset, kernels):
        K = _kernels(kernels)
        H = x.shape[-1]
        y, z, mean, rstd = K.ln_res_fwd(x.reshape(-1, H), resid.reshape(-1, H), gamma, beta, eps, p, seed, offset)
        ctx.save_for_backward(z, mean, rstd, gamma)
        ctx.meta = (K, p, seed, offset, tuple(x.shape), tuple(resid.shape), resid.dtype, gamma.dtype, beta.dtype)
        return y.view(x.shape)

    @staticmethod
    def backward(ctx, dy):
        z, mean, rstd, gamma = ctx.saved_tensors
        K, p, seed, offset, xshape, rshape, rdt, gdt, bdt = ctx.meta
        dx, dres, dg, db = K.ln_res_bwd(dy, z, mean, rstd, gamma, p, seed, offset)
        if rshape[0] != xshape[0]:                  # broadcast residual (layer 0: the shared label embeddings)
            dres = dres.view(xshape).sum(dim=0, keepdim=True, dtype=torch.float32)
        d_res = dres.reshape(rshape).to(rdt) if ctx.needs_input_grad[1] else None
        return dx.view(xshape), d_res, dg.to(gdt), db.to(bdt), None, None, None, None, None


def residual_dropout_layer_norm(x, resid, ln: nn.LayerNorm, p: float, seed: int, offset: int, *, _kernels=None):
    return _ResidualDropoutLayerNorm.apply(x, resid, ln.weight, ln.bias, ln.eps, p, seed, offset, _kernels)


# ---- module tree with the reference's parameter names --------------------------------------------
class _SelfAttn(nn.Module):          # crossattention.self
    def __init__(self, encoder_width: int):
        super().__init__()
        self.query = nn.Linear(TAG_HIDDEN, TAG_HIDDEN)
        self.key = nn.Linear(encoder_width, TAG_HIDDEN)
        self.value = nn.Linear(encoder_width, TAG_HIDDEN)


class _SelfOutput(nn.Module):        # crossattention.output / (layer).output
    def __init__(self, in_features: int):
        super().__init__()
        self.dense = nn.Linear(in_features, TAG_HIDDEN)
        self.LayerNorm = nn.LayerNorm(TAG_HIDDEN, eps=TAG_LN_EPS)


class _CrossAttnBlock(nn.Module):    # crossattention
    def __init__(self, encoder_width: int):
        super().__init__()
        self.self = _SelfAttn(encoder_width)
        self.output = _SelfOutput(TAG_HIDDEN)


class _Intermediate(nn.Module):
    def __init__(self):
        super().__init__()
        self.dense = nn.Linear(TAG_HIDDEN, TAG_FF)


class _TagLayer(nn.Module):
    def __init__(self, encoder_width: int):
        super().__init__()
        self.crossattention = _CrossAttnBlock(encoder_width)
        self.intermediate = _Intermediate()
        self.output = _SelfOutput(TAG_FF)


class _TagEncoder(nn.Module):
    def __init__(self, encoder_width: int):
        super().__init__()
        self.layer = nn.ModuleList([_TagLayer(encoder_width) for _ in range(TAG_LAYERS)])


class _TagBert(nn.Module):
    def __init__(self, encoder_width: int):
        super().__init__()
        self.encoder = _TagEncoder(encoder_width)


class TagHead(nn.Module):
    """Holds ``tag_head``, ``tag_labels``, ``tag_fc`` under the reference's names and implements
    ``tag_forward``.  Use ``TagHead.from_reference(model)`` / ``load_state_dict(filtered)`` to take over the
    weights of a reference ``CLIP`` and ``patch_reference_model(model)`` to route ``model.tag_forward`` here."""

    def __init__(self, embed_dim: int, tag_list: Optional[List[str]] = None, *, fuse_kv: bool = True,
                 fuse_ln: bool = True, _kernels=None):
        super().__init__()
        self.fuse_kv = fuse_kv
        self.fuse_ln = fuse_ln
        self.tag_head = _TagBert(embed_dim)
        self.tag_labels = nn.Embedding(TAG_QUERIES, TAG_HIDDEN)
        self.tag_fc = nn.Linear(TAG_HIDDEN, 1)
        self.tag_list = list(tag_list) if tag_list is not None else None
        self.embed_dim = embed_dim
        self._k = _kernels
        self._step = 0
        self.reset_parameters()

    def reset_parameters(self):
        # BertPreTrainedModel._init_weights (bert.py:631-641): N(0, 0.02) weights, zero biases, unit LayerNorm
        for m in self.tag_head.modules():
            if isinstance(m, nn.Linear):
                nn.init.normal_(m.weight, mean=0.0, std=0.02)
                nn.init.zeros_(m.bias)
            elif isinstance(m, nn.LayerNorm):
                nn.init.ones_(m.weight)
                nn.init.zeros_(m.bias)

    @classmethod
    def from_reference(cls, model: nn.Module, **kw) -> "TagHead":
        sd = {k: v for k, v in model.state_dict().items() if k.startswith(("tag_head.", "tag_labels.", "tag_fc."))}
        d = sd["tag_head.encoder.layer.0.crossattention.self.key.weight"].shape[1]
        head = cls(d, getattr(model, "tag_list", None), **kw)
        head.load_state_dict(sd, strict=True)
        p = next(iter(sd.values()))
        return head.to(device=p.device)

    def tag_forward(self, tag_embeds: torch.Tensor, seed: Optional[int] = None) -> torch.Tensor:
        """tokens [b, N, D] -> tag logits [b, 44]  (model.py:337-352)."""
        if tag_embeds.dim() != 3 or tag_embeds.shape[-1] != self.embed_dim:
            raise ValueError(f"tag_embeds must be [b, N, {self.embed_dim}], got {tuple(tag_embeds.shape)}")
        bs = tag_embeds.shape[0]
        drop = TAG_DROPOUT if self.training else 0.0
        if seed is None:
            seed = int(torch.initial_seed() & 0x7FFFFFFFFFFFFFFF)
        self._step += 1
        h = None                                   # layer-0 hidden state is the shared label embedding
        label = self.tag_labels.weight             # [44, 768]
        layers = list(self.tag_head.encoder.layer)
        # bf16 on the GPU (the reference's amp_bf16 training): key and value projections of ALL layers as one
        # [b*N, D] x [D, 2*L*768] GEMM with a bias epilogue on the package's own tcgen05 kernel; K4 reads the per-layer
        # column slices in place.  fp32 (exact mode) keeps one library Linear per projection.
        amp_bf16 = tag_embeds.is_cuda and torch.is_autocast_enabled("cuda") and \
            torch.get_autocast_dtype("cuda") == torch.bfloat16
        fused_kv = None
        hooked = self._k is not None                 # injected kernel provider (the CPU contract model of the tests)
        if self.fuse_kv and (tag_embeds.is_cuda or hooked) and self.embed_dim % 8 == 0 and \
                (tag_embeds.dtype == torch.bfloat16 or amp_bf16) and \
                hasattr(_kernels(self._k), "tc_linear"):
            w = torch.cat([m.weight for l in layers for m in (l.crossattention.self.key, l.crossattention.self.value)], 0)
            bcat = torch.cat([m.bias for l in layers for m in (l.crossattention.self.key, l.crossattention.self.value)], 0)
            gbuf = _KVGradBuffer(2 * len(layers) * TAG_HIDDEN)
            fused_kv = (_FusedKVProjection.apply(tag_embeds.to(torch.bfloat16).contiguous(), w, bcat, 2 * len(layers),
                                                 self._k, gbuf), gbuf)
        K_ = _kernels(self._k)
        fused_ln = (self.fuse_ln and (amp_bf16 or (hooked and tag_embeds.dtype == torch.bfloat16))
                    and hasattr(K_, "ln_res_fwd") and K_.supports_ln_res(TAG_HIDDEN))
        for li, layer in enumerate(layers):
            ca = layer.crossattention
            if h is None:
                q = ca.self.query(label).unsqueeze(0).expand(bs, -1, -1)     # once, not per sample
                resid = label.unsqueeze(0)
            else:
                q = ca.self.query(h)
                resid = h
            if fused_kv is not None:
                k, v = fused_kv[0][2 * li], fused_kv[0][2 * li + 1]
                sink = (fused_kv[1], 2 * li * TAG_HIDDEN, (2 * li + 1) * TAG_HIDDEN)
                ctx = cross_attention(q.to(k.dtype), k, v, TAG_HEADS, drop, seed, self._step * TAG_LAYERS + li,
                                      _kernels=self._k, _sink=sink)
            else:
                k = ca.self.key(tag_embeds)
                v = ca.self.value(tag_embeds)
                ctx = cross_attention(q, k, v, TAG_HEADS, drop, seed, self._step * TAG_LAYERS + li, _kernels=self._k)
            if fused_ln:
                # dense (library GEMM, bias in its epilogue, bf16 out) -> ONE kernel for dropout + residual + LayerNorm
                base = ((self._step * TAG_LAYERS + li) << 2) | (1 << 40)      # Philox streams distinct from K4's
                a = residual_dropout_layer_norm(ca.output.dense(ctx), resid, ca.output.LayerNorm, drop, seed, base,
                                                _kernels=self._k)
                f = F.gelu(layer.intermediate.dense(a))
                h = residual_dropout_layer_norm(layer.output.dense(f), a, layer.output.LayerNorm, drop, seed, base + 1,
                                                _kernels=self._k)
            else:
                a = F.dropout(ca.output.dense(ctx.to(k.dtype)), drop, self.training)
                a = ca.output.LayerNorm(a + resid)
                f = F.gelu(layer.intermediate.dense(a))
                o = F.dropout(layer.output.dense(f), drop, self.training)
                h = layer.output.LayerNorm(o + a)
        return self.tag_fc(h).squeeze(-1)

    forward = tag_forward

    def control_word_indices(self, tag_logits: torch.Tensor) -> torch.Tensor:
        """int64 [b, 6]: per category the top-1 tag index (model.py:362-370)."""
        n = TAG_QUERIES // 2
        s = tag_logits.sigmoid()
        out, pos = [], 0
        for size in CATEGORY_SIZES:
            score = s[:, pos:pos + size] + s[:, n + pos:n + pos + size]
            out.append(torch.argsort(score, dim=-1, descending=True)[:, :1] + pos)
            pos += size
        return torch.cat(out, dim=-1)

    def prepare_control_words(self, tag_logits: torch.Tensor, samples=None) -> List[str]:
        """list[str], tags joined by ',' (model.py:354-383).  One device->host transfer of [b,6] ints."""
        if self.tag_list is None:
            raise RuntimeError("TagHead.tag_list is not set (pass the model's tag_list)")
        idx = self.control_word_indices(tag_logits).tolist()
        return [",".join(self.tag_list[i] for i in row) for row in idx]


def patch_reference_model(model: nn.Module, *, _kernels=None) -> nn.Module:
    """Route ``model.tag_forward`` (and ``encode_*``'s normalise, via ``model.xtag_normalize``) of a
    reference ``open_clip.CLIP`` instance through this package WITHOUT adding parameters or
    buffers: the fused head reads the model's own ``tag_head`` / ``tag_labels`` / ``tag_fc`` weights."""
    head = TagHead.__new__(TagHead)
    nn.Module.__init__(head)
    # share (not copy) the reference modules' parameters
    object.__setattr__(head, "_ref", model)
    head.embed_dim = model.tag_head.encoder.layer[0].crossattention.self.key.weight.shape[1]
    head.tag_list = getattr(model, "tag_list", None)
    head._k, head._step = _kernels, 0
    head.fuse_kv = True
    head.fuse_ln = True

    def tag_forward(tag_embeds):
        head.__dict__["training"] = model.training
        head._modules["tag_head"] = model.tag_head
        head._modules["tag_labels"] = model.tag_labels
        head._modules["tag_fc"] = model.tag_fc
        try:
            return TagHead.tag_forward(head, tag_embeds)
        finally:
            for n in ("tag_head", "tag_labels", "tag_fc"):
                head._modules.pop(n, None)

    model.tag_forward = tag_forward
    model.xtag_normalize = lambda x: l2_normalize(x, _kernels=_kernels)
    return model
