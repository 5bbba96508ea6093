#!/usr/bin/env python
"""TQN fusion head (model.py:552-561 shape: B queries per sample over P+1 memory tokens, 4 layers, d = 512), bf16:
fusion_scores fwd+bwd with (a) K4 in one launch per layer, (b) K4 in 64-query chunks (round-1 path), (c) eager torch
attention (materialised [B, heads, B, P] scores, the reference's nn.MultiheadAttention formulation).
   python scripts/fusion_time.py [B] [P]"""
import json
import math
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import xtag_clip_b200 as xt  # noqa: E402
from xtag_clip_b200 import fusion_head as fh  # noqa: E402


def timeit(fn, warmup=2, iters=5):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def eager_attention(q, k, v, heads, dropout_p=0.0, seed=0, offset=0, **kw):
    b, Lq, H = q.shape
    dh = H // heads
    qh = q.reshape(b, Lq, heads, dh).permute(0, 2, 1, 3)
    kh = k.reshape(b, -1, heads, dh).permute(0, 2, 1, 3)
    vh = v.reshape(b, -1, heads, dh).permute(0, 2, 1, 3)
    p = torch.softmax((qh @ kh.transpose(-1, -2)) / math.sqrt(dh), dim=-1)
    p = F.dropout(p, dropout_p, dropout_p > 0)
    return (p @ vh).permute(0, 2, 1, 3).reshape(b, Lq, H)


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    P = int(sys.argv[2]) if len(sys.argv) > 2 else 197
    dev = torch.device("cuda")
    head = xt.FusionHead(512, 1, 4).to(dev, torch.bfloat16).train()
    img = torch.randn(B, P, 512, device=dev, dtype=torch.bfloat16, requires_grad=True)
    txt = torch.randn(B, 77, 512, device=dev, dtype=torch.bfloat16, requires_grad=True)

    def step():
        img.grad = txt.grad = None
        head.zero_grad(set_to_none=True)
        s = xt.fusion_scores(head, img, txt)
        xt.DQNCOSLoss()(s.float()).backward()

    rec = dict(B=B, P=P)
    rec["one_launch_ms"] = timeit(step)
    orig_attend = fh.FusionHead._attend

    def chunked(self, q, kv, seed, offset):
        E = self.d_model
        k, v = kv[..., :E], kv[..., E:]
        drop = self.p_drop if self.training else 0.0
        out = [fh.cross_attention(q[:, c:c + 64], k, v, self.nhead, drop, seed, offset * 64 + ci, _kernels=self._k)
               for ci, c in enumerate(range(0, q.shape[1], 64))]
        return torch.cat(out, dim=1)

    fh.FusionHead._attend = chunked
    try:
        rec["chunked_ms"] = timeit(step)
        orig_ca = fh.cross_attention
        fh.cross_attention = eager_attention
        fh.FusionHead._attend = lambda self, q, kv, seed, offset: eager_attention(
            q, kv[..., :self.d_model], kv[..., self.d_model:], self.nhead, self.p_drop if self.training else 0.0)
        try:
            rec["eager_torch_ms"] = timeit(step)
        except torch.OutOfMemoryError:
            rec["eager_torch_ms"] = None
        finally:
            fh.cross_attention = orig_ca
    finally:
        fh.FusionHead._attend = orig_attend
    print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
