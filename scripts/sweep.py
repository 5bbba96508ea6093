#!/usr/bin/env python
"""Config sweep on one B200: fused ClipLoss (this repo) vs the stock PyTorch expression of the reference
(loss.py:116-137: two GEMMs + two F.cross_entropy under bf16 autocast) on the same GPU, plus K3 / K4 bandwidth.
Writes one JSON object per line (gpurun_out/sweep.jsonl when run under gpurun)."""
import json
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import xtag_clip_b200 as xt  # noqa: E402
from xtag_clip_b200.kernels import default_kernels  # noqa: E402


def timeit(fn, warmup=3, iters=10):
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


def feats(B, D, dtype):
    g = torch.Generator().manual_seed(B + D)
    i = torch.randn(B, D, generator=g)
    t = 0.5 * i + 0.5 * torch.randn(B, D, generator=g)
    n = torch.nn.functional.normalize
    return n(i, dim=-1).to(dtype).cuda(), n(t, dim=-1).to(dtype).cuda()


def main():
    out = []
    fused = xt.ClipLoss()
    for B, D in [(256, 512), (512, 512), (1024, 512), (2048, 512), (4096, 512), (8192, 768), (16384, 1024),
                 (32768, 1024)]:
        I, T = feats(B, D, torch.bfloat16)
        I.requires_grad_(True)
        T.requires_grad_(True)
        ls = torch.tensor(2.659, device="cuda", requires_grad=True)

        def step_fused():
            I.grad = T.grad = ls.grad = None
            fused(I, T, ls.exp()).backward()

        def step_torch():
            I.grad = T.grad = ls.grad = None
            with torch.autocast("cuda", dtype=torch.bfloat16):
                li, lt = fused.get_logits(I, T, ls.exp())          # the reference's expressions
                lab = torch.arange(B, device="cuda")
                loss = (torch.nn.functional.cross_entropy(li, lab) + torch.nn.functional.cross_entropy(lt, lab)) / 2
            loss.backward()

        iters = 20 if B <= 8192 else 5
        tf = timeit(step_fused, iters=iters)
        try:
            tt = timeit(step_torch, iters=iters)
        except torch.OutOfMemoryError:
            tt = None
        rec = dict(kind="cliploss_fwd_bwd", B=B, D=D, fused_ms=tf, torch_eager_ms=tt,
                   fused_samples_per_s=B / tf * 1e3, algorithmic_tflops=6.0 * B * B * D / (tf * 1e-3) / 1e12,
                   speedup_vs_torch=(tt / tf) if tt else None)
        print(json.dumps(rec), flush=True)
        out.append(rec)
    K = default_kernels()
    for rows, dim in [(4096, 512), (32768, 1024), (262144, 1024)]:
        x = torch.randn(rows, dim, device="cuda")
        t = timeit(lambda: K.l2norm_fwd(x, torch.bfloat16, 1e-12))
        rec = dict(kind="l2norm_fwd_f32_to_bf16", rows=rows, dim=dim, ms=t, gbs=rows * dim * 6 / (t * 1e-3) / 1e9)
        print(json.dumps(rec), flush=True)
        out.append(rec)
    for b, N in [(1024, 50), (1024, 197), (1024, 257)]:
        q = torch.randn(b, 44, 768, device="cuda", dtype=torch.bfloat16)
        kv = torch.randn(b, N, 1536, device="cuda", dtype=torch.bfloat16)
        k, v = kv[..., :768], kv[..., 768:]
        t = timeit(lambda: K.xattn_fwd(q, k, v, 4, 1 / math.sqrt(192), 0.0, 0, 0))
        rec = dict(kind="xattn_fwd_bf16", b=b, N=N, ms=t, gbs=b * (88 + 2 * N) * 1536 / (t * 1e-3) / 1e9)
        print(json.dumps(rec), flush=True)
        out.append(rec)
        o, lse = K.xattn_fwd(q, k, v, 4, 1 / math.sqrt(192), 0.0, 0, 0)
        do = torch.randn_like(o)
        t = timeit(lambda: K.xattn_bwd(q, k, v, o, do, lse, 4, 1 / math.sqrt(192), 0.0, 0, 0))
        # algorithmic bytes of the backward: read q,k,v,o,do ; write dq,dk,dv
        rec = dict(kind="xattn_bwd_bf16", b=b, N=N, ms=t, gbs=b * (4 * 44 + 4 * N) * 1536 / (t * 1e-3) / 1e9)
        print(json.dumps(rec), flush=True)
        out.append(rec)
    # whole tag head (config 3 shape): fused K4 attention core vs the reference's eager attention (bert.py:219-274:
    # two batched GEMMs + softmax + dropout + permute copies) with the same weights, fwd + bwd, bf16 autocast
    import torch.nn.functional as F
    from xtag_clip_b200 import tag_head as th

    def eager_attention(q, k, v, heads, dropout_p=0.0, seed=0, offset=0, *, _kernels=None):
        b, Lq, H = q.shape
        dh = H // heads
        qh = q.reshape(b, Lq, heads, dh).permute(0, 2, 1, 3)
        kh = k.reshape(b, -1, heads, dh).permute(0, 2, 1, 3)
        vh = v.reshape(b, -1, heads, dh).permute(0, 2, 1, 3)
        p = torch.softmax((qh @ kh.transpose(-1, -2)) / math.sqrt(dh), dim=-1)
        p = F.dropout(p, dropout_p, dropout_p > 0)
        return (p @ vh).permute(0, 2, 1, 3).reshape(b, Lq, H)

    for b, N, D in [(1024, 197, 512), (1024, 50, 512), (256, 257, 1024)]:
        head = xt.TagHead(D).cuda().train()
        tok = torch.randn(b, N, D, device="cuda", dtype=torch.bfloat16, requires_grad=True)

        def step_head():
            tok.grad = None
            head.zero_grad(set_to_none=True)
            with torch.autocast("cuda", dtype=torch.bfloat16):
                out = head.tag_forward(tok)
            out.float().sum().backward()

        t_fused = timeit(step_head, iters=10)
        orig = th.cross_attention
        th.cross_attention = eager_attention
        try:
            t_eager = timeit(step_head, iters=10)
        finally:
            th.cross_attention = orig
        rec = dict(kind="tag_head_fwd_bwd_train", b=b, N=N, D=D, fused_ms=t_fused, eager_attention_ms=t_eager,
                   speedup=t_eager / t_fused)
        print(json.dumps(rec), flush=True)
        out.append(rec)
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/sweep.jsonl", "w") as fh:
        for r in out:
            fh.write(json.dumps(r) + "\n")


if __name__ == "__main__":
    main()
