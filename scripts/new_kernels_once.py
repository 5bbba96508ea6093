#!/usr/bin/env python
"""One tag-head training step (config 3) and one SigLipLoss step (config 5 shape) -- the command profiled with ncu for
the round-2 kernels (K6, the bias-epilogue GEMM, split-K, the sigmoid epilogue)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import xtag_clip_b200 as xt  # noqa: E402
from bench import _synth  # noqa: E402

dev = torch.device("cuda")
head = xt.TagHead(512).to(dev).train()
asl = xt.AsymmetricLoss(gamma_neg=4, gamma_pos=1, clip=0.05)
tok = torch.randn(1024, 197, 512, device=dev, dtype=torch.bfloat16, requires_grad=True)
y = (torch.rand(1024, 22, device=dev) > 0.7).float().repeat(1, 2)
for _ in range(2):
    tok.grad = None
    head.zero_grad(set_to_none=True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        logits = head.tag_forward(tok)
    asl(logits.float(), y).backward()
I, T = _synth(0, 32768, 1024, "cpu", torch.bfloat16)
I = I.to(dev).requires_grad_(True)
T = T.to(dev).requires_grad_(True)
s = torch.tensor(10.0, device=dev, requires_grad=True)
b = torch.tensor(-10.0, device=dev, requires_grad=True)
for _ in range(2):
    xt.SigLipLoss()(I, T, s, b).backward()
torch.cuda.synchronize()
print("ok")
