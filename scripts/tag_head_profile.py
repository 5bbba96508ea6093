#!/usr/bin/env python
"""Per-kernel time of one tag-head training step at BASELINE config 3 (b=1024, N=197, D=512, bf16 autocast).
   python scripts/tag_head_profile.py [--no-fuse]"""
import os
import sys
from collections import defaultdict

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import xtag_clip_b200 as xt  # noqa: E402


def main():
    fuse = "--no-fuse" not in sys.argv
    b, N, D = 1024, 197, 512
    dev = torch.device("cuda")
    head = xt.TagHead(D, fuse_kv=fuse).to(dev).train()
    asl = xt.AsymmetricLoss(gamma_neg=4, gamma_pos=1, clip=0.05)
    tok = torch.randn(b, N, D, device=dev, dtype=torch.bfloat16, requires_grad=True)
    y = (torch.rand(b, 22, device=dev) > 0.7).float().repeat(1, 2)

    def step():
        tok.grad = None
        head.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            logits = head.tag_forward(tok)
        asl(logits.float(), y).backward()

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(3):
            step()
        torch.cuda.synchronize()
    agg = defaultdict(lambda: [0.0, 0])
    t_min, t_max = None, None
    for e in prof.events():
        if e.device_type == torch.autograd.DeviceType.CUDA:
            d = e.time_range.end - e.time_range.start
            agg[e.name[:110]][0] += d / 3
            agg[e.name[:110]][1] += 1
            t_min = e.time_range.start if t_min is None else min(t_min, e.time_range.start)
            t_max = e.time_range.end if t_max is None else max(t_max, e.time_range.end)
    tot = sum(v[0] for v in agg.values())
    print(f"fuse_kv={fuse}: wall {(t_max - t_min) / 3:.0f} us/step, kernel sum {tot:.0f} us/step")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:28]:
        print(f"{v[0]:9.1f} us  x{v[1] // 3:<3d} {k}")


if __name__ == "__main__":
    main()
