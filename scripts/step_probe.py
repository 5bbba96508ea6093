#!/usr/bin/env python
"""Where does a config-5 step spend its time on ONE GPU?  Times, interleaved in the sustained power state:
   graph   ClipLoss(cuda_graph=True) + backward          (what bench.py times)
   eager   ClipLoss() + backward
   raw     K.clip_fwd + K.clip_bwd through the C ABI       (no autograd, no copies)
   python scripts/step_probe.py [--batch 32768] [--dim 1024] [--rounds 6] [--iters 10]"""
import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import xtag_clip_b200 as xt  # noqa: E402
from bench import _synth  # noqa: E402
from xtag_clip_b200.kernels import default_kernels  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32768)
    ap.add_argument("--dim", type=int, default=1024)
    ap.add_argument("--rounds", type=int, default=6)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--heat", type=float, default=3.0)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    I, T = _synth(0, a.batch, a.dim, "cpu", torch.bfloat16)
    I = I.to(dev).requires_grad_(True)
    T = T.to(dev).requires_grad_(True)
    ls = torch.tensor(2.659, device=dev, requires_grad=True)
    graph = xt.ClipLoss(cuda_graph=True)
    eager = xt.ClipLoss()
    K = default_kernels()
    s = torch.tensor([14.285714], device=dev)
    g1 = torch.tensor(1.0, device=dev)
    B = a.batch
    w = (0.5 / B, 0.5 / B, 1.0 / B)

    def step(mod):
        I.grad = T.grad = ls.grad = None
        mod(I, T, ls.exp()).backward()

    def raw():
        row, col, diag = K.clip_fwd(I.detach(), T.detach(), s, 0)
        K.clip_bwd(I.detach(), T.detach(), s, 0, row, col, *w, g1, True, True, torch.bfloat16)

    modes = {"graph": lambda: step(graph), "eager": lambda: step(eager), "raw": raw}
    for f in modes.values():
        for _ in range(3):
            f()
    torch.cuda.synchronize()
    t_end = time.time() + a.heat
    while time.time() < t_end:
        for _ in range(10):
            raw()
        torch.cuda.synchronize()
    acc = {k: [] for k in modes}
    for _ in range(a.rounds):
        for k, f in modes.items():
            f()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(a.iters):
                f()
            e1.record()
            torch.cuda.synchronize()
            acc[k].append(e0.elapsed_time(e1) / a.iters)
    print(json.dumps({k: dict(mean=sum(v) / len(v), min=min(v), max=max(v)) for k, v in acc.items()}))


if __name__ == "__main__":
    main()
