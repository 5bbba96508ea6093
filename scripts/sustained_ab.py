#!/usr/bin/env python
"""Sustained-power A/B on ONE GPU: alternates >= 1 s continuous runs of (a) the raw config-5 step (K.clip_fwd +
K.clip_bwd) under each tuning value and (b) cuBLAS bf16 8192^3 (the driver's "sustained" reference), after a heat-up.
Prints executed TFLOP/s per leg so the kernels and the library are compared in the same power state on the same box.
   python scripts/sustained_ab.py --tunes 0x800,0x1000800 [--secs 1.5] [--rounds 3]"""
import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import _synth  # noqa: E402
from xtag_clip_b200.kernels import CudaKernels  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32768)
    ap.add_argument("--dim", type=int, default=1024)
    ap.add_argument("--tunes", default="0x800,0x1000800")
    ap.add_argument("--secs", type=float, default=1.5)
    ap.add_argument("--rounds", type=int, default=3)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    K = CudaKernels(impl=2)
    I, T = _synth(0, a.batch, a.dim, "cpu", torch.bfloat16)
    I, T = I.to(dev), T.to(dev)
    s = torch.tensor([14.285714], device=dev)
    g1 = torch.tensor(1.0, device=dev)
    B, D = a.batch, a.dim
    w = (0.5 / B, 0.5 / B, 1.0 / B)
    X = torch.randn(8192, 8192, device=dev, dtype=torch.bfloat16)
    Y = torch.randn(8192, 8192, device=dev, dtype=torch.bfloat16)

    def raw():
        row, col, diag = K.clip_fwd(I, T, s, 0)
        K.clip_bwd(I, T, s, 0, row, col, *w, g1, True, True, torch.bfloat16)

    def mm():
        for _ in range(4):
            torch.matmul(X, Y)

    legs = [(f"raw {t}", int(t, 0)) for t in a.tunes.split(",")] + [("cublas 8192^3 x4", None)]
    flop = {None: 4 * 2 * 8192 ** 3}
    acc = {n: [] for n, _ in legs}
    for rnd in range(a.rounds + 1):                  # round 0 = heat-up, discarded
        for name, tune in legs:
            f = mm if tune is None else raw
            if tune is not None:
                K.lib.xtag_set_tune(tune)
            f()
            n = 0
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t_end = time.time() + a.secs
            e0.record()
            while time.time() < t_end:
                for _ in range(5):
                    f()
                n += 5
                torch.cuda.current_stream().synchronize() if n % 40 == 0 else None
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / n
            fl = flop[None] if tune is None else 8.0 * B * B * D
            if rnd:
                acc[name].append(dict(ms=ms, tflops=fl / ms / 1e9))
    out = {n: dict(ms=sum(x["ms"] for x in v) / len(v), tflops_executed=sum(x["tflops"] for x in v) / len(v)) for n, v in acc.items()}
    print(json.dumps(out))
    K.lib.xtag_set_tune(0x800)


if __name__ == "__main__":
    main()
