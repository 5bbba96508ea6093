#!/usr/bin/env python
"""A/B of the runtime tuning bits of the tcgen05 kernels (xtag_set_tune) on one GPU.

    python scripts/tune_sweep.py [--batch 32768] [--dim 1024] [--iters 6] [--tunes 0x0,0x400,...]

For every setting: per-kernel durations from the library's own CUDA events (xtag_prof_*) for the fused forward
(K1), the dS producer and the two gradient GEMMs, at BASELINE config 5's single-GPU shape.  One JSON line per
setting on stdout (and appended to gpurun_out/tune_sweep.jsonl when that directory exists).
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32768)
    ap.add_argument("--dim", type=int, default=1024)
    ap.add_argument("--iters", type=int, default=6)
    ap.add_argument("--tunes", default="0x0,0x400,0x4,0x8,0xc,0x100,0x200,0x308,0x10c")
    args = ap.parse_args()
    import torch
    from xtag_clip_b200.kernels import CudaKernels

    K = CudaKernels(impl=2)
    lib = K.lib
    B, D = args.batch, args.dim
    g = torch.Generator().manual_seed(0)
    I = torch.nn.functional.normalize(torch.randn(B, D, generator=g), dim=-1)
    T = torch.nn.functional.normalize(0.5 * I + 0.5 * torch.randn(B, D, generator=g), dim=-1)
    I, T = I.bfloat16().cuda(), T.bfloat16().cuda()
    s = torch.tensor([14.285714], device="cuda")
    gout = torch.tensor(1.0, device="cuda")
    w = (0.5 / B, 0.5 / B, 1.0 / B)
    names = {0: "fwd", 1: "ds", 2: "gemm"}
    out_path = os.path.join(ROOT, "gpurun_out", "tune_sweep.jsonl")
    ref = None
    for tune in [int(t, 0) for t in args.tunes.split(",")]:
        lib.xtag_set_tune(tune)
        for _ in range(2):
            row, col, diag = K.clip_fwd(I, T, s, 0)
            dA, dB, ds = K.clip_bwd(I, T, s, 0, row, col, *w, gout, True, True, torch.bfloat16)
        torch.cuda.synchronize()
        lib.xtag_prof_enable(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.iters):
            row, col, diag = K.clip_fwd(I, T, s, 0)
            dA, dB, ds = K.clip_bwd(I, T, s, 0, row, col, *w, gout, True, True, torch.bfloat16)
        e1.record()
        torch.cuda.synchronize()
        cap = 16 * args.iters + 16
        tags, tms, work = (ctypes.c_int * cap)(), (ctypes.c_float * cap)(), (ctypes.c_double * cap)()
        n = lib.xtag_prof_read(tags, tms, work, cap)
        lib.xtag_prof_enable(0)
        per = {}
        for i in range(n):
            per.setdefault(names.get(tags[i], str(tags[i])), []).append(tms[i])
        loss = float(K.clip_loss(row, diag, col, 0))
        chk = (float(dA.float().abs().sum()), float(dB.float().abs().sum()), float(ds))
        if ref is None:
            ref = chk
        rec = dict(tune=hex(tune), batch=B, dim=D, ms_per_step=e0.elapsed_time(e1) / args.iters,
                   kernels_ms={k: sum(v) / len(v) for k, v in per.items()},
                   kernels_min_ms={k: min(v) for k, v in per.items()}, loss=loss,
                   grad_l1_vs_first=[c / r if r else None for c, r in zip(chk, ref)])
        line = json.dumps(rec)
        print(line, flush=True)
        if os.path.isdir(os.path.dirname(out_path)):
            with open(out_path, "a") as f:
                f.write(line + "\n")
    lib.xtag_set_tune(0)


if __name__ == "__main__":
    main()
