#!/usr/bin/env python
"""K4 timings on one GPU: forward, two-kernel backward (tune 0) and single-pass backward (tune 0x800) at the tag-head
shapes, as algorithmic GB/s (SURVEY.md section 8d byte counts).  One JSON line per measurement."""
import json
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from xtag_clip_b200.kernels import default_kernels  # noqa: E402


def timeit(fn, warmup=3, iters=20):
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


def profile_once():
    """one forward + backward at b=1024, N=197 with and without dropout (for an ncu capture; the backward kernel is the
    one XTAG_TC_TUNE selects)"""
    K = default_kernels()
    sc = 1 / math.sqrt(192)
    q = torch.randn(1024, 44, 768, device="cuda", dtype=torch.bfloat16)
    kv = torch.randn(1024, 197, 1536, device="cuda", dtype=torch.bfloat16)
    k, v = kv[..., :768], kv[..., 768:]
    for p in (0.0, 0.1):
        o, lse = K.xattn_fwd(q, k, v, 4, sc, p, 1, 2)
        K.xattn_bwd(q, k, v, o, torch.randn_like(o), lse, 4, sc, p, 1, 2)
    torch.cuda.synchronize()


def main():
    if "--profile" in sys.argv:
        return profile_once()
    K = default_kernels()
    sc = 1 / math.sqrt(192)
    out = []
    for b, N in [(1024, 50), (1024, 197), (1024, 257), (4096, 197)]:
        q = torch.randn(b, 44, 768, device="cuda", dtype=torch.bfloat16)
        kv = torch.randn(b, N, 1536, device="cuda", dtype=torch.bfloat16)
        k, v = kv[..., :768], kv[..., 768:]
        for p in (0.0, 0.1):
            t = timeit(lambda: K.xattn_fwd(q, k, v, 4, sc, p, 1, 2))
            out.append(dict(kind="xattn_fwd", b=b, N=N, p_drop=p, ms=t, gbs=b * (88 + 2 * N) * 1536 / (t * 1e-3) / 1e9))
            o, lse = K.xattn_fwd(q, k, v, 4, sc, p, 1, 2)
            do = torch.randn_like(o)
            for tune in (0x000, 0x800):
                old = K.lib.xtag_set_tune(tune)
                try:
                    t = timeit(lambda: K.xattn_bwd(q, k, v, o, do, lse, 4, sc, p, 1, 2))
                finally:
                    K.lib.xtag_set_tune(old)
                out.append(dict(kind="xattn_bwd", tune=hex(tune), b=b, N=N, p_drop=p, ms=t,
                                gbs=b * (4 * 44 + 4 * N) * 1536 / (t * 1e-3) / 1e9))
            print("\n".join(json.dumps(r) for r in out[-3:]), flush=True)
    if os.path.isdir("gpurun_out"):
        with open("gpurun_out/xattn_time.jsonl", "w") as fh:
            for r in out:
                fh.write(json.dumps(r) + "\n")


if __name__ == "__main__":
    main()
