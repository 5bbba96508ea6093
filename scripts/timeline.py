#!/usr/bin/env python
"""Kernel timeline of one bench step (torch.profiler / CUPTI): where the non-kernel time of a step goes.
   python scripts/timeline.py            (1 GPU)      torchrun --nproc-per-node N scripts/timeline.py   (N GPUs)"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import xtag_clip_b200 as xt  # noqa: E402
from bench import _synth  # noqa: E402


def main():
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, D = int(os.environ.get("XB", 32768)), 1024
    b = B // world
    I, T = _synth(rank, b, D, "cpu", torch.bfloat16)
    I = I.to(dev).requires_grad_(True); T = T.to(dev).requires_grad_(True)
    ls = torch.tensor(2.659, device=dev, requires_grad=True)
    mod = xt.ClipLoss(local_loss=world > 1, gather_with_grad=world > 1, cache_labels=True, rank=rank, world_size=world,
                      cuda_graph=bool(int(os.environ.get("XGRAPH", "0"))), exchange=os.environ.get("XEXCH") or None,
                      pull_streams=int(os.environ.get("XPS", "2")))

    def step():
        I.grad = T.grad = ls.grad = None
        mod(I, T, ls.exp()).backward()

    for _ in range(5):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(3):
            step()
        torch.cuda.synchronize()
    if rank == 0:
        evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
        evs.sort(key=lambda e: e.time_range.start)
        t0 = evs[0].time_range.start
        last_end = None
        n = len(evs) // 3
        print(f"world={world} b={b}: {len(evs)} device events in 3 steps; middle step:")
        for e in evs[n:2 * n]:
            gap = (e.time_range.start - last_end) if last_end is not None else 0.0
            print(f"{(e.time_range.start - t0):10.1f} us  dur {e.time_range.end - e.time_range.start:8.1f}  gap {gap:7.1f}  {e.name[:90]}")
            last_end = max(last_end or 0, e.time_range.end)
    if world > 1:
        dist.barrier(); dist.destroy_process_group()


if __name__ == "__main__":
    main()
