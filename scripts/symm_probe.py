#!/usr/bin/env python
"""Probe of torch symmetric memory on the GPU box (2+ GPUs): peer pulls by copy engine, barrier cost, and whether both
overlap with a persistent tcgen05 kernel that occupies every SM."""
import os
import sys
import time

import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from xtag_clip_b200.kernels import default_kernels  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    K = default_kernels()
    b, D = 4096, 1024
    buf = symm_mem.empty((2, b, D), dtype=torch.bfloat16, device=dev)
    hdl = symm_mem.rendezvous(buf, dist.group.WORLD)
    if rank == 0:
        print("rendezvous ok; multicast:", hdl.has_multicast_support, "world", hdl.world_size, flush=True)
    buf.fill_(float(rank + 1))
    hdl.barrier(channel=0)
    peers = [hdl.get_buffer(p, (2, b, D), torch.bfloat16) for p in range(world)]
    out = torch.empty((world, b, D), dtype=torch.bfloat16, device=dev)
    side = torch.cuda.Stream()

    def pull_all(stream):
        with torch.cuda.stream(stream):
            for j in range(1, world):
                p = (rank + j) % world
                out[p].copy_(peers[p][0], non_blocking=True)

    # correctness
    pull_all(side)
    torch.cuda.synchronize()
    hdl.barrier(channel=0)
    ok = all(float(out[p][0, 0]) == p + 1 for p in range(world) if p != rank)
    # timing: pulls alone
    def timed(fn, n=10):
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n * 1e3

    t_pull = timed(lambda: pull_all(torch.cuda.current_stream()))
    t_bar = timed(lambda: hdl.barrier(channel=0))
    A = torch.randn(4096, 8192, device=dev).bfloat16()
    Bm = torch.randn(8192, 8192, device=dev).bfloat16()
    t_gemm = timed(lambda: K.tc_gemm_nt(A, Bm, torch.bfloat16))

    def overlapped():
        cur = torch.cuda.current_stream()
        side.wait_stream(cur)
        K.tc_gemm_nt(A, Bm, torch.bfloat16)
        with torch.cuda.stream(side):
            hdl.barrier(channel=1)
        pull_all(side)
        cur.wait_stream(side)

    t_ovl = timed(overlapped)
    ag = torch.empty((world * b, D), dtype=torch.bfloat16, device=dev)
    t_nccl = timed(lambda: dist.all_gather_into_tensor(ag, buf[0]))
    if rank == 0:
        gb = (world - 1) * b * D * 2 / 1e9
        print(f"world={world} correct={ok} pull {(world-1)}x8MB: {t_pull:.1f} us ({gb / t_pull * 1e6:.0f} GB/s)  barrier {t_bar:.1f} us  "
              f"gemm {t_gemm:.1f} us  gemm||barrier+pull {t_ovl:.1f} us  nccl all_gather {t_nccl:.1f} us", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
