"""Multi-GPU parity worker (launched by tests/test_gpu_dist.py or by hand):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 \
        tests/dist_gpu_worker.py

One process per GPU over NCCL.  Every rank runs xtag_clip_b200.ClipLoss in all four local_loss x gather_with_grad
modes on CUDA and checks its own loss / gradients against the single-process emulation of the reference's
per-rank results (oracle.clip_loss_world, pinned to the reference under gloo by tests/test_oracle_golden.py)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

import oracle  # noqa: E402
import xtag_clip_b200 as xt  # noqa: E402


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-300))


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    failures = []
    # sizes keep the fp64 world emulation (W ranks x 4 modes on the host) to a few seconds even at W = 8
    for (b, D, scale, dtype, ltol, gtol) in [(48, 64, 14.285714, torch.float32, 1e-5, 1e-5),
                                             (256, 256, 30.0, torch.bfloat16, 1e-3, 2e-2),
                                             (384, 512, 14.285714, torch.bfloat16, 1e-3, 2e-2)]:
        g = torch.Generator().manual_seed(100 + b)
        I_all = torch.nn.functional.normalize(torch.randn(b * world, D, generator=g), dim=-1)
        T_all = torch.nn.functional.normalize(0.3 * I_all + 0.7 * torch.randn(b * world, D, generator=g), dim=-1)
        I_all, T_all = I_all.to(dtype), T_all.to(dtype)
        Il = [I_all[r * b:(r + 1) * b].double() for r in range(world)]
        Tl = [T_all[r * b:(r + 1) * b].double() for r in range(world)]
        torch.set_num_threads(max(1, (os.cpu_count() or 8) // world))
        for ll in (False, True):
            for gwg in (False, True):
                if b == 384 and not (ll and gwg):
                    continue                      # largest shape: performance mode only
                losses, dI, dT, ds = oracle.clip_loss_world(Il, Tl, scale, ll, gwg)
                I = I_all[rank * b:(rank + 1) * b].to(dev).requires_grad_(True)
                T = T_all[rank * b:(rank + 1) * b].to(dev).requires_grad_(True)
                s = torch.tensor(scale, device=dev, requires_grad=True)
                mod = xt.ClipLoss(local_loss=ll, gather_with_grad=gwg, cache_labels=True, rank=rank, world_size=world)
                loss = mod(I, T, s)
                loss.backward()
                tag = f"b={b} D={D} {dtype} ll={ll} gwg={gwg} rank={rank}"
                e = (rel(loss, losses[rank]), rel(I.grad, dI[rank]), rel(T.grad, dT[rank]))
                if e[0] > ltol or e[1] > gtol or e[2] > gtol:
                    failures.append(f"{tag}: loss {e[0]:.2e} dI {e[1]:.2e} dT {e[2]:.2e}")
                # d(logit_scale): compare what DDP would reduce (sum over ranks)
                tot = s.grad.detach().clone()
                dist.all_reduce(tot)
                ref_tot = float(sum(ds))
                if abs(tot.item() - ref_tot) > 2e-2 * abs(ref_tot) + 5e-5:
                    failures.append(f"{tag}: dscale sum {tot.item():.6e} vs {ref_tot:.6e}")
    # CUDA-graph replay of the performance mode (captured barriers, copy-engine pulls and stream forks)
    b, D, scale = 512, 256, 14.285714
    g = torch.Generator().manual_seed(7)
    I_all = torch.nn.functional.normalize(torch.randn(b * world, D, generator=g), dim=-1).bfloat16()
    T_all = torch.nn.functional.normalize(0.3 * I_all.float() + 0.7 * torch.randn(b * world, D, generator=g), dim=-1).bfloat16()
    mods = [xt.ClipLoss(local_loss=True, gather_with_grad=True, rank=rank, world_size=world, cuda_graph=cg)
            for cg in (True, False)]
    for it in range(3):
        res = []
        for mod in mods:
            I = (I_all[rank * b:(rank + 1) * b].float() * (1.0 + 0.1 * it)).bfloat16().to(dev).requires_grad_(True)
            T = T_all[rank * b:(rank + 1) * b].to(dev).requires_grad_(True)
            s = torch.tensor(scale, device=dev, requires_grad=True)
            loss = mod(I, T, s)
            loss.backward()
            res.append((loss.detach(), I.grad, T.grad, s.grad))
        for name, a, bb in zip(("loss", "dI", "dT", "ds"), res[0], res[1]):
            if rel(a, bb) > 1e-5:
                failures.append(f"cuda_graph it={it} rank={rank}: {name} differs from eager by {rel(a, bb):.2e}")
    # The SHIPPED shard shapes (BASELINE config 5: b = 4096, D = 1024; config 4: global 8192, D = 768), performance
    # mode, both exchanges, eager and CUDA-graph replay, against the already validated single-GPU ClipLoss on the
    # concatenated batch: mean of the rank losses == global loss, rank feature gradients == W x the global gradient's
    # slice, sum of the rank d(logit_scale) == W x the global one ("Gradient scaling").
    small = os.environ.get("XTAG_DIST_SMALL") == "1"
    shapes = [("C5", 4096, 1024), ("C4", 8192 // world, 768)]
    if small:
        shapes = [("C5/4", 1024, 1024), ("C4/4", 2048 // world, 768)]
    for name, b, D in shapes:
        B = b * world
        g = torch.Generator().manual_seed(31 + D)
        I_all = torch.nn.functional.normalize(torch.randn(B, D, generator=g), dim=-1)
        T_all = torch.nn.functional.normalize(0.5 * I_all + 0.5 * torch.randn(B, D, generator=g), dim=-1).bfloat16()
        I_all = I_all.bfloat16()
        Ig = I_all.to(dev).requires_grad_(True)
        Tg = T_all.to(dev).requires_grad_(True)
        sg = torch.tensor(14.285714, device=dev, requires_grad=True)
        lg = xt.ClipLoss()(Ig, Tg, sg)
        lg.backward()
        lo, hi = rank * b, (rank + 1) * b
        for exch, ps in (("pull", 2), ("pull", 1), ("push", 2)):      # ps: copy streams of the exchange (default 2)
            for cg in (False, True):
                mod = xt.ClipLoss(local_loss=True, gather_with_grad=True, rank=rank, world_size=world, exchange=exch,
                                  cuda_graph=cg, pull_streams=ps)
                for it in range(3 if cg else 2):        # replays after the capture; slot parity of the eager path
                    I = I_all[lo:hi].to(dev).requires_grad_(True)
                    T = T_all[lo:hi].to(dev).requires_grad_(True)
                    s = torch.tensor(14.285714, device=dev, requires_grad=True)
                    loss = mod(I, T, s)
                    loss.backward()
                    tag = f"{name} b={b} D={D} W={world} {exch}/{ps} graph={cg} it={it} rank={rank}"
                    lm = loss.detach().clone()
                    dist.all_reduce(lm)
                    lm /= world
                    e = (rel(lm, lg), rel(I.grad, world * Ig.grad[lo:hi]), rel(T.grad, world * Tg.grad[lo:hi]))
                    if e[0] > 1e-3 or e[1] > 2e-2 or e[2] > 2e-2:
                        failures.append(f"{tag}: mean loss {e[0]:.2e} dI {e[1]:.2e} dT {e[2]:.2e}")
                    tot = s.grad.detach().clone()
                    dist.all_reduce(tot)
                    ref_tot = world * float(sg.grad)
                    if abs(tot.item() - ref_tot) > 2e-2 * abs(ref_tot) + 5e-5:
                        failures.append(f"{tag}: dscale sum {tot.item():.6e} vs {ref_tot:.6e}")
                path = mod.last_path
                if path["exchange"] != f"symmetric-memory {exch}" or path["cuda_graph"] != cg:
                    failures.append(f"{tag}: fell back to {path}")
        del Ig, Tg, lg
        torch.cuda.empty_cache()
    # rank skew: the flag-gated K1 of the push exchange waits inside the kernel for a late peer (here 1.5 s; the spin
    # budget of the tests is 20 s) and still produces the right loss
    b, D = 512, 256
    g = torch.Generator().manual_seed(5)
    I_all = torch.nn.functional.normalize(torch.randn(b * world, D, generator=g), dim=-1).bfloat16()
    T_all = torch.nn.functional.normalize(0.3 * I_all.float() + 0.7 * torch.randn(b * world, D, generator=g), dim=-1).bfloat16()
    mod = xt.ClipLoss(local_loss=True, gather_with_grad=True, rank=rank, world_size=world, exchange="push")
    vals = []
    for it in range(3):
        if it == 1 and rank == world - 1:
            torch.cuda.synchronize()
            import time
            # XTAG_LONG_SKEW_S=70 (opt-in, with XTAG_SPIN_TIMEOUT_MS unset or larger): a rank that stalls for over a
            # minute -- data-loader hiccup, rank-0-only checkpoint -- must not trap the waiting ranks
            time.sleep(float(os.environ.get("XTAG_LONG_SKEW_S", "1.5")))
        I = I_all[rank * b:(rank + 1) * b].to(dev).requires_grad_(True)
        T = T_all[rank * b:(rank + 1) * b].to(dev).requires_grad_(True)
        loss = mod(I, T, torch.tensor(14.285714, device=dev))
        loss.backward()
        vals.append((loss.detach().clone(), I.grad.clone()))
    torch.cuda.synchronize()
    for it in (1, 2):
        if rel(vals[it][0], vals[0][0]) > 1e-6 or rel(vals[it][1], vals[0][1]) > 1e-6:
            failures.append(f"skewed push step {it} rank={rank}: differs from the unskewed step")
    # sharded SigLipLoss against the oracle's emulation of the reference's per-rank results
    b, D, scale, bias = 256, 256, 12.0, -8.0
    g = torch.Generator().manual_seed(77)
    I_all = torch.nn.functional.normalize(torch.randn(b * world, D, generator=g), dim=-1).bfloat16()
    T_all = torch.nn.functional.normalize(0.4 * I_all.float() + 0.6 * torch.randn(b * world, D, generator=g), dim=-1).bfloat16()
    Il = [I_all[r * b:(r + 1) * b].double() for r in range(world)]
    Tl = [T_all[r * b:(r + 1) * b].double() for r in range(world)]
    lo, dI, dT, ds, db = oracle.siglip_loss_world(Il, Tl, scale, bias)
    I = I_all[rank * b:(rank + 1) * b].to(dev).requires_grad_(True)
    T = T_all[rank * b:(rank + 1) * b].to(dev).requires_grad_(True)
    s = torch.tensor(scale, device=dev, requires_grad=True)
    bb = torch.tensor(bias, device=dev, requires_grad=True)
    loss = xt.SigLipLoss(rank=rank, world_size=world)(I, T, s, bb)
    loss.backward()
    e = (rel(loss, lo[rank]), rel(I.grad, dI[rank]), rel(T.grad, dT[rank]), rel(s.grad, ds[rank]), rel(bb.grad, db[rank]))
    if e[0] > 1e-3 or max(e[1:]) > 2e-2:
        failures.append(f"siglip W={world} rank={rank}: loss {e[0]:.2e} dI {e[1]:.2e} dT {e[2]:.2e} ds {e[3]:.2e} db {e[4]:.2e}")
    n_fail = torch.tensor([len(failures)], device=dev)
    dist.all_reduce(n_fail)
    for f in failures:
        print("FAIL", f, flush=True)
    if rank == 0:
        print(f"dist_gpu_worker: world={world} total_failures={int(n_fail)}", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(1 if int(n_fail) else 0)


if __name__ == "__main__":
    main()
