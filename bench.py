#!/usr/bin/env python
"""bench.py -- headline benchmark of the XTag-CLIP contrastive head on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Metric (BASELINE.json): contrastive-loss fwd+bwd samples/s at global batch; % of bf16 tensor peak.
Workload: config 5 -- ViT-H-14 shape, GLOBAL batch 32768, embed dim 1024, bf16, one process per GPU, rank r owns
rows [r*b, (r+1)*b) with b = 32768/N (local_loss + gather_with_grad, the mode scripts/h14_224_32_finetune.sh of
the reference uses).  Total work is fixed as N grows => "scaling": "strong".
A step = ClipLoss(...)(image_features, text_features, logit_scale) + loss.backward() through the public drop-in
API of xtag_clip_b200 on synthetic, L2-normalised features.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "contrastive-loss fwd+bwd samples/s at global batch"
UNIT = "samples/s"
GLOBAL_BATCH = 32768
DIM = 1024
LOGIT_SCALE = 14.285714          # exp(2.659), the reference's init (model.py:263)
CPU_SHARDS = 8                   # bounded CPU sample: 1 of 8 row shards of the whole-batch problem


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return dict(burst=float(d["bf16_tflops"]), sustained=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])),
                    hbm=float(d["hbm_gbs"]), src="measured")
    return dict(burst=1590.0, sustained=1400.0, hbm=6650.0, src="fallback")


def _synth(rank: int, b: int, d: int, device, dtype):
    import torch
    g = torch.Generator().manual_seed(1234 + rank)
    i_raw = torch.randn(b, d, generator=g)
    t_raw = 0.5 * i_raw + 0.5 * torch.randn(b, d, generator=g)
    I = torch.nn.functional.normalize(i_raw, dim=-1).to(dtype)
    T = torch.nn.functional.normalize(t_raw, dim=-1).to(dtype)
    return I, T


# ------------------------------------------------------------------------------------------------------------
def workload_config(B: int, D: int, world: int):
    """`config` of the JSON line: names the workload only, so both arms (ours / --impl reference) print the same."""
    b = B // world
    return dict(workload=f"C5: ViT-H-14 contrastive head (ClipLoss fwd+bwd), global batch {B}, dim {D}",
                global_batch=B, dim=D, per_rank_batch=b,
                parallelism=f"dp{world} (rows sharded, local_loss + gather_with_grad)" if world > 1 else "dp1",
                l2=f"no flush: per-step working set (features + {b}x{B} logit-sized matrices) = "
                   f"{(2 * b * B * 2 * 2 + 4 * B * D * 2) / 2**20:.0f} MiB > 126 MB L2")


def cpu_shard_step_time(B: int, D: int, shards: int, iters: int, warmup: int):
    """The reference's algorithm on the host cores, bounded: ONE of `shards` row shards of the whole-batch problem --
    rows [0, B/shards) of both B x B logit matrices against all B columns, forward + backward, exactly what one rank of
    the reference's local_loss world computes (oracle.clip_loss_local_rank restates loss.py:95-96, 116-118, 134-137;
    the Python reference cannot travel to the GPU box).  The shards partition the rows, so a whole-batch step is
    `shards` such steps: no B^2 extrapolation.  fp32, every host thread."""
    import torch
    import oracle
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    I, T = _synth(0, B, D, "cpu", torch.float32)
    b = B // shards
    Il = I[:b].clone().requires_grad_(True)
    Tl = T[:b].clone().requires_grad_(True)
    s = torch.tensor(LOGIT_SCALE, requires_grad=True)
    times = []
    for it in range(warmup + iters):
        Il.grad = Tl.grad = s.grad = None
        t0 = time.perf_counter()
        loss = oracle.clip_loss_local_rank(Il, Tl, I, T, s, 0)
        loss.backward()
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    return times, cores, float(loss)


def cpu_full_step_time(B: int, D: int):
    """ONE real whole-batch step of the reference's W=1 path (oracle.clip_loss_single: two B x B logit matrices) when
    the host has the memory for it (~12 B x B fp32 matrices live at the peak); None otherwise."""
    import torch
    import oracle
    try:
        import psutil
        if psutil.virtual_memory().available < 14 * B * B * 4:
            return None
    except Exception:
        return None
    I, T = _synth(0, B, D, "cpu", torch.float32)
    I.requires_grad_(True)
    T.requires_grad_(True)
    s = torch.tensor(LOGIT_SCALE, requires_grad=True)
    t0 = time.perf_counter()
    oracle.clip_loss_single(I, T, s).backward()
    return time.perf_counter() - t0


def cpu_baseline_record(full_B: int, D: int, full_step: bool):
    times, cores, _ = cpu_shard_step_time(full_B, D, CPU_SHARDS, iters=3, warmup=1)
    t = statistics.median(times)
    value = full_B / (CPU_SHARDS * t)
    rec = dict(value=value, unit=UNIT, cores=cores, kind="port",
               sample=f"oracle port of the reference ClipLoss (fp32, torch CPU, {cores} threads): rows [0, {full_B // CPU_SHARDS})"
                      f" of both {full_B} x {full_B} logit matrices (1 of {CPU_SHARDS} row shards of the whole-batch step = one "
                      f"rank of the reference's local_loss world), fwd+bwd, median of 3: {t * 1e3:.0f} ms; whole-batch step = "
                      f"{CPU_SHARDS} shards = {CPU_SHARDS * t * 1e3:.0f} ms")
    if full_step:
        try:
            tf = cpu_full_step_time(full_B, D)
        except Exception as e:          # e.g. the host ran out of memory after all
            tf = None
            rec["full_batch_step_error"] = f"{type(e).__name__}: {e}"[:200]
        if tf is not None:
            rec["measured_full_batch_step_ms"] = tf * 1e3
            rec["measured_full_batch_samples_per_s"] = full_B / tf
    return rec


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path (the oracle port: the reference is Python
    and /root/reference does not exist on the GPU box) on this arm's config, metric and unit.  A step is one bounded
    sample of the workload (one of 8 row shards, see cpu_shard_step_time); rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    K, W = args.steps, args.warmup
    B, D = args.batch, args.dim
    times, cores, loss = cpu_shard_step_time(B, D, CPU_SHARDS, iters=max(K, 1), warmup=max(W, 1))
    t = sum(times) / len(times)
    value = B / (CPU_SHARDS * t)
    sample = (f"oracle port of the reference ClipLoss (fp32, torch CPU, {cores} threads): each step is fwd+bwd of rows "
              f"[0, {B // CPU_SHARDS}) of both {B} x {B} logit matrices against all {B} columns -- 1 of {CPU_SHARDS} row "
              f"shards of the whole-batch step, what one rank of the reference's local_loss world computes "
              f"({t * 1e3:.0f} ms mean of {len(times)}); ms_per_step and value are for the whole batch = {CPU_SHARDS} "
              f"shards (the shards partition the rows exactly)")
    line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=args.gpus, steps=K, warmup=W,
                ms_per_step=CPU_SHARDS * t * 1e3, higher_is_better=True, scaling="strong", vs_baseline=None,
                dtype="f32", data="synthetic", impl="reference",
                config=workload_config(B, D, max(args.gpus, 1)),
                cpu_baseline=dict(value=value, unit=UNIT, cores=cores, kind="port", sample=sample),
                e2e=dict(value=value, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                gpu_launches=0, loss_of_sample=loss)
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------
class ClockSampler:
    QUERY = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                       "-i", str(self.idx), "-lms", "200"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def wait_first_sample(self, timeout=20.0):
        """nvidia-smi takes a second or more to initialise (and holds driver locks while it does): the timed
        region must not start before the sampler is in steady state."""
        t0 = time.time()
        while self.p is not None and time.time() - t0 < timeout:
            if os.path.getsize(self.f.name) > 0:
                return
            time.sleep(0.05)

    @staticmethod
    def _ts(s):
        import datetime
        try:
            return datetime.datetime.strptime(s.strip(), "%Y/%m/%d %H:%M:%S.%f").timestamp()
        except Exception:
            return None

    def stop(self, t_lo=None, t_hi=None):
        if self.p is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(",") for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons, power = [], [], set(), []
        for r in rows:
            try:
                r = [x.strip() for x in r]
                ts = self._ts(r[0])
                if t_lo is not None and ts is not None and not (t_lo - 0.25 <= ts <= t_hi + 0.25):
                    continue                      # keep only samples taken during the timed region
                sm.append(float(r[1])); mx.append(float(r[2])); power.append(float(r[3]))
                names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
                for n, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        if not sm:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["no samples inside the timed region"])
        loaded = [s for s, p in zip(sm, power) if p >= 0.5 * max(power)] or sm
        return dict(sm_mhz=statistics.median(loaded), sm_max_mhz=max(mx), reasons=sorted(reasons),
                    samples=len(sm), power_w_max=max(power))


def _bind_to_gpu_numa_node(local_rank: int):
    """Pin this process to the CPUs of the GPU's NUMA node BEFORE any pinned host memory is allocated: with 8 ranks
    on one host the per-step H2D copies otherwise cross the socket interconnect (measured: 23 GB/s per GPU with all 8
    copying at once).  Best effort; returns what it did for the JSON line."""
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = int(vis.split(",")[local_rank]) if vis and vis.split(",")[local_rank].strip().isdigit() else local_rank
        h = pynvml.nvmlDeviceGetHandleByIndex(phys)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        dom, rest = bus.split(":", 1)
        bdf = f"{int(dom, 16):04x}:{rest.lower()}"
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read().strip())
        if node < 0:
            return dict(numa_node=node, bound=False)
        cpus = []
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.extend(range(int(lo), int(hi or lo) + 1))
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if not allowed:
            return dict(numa_node=node, bound=False)
        os.sched_setaffinity(0, allowed)
        try:                                   # prefer memory of that node for the pinned buffers allocated next
            import ctypes as _ct
            libnuma = _ct.CDLL("libnuma.so.1")
            if libnuma.numa_available() >= 0:
                libnuma.numa_set_preferred(node)
        except Exception:
            pass
        return dict(numa_node=node, cpus=len(allowed), bound=True)
    except Exception as e:
        return dict(bound=False, why=f"{type(e).__name__}: {e}"[:120])


def _ncu_traffic(kernel_sig: str):
    """DRAM bytes (read + write) per launch of the kernel whose name contains `kernel_sig`, from the newest committed
    `ncu --set full` raw page under profiles/ (1 GPU, config 5).  -> (bytes | None, file | None)"""
    import csv
    import glob
    mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    for f in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_ncu_full*raw.csv")), reverse=True):
        try:
            rows = list(csv.reader(open(f)))
            hdr, units = rows[0], rows[1]
            kn, rd, wr = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
            vals = [float(r[rd]) * mult[units[rd]] + float(r[wr]) * mult[units[wr]]
                    for r in rows[2:] if len(r) > max(rd, wr) and kernel_sig in r[kn]]
            if vals:
                return sum(vals) / len(vals), os.path.relpath(f, ROOT)
        except Exception:
            continue
    return None, None


def _time_steps(fn, iters, warmup=3):
    import torch
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


def torch_b200_record(I_dev, T_dev, log_scale):
    """Stock PyTorch on the SAME GPU in the SAME run: the reference's ClipLoss expressions (loss.py:116-137: two
    logit GEMMs, F.cross_entropy on each) under torch.autocast(bf16), eager -- the number the kernels have to beat."""
    import torch
    import torch.nn.functional as F
    B = I_dev.shape[0]
    labels = torch.arange(B, device=I_dev.device)

    def step():
        I_dev.grad = T_dev.grad = log_scale.grad = None
        with torch.autocast("cuda", dtype=torch.bfloat16):
            s = log_scale.exp()
            li = s * I_dev @ T_dev.T
            lt = s * T_dev @ I_dev.T
            loss = (F.cross_entropy(li, labels) + F.cross_entropy(lt, labels)) / 2
        loss.backward()
        return loss

    try:
        ms = _time_steps(step, iters=5, warmup=2)
        return dict(ms_per_step=ms, value=B / ms * 1e3, unit=UNIT, loss=float(step()),
                    what="stock torch eager ClipLoss math under bf16 autocast, same GPU, same inputs")
    except Exception as e:
        return dict(error=f"{type(e).__name__}: {e}"[:200])
    finally:
        I_dev.grad = T_dev.grad = log_scale.grad = None
        torch.cuda.empty_cache()


def config_subrecords(dev, world, rank):
    """The other BASELINE configs as sub-records (1 GPU: C2 sweep and the C3 tag head; N > 1: C4 at this world size)."""
    import torch
    import torch.distributed as dist
    import xtag_clip_b200 as xt
    out = {}
    if world == 1:
        c2 = []
        for B in (256, 512, 1024, 2048, 4096):
            I, T = _synth(0, B, 512, "cpu", torch.bfloat16)
            I = I.to(dev).requires_grad_(True)
            T = T.to(dev).requires_grad_(True)
            ls = torch.tensor(2.659260036932778, device=dev, requires_grad=True)
            rec = dict(B=B, D=512)
            for name, mod in (("graph", xt.ClipLoss(cuda_graph=True)), ("eager", xt.ClipLoss())):
                def step():
                    I.grad = T.grad = ls.grad = None
                    mod(I, T, ls.exp()).backward()
                ms = _time_steps(step, iters=50, warmup=5)
                rec[f"{name}_ms"] = ms
            rec["samples_per_s"] = B / min(rec["graph_ms"], rec["eager_ms"]) * 1e3
            c2.append(rec)
        out["C2_vit_b32_head_sweep_bf16"] = c2
        try:
            B, D = GLOBAL_BATCH, DIM
            I, T = _synth(0, B, D, "cpu", torch.bfloat16)
            I = I.to(dev).requires_grad_(True)
            T = T.to(dev).requires_grad_(True)
            ls = torch.tensor(2.3, device=dev, requires_grad=True)
            lb = torch.tensor(-10.0, device=dev, requires_grad=True)
            sig = xt.SigLipLoss()

            def step_sig():
                I.grad = T.grad = ls.grad = lb.grad = None
                sig(I, T, ls.exp(), lb).backward()

            ms = _time_steps(step_sig, iters=10, warmup=3)
            out["siglip_C5_shape"] = dict(B=B, D=D, ms_per_step=ms, samples_per_s=B / ms * 1e3,
                                          algorithmic_tflops=6.0 * B * B * D / (ms * 1e-3) / 1e12,
                                          what="SigLipLoss fwd+bwd (one pass: loss + staged logit gradient; two gradient "
                                               "GEMMs): 6*B^2*D executed = algorithmic")
            del I, T
            torch.cuda.empty_cache()
        except Exception as e:
            out["siglip_C5_shape"] = dict(error=f"{type(e).__name__}: {e}"[:200])
        try:
            b, N, D = 1024, 197, 512
            head = xt.TagHead(D).to(dev).train()
            asl = xt.AsymmetricLoss(gamma_neg=4, gamma_pos=1, clip=0.05)
            tok = torch.randn(b, N, D, device=dev, dtype=torch.bfloat16, requires_grad=True)
            y = (torch.rand(b, 22, device=dev) > 0.7).float().repeat(1, 2)

            def step_head():
                tok.grad = None
                head.zero_grad(set_to_none=True)
                with torch.autocast("cuda", dtype=torch.bfloat16):
                    logits = head.tag_forward(tok)
                asl(logits.float(), y).backward()

            ms = _time_steps(step_head, iters=10, warmup=3)
            # the same head on stock PyTorch ops only (the reference's formulation: one Linear per projection, eager
            # softmax attention with its [b, 4, 44, N] score tensor, separate dropout / add / LayerNorm kernels)
            import math
            import torch.nn.functional as F
            from xtag_clip_b200 import tag_head as th

            def eager_attention(q, k, v, heads, dropout_p=0.0, seed=0, offset=0, **kw):
                bb, Lq, H = q.shape
                dh = H // heads
                qh = q.reshape(bb, Lq, heads, dh).permute(0, 2, 1, 3)
                kh = k.reshape(bb, -1, heads, dh).permute(0, 2, 1, 3)
                vh = v.reshape(bb, -1, heads, dh).permute(0, 2, 1, 3)
                pr = torch.softmax((qh @ kh.transpose(-1, -2)) / math.sqrt(dh), dim=-1)
                pr = F.dropout(pr, dropout_p, dropout_p > 0)
                return (pr @ vh).permute(0, 2, 1, 3).reshape(bb, Lq, H)

            orig, flags = th.cross_attention, (head.fuse_kv, head.fuse_ln)
            th.cross_attention, head.fuse_kv, head.fuse_ln = eager_attention, False, False
            try:
                ms_lib = _time_steps(step_head, iters=10, warmup=3)
            finally:
                th.cross_attention = orig
                head.fuse_kv, head.fuse_ln = flags
            out["C3_tag_head_fwd_bwd_train_bf16"] = dict(b=b, N=N, D=D, ms_per_step=ms, samples_per_s=b / ms * 1e3,
                                                         torch_b200_ms_per_step=ms_lib, speedup_vs_torch_b200=ms_lib / ms,
                                                         what="TagHead.tag_forward (2 cross-attention layers, dropout 0.1)"
                                                              " + AsymmetricLoss, fwd+bwd, bf16 autocast; torch_b200 = the "
                                                              "same head on stock PyTorch ops only, same GPU")
        except Exception as e:
            out["C3_tag_head_fwd_bwd_train_bf16"] = dict(error=f"{type(e).__name__}: {e}"[:200])
    else:
        B, D = 8192, 768
        b = B // world
        I, T = _synth(rank, b, D, "cpu", torch.bfloat16)
        I = I.to(dev).requires_grad_(True)
        T = T.to(dev).requires_grad_(True)
        ls = torch.tensor(2.659260036932778, device=dev, requires_grad=True)
        mod = xt.ClipLoss(local_loss=True, gather_with_grad=True, rank=rank, world_size=world, cuda_graph=True)

        def step():
            I.grad = T.grad = ls.grad = None
            mod(I, T, ls.exp()).backward()

        for _ in range(5):
            step()
        dist.barrier()
        ms = _time_steps(step, iters=50, warmup=3)
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t)
        out["C4_vit_l16_global8192_d768"] = dict(B=B, D=D, world=world, per_rank_batch=b, ms_per_step=ms,
                                                 samples_per_s=B / ms * 1e3,
                                                 algorithmic_tflops_per_gpu=6.0 * B * B * D / world / (ms * 1e-3) / 1e12,
                                                 path=mod.last_path)
    return out


def run_ours(args):
    from xtag_clip_b200._cuda_probe import wait_for_cuda
    wait_for_cuda()
    import torch
    import torch.distributed as dist
    import xtag_clip_b200 as xt
    from xtag_clip_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torchrun for --gpus > 1")
    assert torch.cuda.is_available(), "bench.py --impl ours needs a CUDA device (there is no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = _bind_to_gpu_numa_node(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, D = args.batch, args.dim
    assert B % world == 0
    b = B // world
    K, W = args.steps, max(args.warmup, 3)
    lib = _lib.load()

    I_host, T_host = _synth(rank, b, D, "cpu", torch.bfloat16)
    I_pin, T_pin = I_host.pin_memory(), T_host.pin_memory()
    I_dev = I_pin.to(dev).requires_grad_(True)
    T_dev = T_pin.to(dev).requires_grad_(True)
    log_scale = torch.tensor(2.659260036932778, device=dev, requires_grad=True)     # ln(1/0.07)
    # cuda_graph=True: the public option that replays a captured step (the 8-GPU shard is launch-bound otherwise)
    loss_mod = xt.ClipLoss(local_loss=world > 1, gather_with_grad=world > 1, cache_labels=True, rank=rank,
                           world_size=world, cuda_graph=not args.no_graph, pull_streams=args.pull_streams,
                           exchange=args.exchange)
    loss_eager = xt.ClipLoss(local_loss=world > 1, gather_with_grad=world > 1, cache_labels=True, rank=rank,
                             world_size=world, pull_streams=args.pull_streams, exchange=args.exchange)

    def step(I, T, mod=None):
        I.grad = T.grad = log_scale.grad = None
        loss = (mod or loss_mod)(I, T, log_scale.exp())
        loss.backward()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
        sampler.wait_first_sample()
    for _ in range(W):
        loss = step(I_dev, T_dev)
    barrier()
    # the fast path must be the one that is timed: a silent degradation (graph capture refused -> eager launches,
    # symmetric memory unavailable -> NCCL P2P) would still print a number
    path = loss_mod.last_path
    if not args.no_graph and not path["cuda_graph"]:
        raise SystemExit(f"bench.py: CUDA-graph capture fell back to eager launches ({path}); refusing to time it")
    if world > 1 and not str(path["exchange"]).startswith("symmetric-memory"):
        raise SystemExit(f"bench.py: the symmetric-memory exchange fell back to {path['exchange']}; refusing to time it")

    # Device-resident leg: the features live in the captured step's own input slots (the API a producer uses to write
    # its normalised features in place: ClipLoss.graph_input_slots), so no per-step input copy is timed
    slots = None if args.no_graph else loss_mod.graph_input_slots(b, D, torch.bfloat16)
    if slots is not None:
        slots[0].copy_(I_dev.detach())
        slots[1].copy_(T_dev.detach())
        I_dev = slots[0].requires_grad_(True)
        T_dev = slots[1].requires_grad_(True)
        for _ in range(2):
            step(I_dev, T_dev)
        barrier()

    # ---- device-resident leg (value) ----
    n0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_lo = time.time()
    e0.record()
    for _ in range(K):
        loss = step(I_dev, T_dev)
    e1.record()
    barrier()
    t_hi = time.time()
    ms = e0.elapsed_time(e1)
    launches = _lib.launch_count() - n0
    # Sustained window: the SAME steps for >= 1.2 s more (on every rank -- the steps contain collectives).  It yields
    # the clock samples of this exact load (the timed region above can be shorter than the 200 ms sampling period)
    # and a second throughput figure taken in the settled power state, comparable between N = 1 and N = 8.
    ms_max = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms_max, op=dist.ReduceOp.MAX)
    n_sus = int(1200.0 / max(float(ms_max) / K, 1e-3)) + 1
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    for _ in range(n_sus):
        step(I_dev, T_dev)
    s1.record()
    barrier()
    t_hi = time.time()
    ms_sus = s0.elapsed_time(s1)
    clocks = sampler.stop(t_lo, t_hi) if sampler else None
    loss_val = float(loss.item())

    # ---- end-to-end leg: host buffers in, loss out, copies inside the timed region ----
    # Every step copies ITS inputs from pinned host memory and reads its loss back.  Like any input pipeline the
    # copy of step i+1 is issued on a copy stream before step i's result is awaited, so H2D overlaps compute.
    copy_stream = torch.cuda.Stream(device=dev)
    bufs = [(torch.empty_like(I_dev), torch.empty_like(T_dev)) for _ in range(2)]

    def h2d(slot):
        with torch.cuda.stream(copy_stream):
            bufs[slot][0].detach().copy_(I_pin, non_blocking=True)
            bufs[slot][1].detach().copy_(T_pin, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return ev

    loss_pin = torch.zeros(2, dtype=torch.float32).pin_memory()

    def e2e_loop(n):
        """Software-pipelined like a training input pipeline with asynchronous loss logging: while step i runs on the
        GPU, the H2D copy of step i+1's inputs is in flight on the copy stream and the host reads step i-1's loss
        (its D2H copy into pinned memory was enqueued right behind step i-1).  Every step's inputs cross PCIe and
        every step's loss is read back inside the timed region; the host never idles the GPU between steps."""
        for bi, bt in bufs:
            bi.requires_grad_(True)
            bt.requires_grad_(True)
        main = torch.cuda.current_stream()
        ev = h2d(0)
        last, done, read_ev = 0.0, None, [None, None]
        for i in range(n):
            main.wait_event(ev)
            cur = i & 1
            if i + 1 < n:
                if done is not None:
                    copy_stream.wait_event(done)                       # slot cur^1 was last read by step i-1
                ev = h2d(cur ^ 1)                                      # next step's inputs, overlapping step i
            loss_t = step(bufs[cur][0], bufs[cur][1])
            loss_pin[cur].copy_(loss_t.detach(), non_blocking=True)    # D2H read of this step's result
            done = torch.cuda.Event()
            done.record(main)
            read_ev[cur] = done
            if read_ev[cur ^ 1] is not None:                           # consume step i-1's loss while step i runs
                read_ev[cur ^ 1].synchronize()
                last = float(loss_pin[cur ^ 1])
        read_ev[(n - 1) & 1].synchronize()
        return float(loss_pin[(n - 1) & 1])

    e2e_loop(2)
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    e2e_loop(K)
    f1.record()
    barrier()
    ms_e2e = f0.elapsed_time(f1)
    # what the H2D copies alone cost when every rank copies at once (diagnostic for the e2e figure at N > 1)
    barrier()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record(copy_stream)
    for i in range(10):
        h2d(i & 1)
    g1.record(copy_stream)
    barrier()
    ms_h2d = g0.elapsed_time(g1) / 10

    # ---- per-kernel pass for the roofline (CUDA events recorded inside the library on the launch stream) ----
    # (eager module: the library's per-launch events are recorded at launch time, a graph replay launches nothing
    #  from the host)
    for _ in range(2):
        step(I_dev, T_dev, loss_eager)
    barrier()
    lib.xtag_prof_enable(1)
    for _ in range(K):
        step(I_dev, T_dev, loss_eager)
    torch.cuda.synchronize()
    cap = 64 * K + 64
    tags = (ctypes.c_int * cap)()
    tms = (ctypes.c_float * cap)()
    work = (ctypes.c_double * cap)()
    n = lib.xtag_prof_read(tags, tms, work, cap)
    lib.xtag_prof_enable(0)
    per = {}
    for i in range(n):
        per.setdefault(tags[i], []).append((tms[i], work[i]))
    barrier()

    t_ms = torch.tensor([ms, ms_e2e, ms_sus, ms_h2d], device=dev, dtype=torch.float64)
    l_cnt = torch.tensor([launches], device=dev, dtype=torch.int64)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(l_cnt, op=dist.ReduceOp.SUM)
    ms, ms_e2e, ms_sus, ms_h2d = (float(x) for x in t_ms)

    extra = {}
    if not args.no_extras:
        try:
            extra = config_subrecords(dev, world, rank)
        except Exception as e:
            extra = dict(error=f"{type(e).__name__}: {e}"[:300])
        if world == 1:
            extra["torch_b200"] = torch_b200_record(I_dev, T_dev, log_scale)

    if rank == 0:
        peaks = _peaks()
        value = B * K / (ms * 1e-3)
        e2e_value = B * K / (ms_e2e * 1e-3)
        names = {0: "tc_gemm_kernel<EPI_LSE> (K1 fused forward)", 1: "tc_gemm_kernel<EPI_DS> (K2 dS producer)",
                 2: "tc_gemm_kernel<EPI_STORE> (K2 dI/dT GEMM)"}
        sigs = {0: "tc_gemm_kernel<0,", 1: "tc_gemm_kernel<1,", 2: "tc_gemm_kernel<2,"}
        kern = {}
        for tag, recs in per.items():
            tot_ms = sum(r[0] for r in recs)
            tot_w = sum(r[1] for r in recs)
            tf = tot_w / (tot_ms * 1e-3) / 1e12 if tot_ms > 0 else None
            kern[names.get(tag, str(tag))] = dict(launches_per_step=len(recs) / K, avg_ms=tot_ms / len(recs),
                                                  share_of_step=tot_ms / K / (ms / K), tflops=tf,
                                                  frac_of_sustained_peak=tf / peaks["sustained"] if tf else None,
                                                  frac_of_burst_peak=tf / peaks["burst"] if tf else None)

        def roof(tag):
            recs = per[tag]
            ach = sum(r[1] for r in recs) / (sum(r[0] for r in recs) * 1e-3) / 1e12
            traffic, src = _ncu_traffic(sigs.get(tag, "?")) if (world == 1 and B == GLOBAL_BATCH and D == DIM) else (None, None)
            return dict(bound="tensor", kernel=names.get(tag), achieved=ach, peak=peaks["sustained"], unit="TFLOP/s",
                        frac=ach / peaks["sustained"], frac_of_burst_peak=ach / peaks["burst"], traffic=traffic,
                        traffic_src=f"{src}: dram__bytes_read.sum + dram__bytes_write.sum per launch" if src else None,
                        peak_kind=f"bf16 sustained, {peaks['src']} (kernel timed inside a long step); burst peak "
                                  f"{peaks['burst']}",
                        flops_per_launch=sum(r[1] for r in recs) / len(recs),
                        avg_launch_ms=sum(r[0] for r in recs) / len(recs))

        roofline = roofline_longest = None
        if per:
            dom_tag = max(per, key=lambda t: sum(r[0] for r in per[t]))                 # largest share of the step
            long_tag = max(per, key=lambda t: sum(r[0] for r in per[t]) / len(per[t]))  # longest single launch
            roofline = roof(dom_tag)
            roofline_longest = roof(long_tag)
        step_tflops_per_gpu = 6.0 * B * B * D / world / (ms / K * 1e-3) / 1e12
        sus_tflops_per_gpu = 6.0 * B * B * D / world / (ms_sus / n_sus * 1e-3) / 1e12
        cpu = cpu_baseline_record(B, D, full_step=not args.no_cpu_full) if world == 1 and not args.no_cpu else None
        line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=K, warmup=W, ms_per_step=ms / K,
                    higher_is_better=True, scaling="strong", vs_baseline=None, dtype="bf16", data="synthetic",
                    config=workload_config(B, D, world),
                    run=dict(launch="eager" if args.no_graph else "CUDA-graph replay (ClipLoss(cuda_graph=True))",
                             path=path, numa=numa, inputs_in_graph_slots=slots is not None,
                             exchange=None if world == 1 else
                             "feature all-gather, column-LSE combine and dT reduce-scatter by copy engines / kernels over "
                             "peer-mapped symmetric memory (torch.distributed._symmetric_memory); the NCCL process group "
                             "only carries the rendezvous and bench.py's barriers"),
                    clocks=clocks,
                    e2e=dict(value=e2e_value, unit=UNIT, h2d_bytes_per_step=2 * b * D * 2 * world,
                             d2h_bytes_per_step=4 * world, ms_per_step=ms_e2e / K, h2d_only_ms_per_step=ms_h2d),
                    gpu_launches=int(l_cnt[0]),
                    roofline=roofline, roofline_longest_launch=roofline_longest,
                    tensor_frac_of_step=dict(algorithmic_tflops_per_gpu=step_tflops_per_gpu,
                                             frac_of_burst_peak=step_tflops_per_gpu / peaks["burst"],
                                             frac_of_sustained_peak=step_tflops_per_gpu / peaks["sustained"],
                                             note="6*B^2*D / (W * t_step); backward recompute not credited"),
                    sustained=dict(steps=n_sus, window_s=ms_sus * 1e-3, ms_per_step=ms_sus / n_sus,
                                   value=B * n_sus / (ms_sus * 1e-3), unit=UNIT,
                                   algorithmic_tflops_per_gpu=sus_tflops_per_gpu,
                                   frac_of_burst_peak=sus_tflops_per_gpu / peaks["burst"],
                                   frac_of_sustained_peak=sus_tflops_per_gpu / peaks["sustained"],
                                   note="same steps, run back to back for >= 1.2 s right after the timed region"),
                    kernels=kern, loss=loss_val, **({"configs": extra} if extra else {}))
        if cpu:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=int(os.environ.get("WORLD_SIZE", "1")))
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=GLOBAL_BATCH, help="global batch (default: BASELINE config 5)")
    ap.add_argument("--dim", type=int, default=DIM)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
    ap.add_argument("--no-cpu-full", action="store_true", help="skip the one real whole-batch CPU step of cpu_baseline")
    ap.add_argument("--no-extras", action="store_true", help="skip the C2 / C3 / C4 / torch_b200 sub-records")
    ap.add_argument("--no-graph", action="store_true", help="eager launches instead of CUDA-graph replay")
    ap.add_argument("--pull-streams", type=int, default=2,
                    help="copy streams of the streamed feature exchange (see symm.gather_streamed / gather_pushed)")
    ap.add_argument("--exchange", default=None, choices=["pull", "push"],
                    help="feature exchange of the sharded forward (default: the library default / XTAG_EXCHANGE)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
