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
CPU_SAMPLE_B = 4096              # bounded CPU sample: the C5 shape at 1/8 of the batch


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
def cpu_reference_step_time(B: int, D: int, iters: int = 5, warmup: int = 2):
    """The reference's own algorithm for this path on the host cores: oracle/clip_oracle.py (a restatement of
    src/open_clip/loss.py:104-139; the Python reference itself cannot travel to the GPU box), fp32, all threads."""
    import torch
    import oracle
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    I, T = _synth(0, B, D, "cpu", torch.float32)
    I.requires_grad_(True)
    T.requires_grad_(True)
    s = torch.tensor(LOGIT_SCALE, requires_grad=True)
    best = float("inf")
    for it in range(warmup + iters):
        I.grad = T.grad = s.grad = None
        t0 = time.perf_counter()
        loss = oracle.clip_loss_single(I, T, s)
        loss.backward()
        dt = time.perf_counter() - t0
        if it >= warmup:
            best = min(best, dt)
    return best, cores


def cpu_baseline_record(full_B: int):
    t, cores = cpu_reference_step_time(CPU_SAMPLE_B, DIM)
    raw = CPU_SAMPLE_B / t
    # per-step cost of this path is proportional to B^2 (B x B logits), so samples/s at the full batch is
    # raw * (B_sample / B_full)
    value = raw * CPU_SAMPLE_B / full_B
    return dict(value=value, unit=UNIT, cores=cores, kind="port",
                sample=f"oracle ClipLoss fwd+bwd fp32, B={CPU_SAMPLE_B} D={DIM}, best of 5: {t * 1e3:.1f} ms/step "
                       f"= {raw:.0f} samples/s at B={CPU_SAMPLE_B}; scaled by B_sample/B_full (cost ~ B^2) to "
                       f"global batch {full_B}")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    K, W = args.steps, args.warmup
    t, cores = cpu_reference_step_time(CPU_SAMPLE_B, DIM, iters=max(K, 1), warmup=max(W, 1))
    raw = CPU_SAMPLE_B / t
    value = raw * CPU_SAMPLE_B / GLOBAL_BATCH
    ms_full = t * 1e3 * (GLOBAL_BATCH / CPU_SAMPLE_B) ** 2
    sample = (f"oracle port of the reference ClipLoss (fp32, torch CPU, {cores} threads): each step is fwd+bwd at "
              f"B={CPU_SAMPLE_B}, D={DIM} ({t * 1e3:.1f} ms, {raw:.0f} samples/s); value is scaled to the global batch "
              f"{GLOBAL_BATCH} by B_sample/B_full because the step cost grows as B^2")
    line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=args.gpus, steps=K, warmup=W, ms_per_step=ms_full,
                higher_is_better=True, scaling="strong", vs_baseline=None, dtype="f32", data="synthetic",
                impl="reference",
                config=dict(workload="C5: ViT-H-14 contrastive head, global batch 32768, dim 1024 (CPU: bounded sample)",
                            global_batch=GLOBAL_BATCH, dim=DIM),
                cpu_baseline=dict(value=value, unit=UNIT, cores=cores, kind="port", sample=sample),
                e2e=dict(value=value, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                gpu_launches=0)
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
                             world_size=world)

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
    # a timed region shorter than the 200 ms sampling period yields no clock sample: keep the SAME steps running
    # (untimed, on every rank -- the steps contain collectives) until a few samples under this exact load exist
    ms_max = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms_max, op=dist.ReduceOp.MAX)
    if float(ms_max) < 600.0:
        extra = int(800.0 / max(float(ms_max) / K, 1e-3)) + 1
        for _ in range(extra):
            step(I_dev, T_dev)
        barrier()
        t_hi = time.time()
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

    t_ms = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
    l_cnt = torch.tensor([launches], device=dev, dtype=torch.int64)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(l_cnt, op=dist.ReduceOp.SUM)
    ms, ms_e2e = float(t_ms[0]), float(t_ms[1])

    if rank == 0:
        peaks = _peaks()
        value = B * K / (ms * 1e-3)
        e2e_value = B * K / (ms_e2e * 1e-3)
        names = {0: "tc_gemm_kernel<EPI_LSE> (K1 fused forward)", 1: "tc_gemm_kernel<EPI_DS> (K2 dS producer)",
                 2: "tc_gemm_kernel<EPI_STORE> (K2 dI/dT GEMM)"}
        kern = {}
        for tag, recs in per.items():
            tot_ms = sum(r[0] for r in recs)
            tot_w = sum(r[1] for r in recs)
            kern[names.get(tag, str(tag))] = dict(launches_per_step=len(recs) / K, avg_ms=tot_ms / len(recs),
                                                  share_of_step=tot_ms / K / (ms / K),
                                                  tflops=tot_w / (tot_ms * 1e-3) / 1e12 if tot_ms > 0 else None)
        dom_tag = max(per, key=lambda t: sum(r[0] for r in per[t])) if per else None
        roofline = None
        if dom_tag is not None:
            recs = per[dom_tag]
            ach = sum(r[1] for r in recs) / (sum(r[0] for r in recs) * 1e-3) / 1e12
            # DRAM bytes per launch of this kernel from the committed `ncu --set full` capture (1 GPU, this config)
            traffic = 3.32e9 if (dom_tag == 2 and world == 1 and B == GLOBAL_BATCH and D == DIM) else None
            roofline = dict(bound="tensor", kernel=names.get(dom_tag), achieved=ach, peak=peaks["sustained"],
                            unit="TFLOP/s", frac=ach / peaks["sustained"], traffic=traffic,
                            traffic_src="profiles/r1_ncu_full_final_raw.csv (dram__bytes_read+write per launch; "
                                        "algorithmic 2.27e9)" if traffic else None,
                            peak_kind=f"bf16 sustained, {peaks['src']} (kernel timed inside a long step); "
                                      f"burst peak {peaks['burst']}",
                            flops_per_launch=sum(r[1] for r in recs) / len(recs),
                            avg_launch_ms=sum(r[0] for r in recs) / len(recs))
        step_tflops_per_gpu = 6.0 * B * B * D / world / (ms / K * 1e-3) / 1e12
        cpu = cpu_baseline_record(B) if world == 1 and not args.no_cpu else None
        line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=K, warmup=W, ms_per_step=ms / K,
                    higher_is_better=True, scaling="strong", vs_baseline=None, dtype="bf16", data="synthetic",
                    config=dict(workload=f"C5: ViT-H-14 contrastive head (ClipLoss fwd+bwd), global batch {B}, dim {D}, "
                                         f"{'local_loss+gather_with_grad, ' if world > 1 else ''}bf16",
                                global_batch=B, dim=D, per_rank_batch=b, parallelism=f"dp{world} (rows sharded)",
                                launch="eager" if args.no_graph else "CUDA-graph replay (ClipLoss(cuda_graph=True))",
                                l2=f"no flush: per-step working set (features + {b}x{B} bf16 dS x2) = "
                                   f"{(2 * b * B * 2 * 2 + 4 * B * D * 2) / 2**20:.0f} MiB > 126 MB L2"),
                    clocks=clocks,
                    e2e=dict(value=e2e_value, unit=UNIT, h2d_bytes_per_step=2 * b * D * 2 * world,
                             d2h_bytes_per_step=4 * world, ms_per_step=ms_e2e / K),
                    gpu_launches=int(l_cnt[0]),
                    roofline=roofline,
                    tensor_frac_of_step=dict(algorithmic_tflops_per_gpu=step_tflops_per_gpu,
                                             frac_of_burst_peak=step_tflops_per_gpu / peaks["burst"],
                                             frac_of_sustained_peak=step_tflops_per_gpu / peaks["sustained"],
                                             note="6*B^2*D / (W * t_step); backward recompute not credited"),
                    kernels=kern, loss=loss_val)
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
