"""world_size 2 and 3 over gloo on CPU: the multi-rank host logic of xtag_clip_b200.ClipLoss (collectives, label
offsets, gradient weights per mode) against what the REFERENCE produced on each rank (tests/golden/clip_dist.npz,
generated under gloo by oracle/make_golden.py).  Kernels are the contract model (tests/kernel_model.py)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    import torch.distributed as dist
    import xtag_clip_b200 as xt
    from kernel_model import ModelKernels
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    g = np.load(os.path.join(HERE, "golden", "clip_dist.npz"))
    pre = f"w{world}_"
    b = int(g[pre + "b"])
    res = {}
    try:
        for ll in (False, True):
            for gwg in (False, True):
                I = torch.from_numpy(g[pre + "I"])[rank * b:(rank + 1) * b].clone().requires_grad_(True)
                T = torch.from_numpy(g[pre + "T"])[rank * b:(rank + 1) * b].clone().requires_grad_(True)
                s = torch.tensor(float(g[pre + "scale"]), dtype=torch.float64, requires_grad=True)
                k = ModelKernels()
                mod = xt.ClipLoss(local_loss=ll, gather_with_grad=gwg, cache_labels=True, rank=rank,
                                  world_size=world, _kernels=k)
                loss = mod(I, T, s)
                loss.backward()
                key = f"ll{int(ll)}_gwg{int(gwg)}_"
                res[key + "loss"] = loss.item()
                res[key + "dI"] = I.grad.numpy()
                res[key + "dT"] = T.grad.numpy()
                res[key + "ds"] = float(s.grad)
                res[key + "calls"] = list(k.calls)
        q.put((rank, res, None))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, None, traceback.format_exc()))
    dist.barrier()
    dist.destroy_process_group()


def rel_err(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


@pytest.mark.parametrize("world,port", [(2, 29721), (3, 29722)])
def test_cliploss_multirank_matches_reference(world, port, golden_dir):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = {}
    for _ in range(world):
        rank, res, err = q.get(timeout=240)
        assert err is None, err
        got[rank] = res
    for p in procs:
        p.join(timeout=60)
    g = np.load(os.path.join(golden_dir, "clip_dist.npz"))
    pre = f"w{world}_"
    b = int(g[pre + "b"])
    for ll in (0, 1):
        for gwg in (0, 1):
            key = f"ll{ll}_gwg{gwg}_"
            for r in range(world):
                ref = pre + f"ll{ll}_gwg{gwg}_r{r}_"
                assert rel_err(got[r][key + "loss"], g[ref + "loss"]) < 1e-5
                assert rel_err(got[r][key + "dI"], g[ref + "dI"]) < 1e-5, (ll, gwg, r)
                assert rel_err(got[r][key + "dT"], g[ref + "dT"]) < 1e-5, (ll, gwg, r)
                if not (ll and gwg):
                    assert rel_err(got[r][key + "ds"], g[ref + "dscale"]) < 2e-4, (ll, gwg, r)   # cancelling sum on fp32 LSEs
            # documented deviation (loss.py docstring): per-rank d(logit_scale) differs in the fused
            # local_loss+gather_with_grad mode, its sum over ranks (what DDP reduces) is identical
            tot = sum(got[r][key + "ds"] for r in range(world))
            ref_tot = sum(float(g[pre + f"ll{ll}_gwg{gwg}_r{r}_dscale"]) for r in range(world))
            assert rel_err(tot, ref_tot) < 2e-4
    # the performance mode (local_loss + gather_with_grad) is chunk-pipelined: one forward call per arriving column
    # block -- the rank's own block first, carrying the labels; the other blocks with label_offset -1 -- on the
    # [b, G*b] row block only (no B x B work, no second logits matrix); the backward is the dB call followed by the
    # dA call that reuses the staged dS while the reduce-scatter runs
    D = g[pre + "I"].shape[1]
    G = 2 if (world >= 4 and world % 2 == 0) else 1
    for r in range(world):
        calls = got[r]["ll1_gwg1_calls"]
        fwd = [c for c in calls if c[0] == "clip_fwd"]
        assert len(fwd) == world // G
        assert fwd[0] == ("clip_fwd", (b, D), (G * b, D), (r % G) * b)
        assert all(c == ("clip_fwd", (b, D), (G * b, D), -1) for c in fwd[1:])
        bwd = [c for c in calls if c[0] == "clip_bwd"]
        assert len(bwd) == 2 and all(c[1] == (b, D) and c[2] == (world * b, D) and c[3] == r * b for c in bwd)


def _worker4(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    import torch.distributed as dist
    import oracle
    import xtag_clip_b200 as xt
    from kernel_model import ModelKernels
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    try:
        b, D, scale = 5, 16, 20.0
        gen = torch.Generator().manual_seed(99)
        I_all = torch.nn.functional.normalize(torch.randn(world * b, D, generator=gen, dtype=torch.float64), dim=-1)
        T_all = torch.nn.functional.normalize(0.3 * I_all + 0.7 * torch.randn(world * b, D, generator=gen,
                                                                              dtype=torch.float64), dim=-1)
        Il = [I_all[r * b:(r + 1) * b] for r in range(world)]
        Tl = [T_all[r * b:(r + 1) * b] for r in range(world)]
        errs = []
        # pipelined with deferred block reductions, pipelined with per-block reductions, not pipelined
        for pipeline, fwd_blocks in ((True, True), (False, True), (True, False)):
            losses, dI, dT, ds = oracle.clip_loss_world(Il, Tl, scale, True, True)
            I = Il[rank].clone().requires_grad_(True)
            T = Tl[rank].clone().requires_grad_(True)
            s = torch.tensor(scale, dtype=torch.float64, requires_grad=True)
            k = ModelKernels(fwd_blocks=fwd_blocks)
            mod = xt.ClipLoss(local_loss=True, gather_with_grad=True, rank=rank, world_size=world, pipeline=pipeline,
                              _kernels=k)
            loss = mod(I, T, s)
            loss.backward()
            errs.append((rel_err(loss.item(), losses[rank].item()), rel_err(I.grad.numpy(), dI[rank].numpy()),
                         rel_err(T.grad.numpy(), dT[rank].numpy()),
                         [c for c in k.calls if c[0] == "clip_fwd"]))
        q.put((rank, errs, None))
    except Exception:  # pragma: no cover
        import traceback
        q.put((rank, None, traceback.format_exc()))
    dist.barrier()
    dist.destroy_process_group()


def test_cliploss_world4_pipelined_groups_of_two():
    """world_size 4: the chunk pipeline exchanges groups of two ranks (2b columns per forward call); checked against
    the oracle's single-process emulation of the reference, and against the non-pipelined schedule."""
    world, port = 4, 29723
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker4, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for _ in range(world):
        rank, errs, err = q.get(timeout=240)
        assert err is None, err
        for e_loss, e_di, e_dt, fwd in errs:
            assert e_loss < 1e-5 and e_di < 1e-5 and e_dt < 1e-5, (rank, e_loss, e_di, e_dt)
        fwd_pipe = errs[0][3]
        assert [c[2][0] for c in fwd_pipe] == [10, 10] and fwd_pipe[0][3] == (rank % 2) * 5 and fwd_pipe[1][3] == -1
        assert [c[2][0] for c in errs[1][3]] == [20]
        assert errs[2][3] == fwd_pipe          # the two reduction strategies launch the same forward blocks
    for p in procs:
        p.join(timeout=60)


class FakeSymmExchange:
    """Stand-in for xtag_clip_b200.symm.SymmExchange over gloo (TEST INFRASTRUCTURE): same methods and bookkeeping,
    the peer-memory transfers replaced by collectives, so the host orchestration of the default multi-GPU path -- the
    streamed forward and the symmetric reduce-scatter backward of xtag_clip_b200/loss.py -- runs in the CPU tests."""

    def __init__(self, world, rank, b, D, dist):
        self.W, self.r, self.b, self.D, self.B, self.dist = world, rank, b, D, world * b, dist
        self.step = self.slot = 0
        self.flags = torch.zeros(world, dtype=torch.int32)
        self.epoch = torch.zeros(1, dtype=torch.int32)
        self.col = torch.zeros(self.B, dtype=torch.float32)
        self.dT = torch.zeros(self.B, D, dtype=torch.bfloat16)
        self._pushed = False
        self.bwd_pending = False
        self.log = []

    def begin_step(self):
        self.step += 1
        self.slot = self.step & 1

    def gather_streamed(self, x, out_all, pull_streams=1):
        self.epoch.add_(1)
        parts = [torch.empty_like(x) for _ in range(self.W)]
        self.dist.all_gather(parts, x.contiguous())
        for p in range(self.W):
            out_all[p * self.b:(p + 1) * self.b].copy_(parts[p])
        self.flags.fill_(int(self.epoch))
        self.log.append("gather_streamed")
        return [(self.r + j) % self.W for j in range(self.W)], [False] + [True] * (self.W - 1)

    def gather_pushed(self, x, streams=2):
        self.epoch.add_(1)
        parts = [torch.empty_like(x) for _ in range(self.W)]
        self.dist.all_gather(parts, x.contiguous())
        gbuf = torch.cat(parts, dim=0)
        self.flags.fill_(int(self.epoch))
        self._pushed = True
        self.log.append("gather_pushed")
        return gbuf, [(self.r - j) % self.W for j in range(self.W)], [False] + [True] * (self.W - 1), self.flags

    def push_step_done(self):
        self.bwd_pending = False
        self.log.append("push_step_done")

    def end_gather(self, streamed=False):
        self.log.append(("end_gather", streamed))

    def col_buffer(self):
        return self.col

    def combine_cols(self, K):
        parts = [torch.empty_like(self.col) for _ in range(self.W)]
        self.dist.all_gather(parts, self.col)
        return K.lse_combine_ptrs(parts, self.W, self.B)

    def dT_buffer(self):
        return self.dT

    def reduce_scatter_begin(self):
        self.log.append("reduce_scatter_begin")

    def reduce_scatter_end(self, K):
        parts = [torch.empty_like(self.dT) for _ in range(self.W)]
        self.dist.all_gather(parts, self.dT)
        lo, hi = self.r * self.b, (self.r + 1) * self.b
        return K.sum_ptrs_bf16([p[lo:hi] for p in parts], self.W, (self.b, self.D))


def _worker_symm(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    import torch.distributed as dist
    import oracle
    import xtag_clip_b200 as xt
    from xtag_clip_b200 import loss as xt_loss
    from kernel_model import ModelKernels
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    try:
        b, D, scale = 256, 16, 20.0                      # b is a multiple of 256: the streamed forward is eligible
        gen = torch.Generator().manual_seed(5)
        nrm = torch.nn.functional.normalize
        I_all = nrm(torch.randn(world * b, D, generator=gen), dim=-1).bfloat16()
        T_all = nrm(0.3 * I_all.float() + 0.7 * torch.randn(world * b, D, generator=gen), dim=-1).bfloat16()
        Il = [I_all[r * b:(r + 1) * b].double() for r in range(world)]
        Tl = [T_all[r * b:(r + 1) * b].double() for r in range(world)]
        losses, dI, dT, ds = oracle.clip_loss_world(Il, Tl, scale, True, True)
        fake = FakeSymmExchange(world, rank, b, D, dist)
        xt_loss._Comm.symm_exchange = lambda self, x: fake if x.dtype == torch.bfloat16 else None
        res = []
        for streamed in (True, False, "push"):
            k = ModelKernels()
            I = I_all[rank * b:(rank + 1) * b].clone().requires_grad_(True)
            T = T_all[rank * b:(rank + 1) * b].clone().requires_grad_(True)
            s = torch.tensor(scale, requires_grad=True)
            mod = xt.ClipLoss(local_loss=True, gather_with_grad=True, rank=rank, world_size=world, _kernels=k,
                              stream_forward=bool(streamed), exchange="push" if streamed == "push" else "pull")
            if not streamed:        # the per-block path pulls through gather_pipelined, which the fake does not model
                xt_loss._Comm.symm_exchange = lambda self, x: None
            else:
                fake.log.clear()
                xt_loss._Comm.symm_exchange = lambda self, x: fake if x.dtype == torch.bfloat16 else None
            loss = mod(I, T, s)
            loss.backward()
            res.append(dict(e_loss=rel_err(loss.item(), losses[rank].item()), e_di=rel_err(I.grad.float().numpy(), dI[rank].numpy()),
                            e_dt=rel_err(T.grad.float().numpy(), dT[rank].numpy()), calls=list(k.calls), log=list(fake.log)))
        # push-exchange guard: the gather buffer of a pushed forward is peer-writable and saved for its backward, so a
        # forward issued while that backward is outstanding, and a no_grad forward, must take the pull exchange
        fake.log.clear()
        fake.bwd_pending = False
        xt_loss._Comm.symm_exchange = lambda self, x: fake if x.dtype == torch.bfloat16 else None
        mod = xt.ClipLoss(local_loss=True, gather_with_grad=True, rank=rank, world_size=world, _kernels=ModelKernels(),
                          exchange="push")
        I = I_all[rank * b:(rank + 1) * b].clone().requires_grad_(True)
        T = T_all[rank * b:(rank + 1) * b].clone().requires_grad_(True)
        s = torch.tensor(scale, requires_grad=True)
        l1 = mod(I, T, s)
        l2 = mod(I, T, s)
        with torch.no_grad():
            mod(I, T, s)
        l1.backward()
        l3 = mod(I, T, s)
        res.append(dict(log=list(fake.log), same=float((l1 - l2).abs() + (l1 - l3).abs()),
                        e_di=rel_err(I.grad.float().numpy(), dI[rank].numpy())))
        q.put((rank, res, None))
    except Exception:  # pragma: no cover
        import traceback
        q.put((rank, None, traceback.format_exc()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,port", [(2, 29731), (4, 29732)])
def test_streamed_forward_and_symmetric_backward_host_logic(world, port):
    """The DEFAULT multi-GPU control flow (flag-gated streamed forward + symmetric-memory reduce-scatter backward) with
    a gloo-backed stand-in for the peer-memory exchange: per-rank loss and gradients against the oracle's emulation of
    the reference, the schedule handed to the kernel (ring order from the own block), and the backward's call pattern
    (dS + dB GEMM into the exchange buffer, reduce-scatter started, dA GEMM reusing the staged dS, reduce finished)."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_symm, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for _ in range(world):
        rank, res, err = q.get(timeout=300)
        assert err is None, err
        st, blk, push, guard = res
        names = [x if isinstance(x, str) else x[0] for x in guard["log"]]
        assert [n for n in names if n.startswith("gather")] == ["gather_pushed", "gather_streamed", "gather_streamed",
                                                                "gather_pushed"], names
        assert names.count("push_step_done") == 1 and guard["same"] < 1e-6 and guard["e_di"] < 2e-2
        for r_ in (st, blk, push):
            assert r_["e_loss"] < 1e-3 and r_["e_di"] < 2e-2 and r_["e_dt"] < 2e-2, (rank, r_["e_loss"], r_["e_di"], r_["e_dt"])
        fwd = [c for c in st["calls"] if c[0] == "clip_fwd_stream"]
        assert len(fwd) == 1 and fwd[0][1] == (256, 16) and fwd[0][2] == (world * 256, 16) and fwd[0][3] == rank * 256
        assert fwd[0][4] == tuple((rank + j) % world for j in range(world))
        bwd = [c for c in st["calls"] if c[0] == "clip_bwd"]
        assert len(bwd) == 2 and all(c[1] == (256, 16) and c[2] == (world * 256, 16) and c[3] == rank * 256 for c in bwd)
        assert st["log"] == ["gather_streamed", ("end_gather", True), "reduce_scatter_begin"]
        # push exchange (experimental): blocks arrive r-1, r-2, ...; the backward ends with the slot-release barrier
        fwd = [c for c in push["calls"] if c[0] == "clip_fwd_stream"]
        assert len(fwd) == 1 and fwd[0][4] == tuple((rank - j) % world for j in range(world))
        assert push["log"] == ["gather_pushed", ("end_gather", True), "reduce_scatter_begin", "push_step_done"]
    for p in procs:
        p.join(timeout=60)


def _worker_siglip(rank, world, port, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    import xtag_clip_b200 as xt
    from kernel_model import ModelKernels
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "siglip.npz"))
        pre = f"w{world}_"
        b = int(g[pre + "b"])
        I = torch.from_numpy(g[pre + "I"])[rank * b:(rank + 1) * b].clone().requires_grad_(True)
        T = torch.from_numpy(g[pre + "T"])[rank * b:(rank + 1) * b].clone().requires_grad_(True)
        s = torch.tensor(float(g[pre + "scale"]), dtype=torch.float64, requires_grad=True)
        bi = torch.tensor(float(g[pre + "bias"]), dtype=torch.float64, requires_grad=True)
        k = ModelKernels()
        loss = xt.SigLipLoss(rank=rank, world_size=world, dist_impl="gather", compute_dtype=torch.bfloat16,
                             _kernels=k)(I, T, s, bi)
        loss.backward()
        key = f"{pre}r{rank}_"
        # the drop-in rounds the features to bf16 (its compute dtype); the fixture is the reference's fp64 run
        res = dict(e_loss=rel_err(loss.item(), float(g[key + "loss"])), e_di=rel_err(I.grad.numpy(), g[key + "dI"]),
                   e_dt=rel_err(T.grad.numpy(), g[key + "dT"]),
                   e_ds=abs(float(s.grad) - float(g[key + "dscale"])) / max(abs(float(g[key + "dscale"])), 1e-12),
                   e_db=abs(float(bi.grad) - float(g[key + "dbias"])) / max(abs(float(g[key + "dbias"])), 1e-12),
                   calls=list(k.calls))
        q.put((rank, res, None))
    except Exception:  # pragma: no cover
        import traceback
        q.put((rank, None, traceback.format_exc()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,port", [(2, 29741), (3, 29742)])
def test_siglip_sharded_host_logic_vs_reference_fixture(world, port):
    """Sharded SigLipLoss under gloo with the CPU contract model: every rank's loss and gradients against what the
    REFERENCE's SigLipLoss(dist_impl='gather') produced under gloo (tests/golden/siglip.npz) -- own block with
    positives at label offset rank*b, text gradients reduce-scattered, per-rank scale / bias gradients."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_siglip, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for _ in range(world):
        rank, res, err = q.get(timeout=300)
        assert err is None, err
        assert res["e_loss"] < 2e-2 and res["e_di"] < 3e-2 and res["e_dt"] < 3e-2 and res["e_ds"] < 3e-2 and res["e_db"] < 3e-2, res
        fwd = [c for c in res["calls"] if c[0] == "siglip_fwd"]
        assert len(fwd) == 1 and fwd[0][3] == rank * fwd[0][1][0] and fwd[0][2][0] == world * fwd[0][1][0]
    for p in procs:
        p.join(timeout=60)
