"""NVLink exchange for the sharded contrastive head over peer-mapped ("symmetric") memory.

The sharded loss has one exchange step each way: forward needs every rank's text features
and the [B] partial column log-sum-exps, backward reduces the [B, D] text-gradient partials onto their owners.  The
reference does this with blocking NCCL collectives (loss.py:51-52 and the reduce-scatter inside
torch.distributed.nn.all_gather's backward).  NCCL kernels need SMs, but K1/K2 are persistent kernels that occupy
all 148 of them, so a concurrent NCCL kernel simply waits (measured: 0.3-0.9 ms stalls at 8 GPUs).  Here every bulk
transfer is a *pull by the copy engines* from the peers' symmetric buffers -- no SMs involved -- so it overlaps with
the tcgen05 kernels:

  forward   features -> own slot; barrier; pull peer blocks chunk group by chunk group on two copy streams while
            K1 already runs on the blocks that have landed
  forward   K1 writes its column-LSE partials straight into a symmetric [B] buffer; barrier; one small kernel
            reads all W peer buffers over NVLink and LSE-combines them
  backward  the dB GEMM writes its [B, D] partial into a symmetric buffer; barrier; each rank pulls its own block
            from the W-1 peers (copy engines, overlapped with the dA GEMM); one kernel sums the W blocks

Buffers are double-buffered by step parity; together with the one barrier per use and step this orders every
overwrite after all peers' reads of the previous use.
torch supplies the allocation / rendezvous / barrier (torch.distributed._symmetric_memory); it is plumbing.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


def available(device: torch.device) -> bool:
    if device.type != "cuda":
        return False
    try:
        import torch.distributed._symmetric_memory as symm_mem  # noqa: F401
        return True
    except Exception:
        return False


class SymmExchange:
    def __init__(self, world: int, rank: int, group, device: torch.device, b: int, D: int,
                 feat_dtype: torch.dtype, grad_dtype: torch.dtype):
        import torch.distributed._symmetric_memory as symm_mem
        self.W, self.r, self.b, self.D = world, rank, b, D
        self.B = world * b
        self.feat_dtype, self.grad_dtype = feat_dtype, grad_dtype
        grp = group if group is not None else dist.group.WORLD
        self._grp = grp
        self.feat = symm_mem.empty((2, b, D), dtype=feat_dtype, device=device)
        self.h_feat = symm_mem.rendezvous(self.feat, grp)
        self.col = symm_mem.empty((2, self.B), dtype=torch.float32, device=device)
        self.h_col = symm_mem.rendezvous(self.col, grp)
        self.dT = symm_mem.empty((2, self.B, D), dtype=grad_dtype, device=device)
        self.h_dT = symm_mem.rendezvous(self.dT, grp)
        self.peer_feat = [self.h_feat.get_buffer(p, (2, b, D), feat_dtype) for p in range(world)]
        self.peer_dT = [self.h_dT.get_buffer(p, (2, self.B, D), grad_dtype) for p in range(world)]
        peer_col = [self.h_col.get_buffer(p, (2, self.B), torch.float32) for p in range(world)]
        # device arrays of peer pointers for the combine kernel, one per slot
        self.col_ptrs = [torch.tensor([peer_col[p][s].data_ptr() for p in range(world)], dtype=torch.int64,
                                      device=device) for s in range(2)]
        self._keep = peer_col
        self.pull_tmp = torch.empty((world, b, D), dtype=grad_dtype, device=device)
        # operand pointers of the final sum, per slot: the W-1 pulled blocks and the rank's own block of its partial
        self.sum_ptrs = [torch.tensor([self.pull_tmp[p].data_ptr() if p != rank
                                       else self.dT[s][rank * b:(rank + 1) * b].data_ptr() for p in range(world)],
                                      dtype=torch.int64, device=device) for s in range(2)]
        self.s1 = torch.cuda.Stream(device=device)
        self.s2 = torch.cuda.Stream(device=device)
        self.s3 = torch.cuda.Stream(device=device)      # push exchange: slot-release barrier at the end of a backward
        self.bwd_pending = False          # push exchange: a forward's gather buffer is still needed by its backward
        self.step = 0
        self.slot = 0
        # streamed forward: per-block ready flags, written by the copy stream right behind each pulled block with the
        # step's epoch value (a device counter, so a captured step replays with fresh values)
        self.flags = torch.zeros(world, dtype=torch.int32, device=device)
        # the epoch of the NEXT flag-gated forward; bumped on the device at the end of each one (combine_cols_loss)
        self.epoch = torch.ones(1, dtype=torch.int32, device=device)
        self._streamed_two = False        # the last streamed gather forked the second copy stream
        self._pushed = False              # the push exchange is in use (its backward ends with a slot-release barrier)
        self.gbuf = None                  # push exchange: peer-writable gather buffers, allocated on first use
        self._loss_scratch = None         # ticket + partials of the fused combine + loss kernel
        self.gg = None                    # captured steps: ONE symmetric gather buffer [B, D] (make_graph_gather)
        self.peer_gg = None

    # ---- schedule ---------------------------------------------------------------------------------------------
    def _blocks(self) -> List[Tuple[int, int, List[int]]]:
        """Column blocks in consumption order: (first rank, last rank + 1, peers to pull).  The rank's own chunk
        first (no wait), then its group partner, then the other groups of G ranks walking away from the rank."""
        W, r = self.W, self.r
        G = 2 if (W >= 4 and W % 2 == 0) else 1
        NG, g0 = W // G, r // G
        out = [(r, r + 1, [])]
        if G == 2:
            out.append((r ^ 1, (r ^ 1) + 1, [r ^ 1]))
        for j in range(1, NG):
            g = (g0 - j) % NG
            out.append((g * G, g * G + G, list(range(g * G, g * G + G))))
        return out

    # ---- captured steps: the gather buffer itself lives in symmetric memory -----------------------------------
    def make_graph_gather(self) -> torch.Tensor:
        """A CUDA-graph replay uses one fixed gather buffer anyway, so for captured steps it is allocated in symmetric
        memory: the rank's own block of it doubles as the text-feature INPUT SLOT of the step (the producer writes its
        features there) and as the source the peers pull from -- no own-block copies at all (two 8 MiB copies per
        step at config 5 / 8 GPUs otherwise).  Safe with a single buffer: a rank overwrites its own block for step
        t+1 only after its step-t backward, and every peer finished pulling that block before this rank could pass
        step t's column-LSE barrier.  Collective: every rank calls it at the same point."""
        if self.gg is None:
            import torch.distributed._symmetric_memory as symm_mem
            self.gg = symm_mem.empty((self.B, self.D), dtype=self.feat_dtype, device=self.feat.device)
            self.h_gg = symm_mem.rendezvous(self.gg, self._grp)
            self.peer_gg = [self.h_gg.get_buffer(p, (self.B, self.D), self.feat_dtype) for p in range(self.W)]
        return self.gg

    def own_block(self, buf: torch.Tensor) -> torch.Tensor:
        return buf[self.r * self.b:(self.r + 1) * self.b]

    # ---- forward: feature gather ------------------------------------------------------------------------------
    def begin_step(self):
        self.step += 1
        self.slot = self.step & 1

    def gather_pipelined(self, x: torch.Tensor, out_all: torch.Tensor):
        """-> [(row_lo, row_hi, event | None)]: blocks of out_all in the order they become valid."""
        b, s = self.b, self.slot
        cur = torch.cuda.current_stream()
        out_all[self.r * b:(self.r + 1) * b].copy_(x)
        self.s1.wait_stream(cur)
        plan = []
        with torch.cuda.stream(self.s1):
            self.feat[s].copy_(x)
            self.h_feat.barrier(channel=0)                  # every rank's slot s is written
        ready = torch.cuda.Event()
        ready.record(self.s1)
        self.s2.wait_event(ready)
        k = 0
        for lo_r, hi_r, peers in self._blocks():
            evs = []
            for p in peers:
                st = self.s1 if (k & 1) == 0 else self.s2
                k += 1
                with torch.cuda.stream(st):
                    out_all[p * b:(p + 1) * b].copy_(self.peer_feat[p][s], non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(st)
                evs.append(ev)
            plan.append((lo_r * b, hi_r * b, evs))
        return plan

    def gather_streamed(self, x: torch.Tensor, out_all: torch.Tensor, pull_streams: int = 1):
        """Pull the peers' blocks in ring order r+1, r+2, ..., each followed on its stream by a 4-byte copy of the
        step's epoch into its ready flag.  Returns (order, wait) for `clip_fwd_stream`: the persistent K1 launch on the
        compute stream starts with the rank's own block and picks up each peer block the moment its flag flips -- no
        event waits, no launch per block.  `end_gather` joins the copy stream(s) afterwards.
        pull_streams = 1: all pulls on ONE copy stream -- a fixed arrival order, but the flag copy
        and the scheduling gaps around it sit between consecutive block copies (measured: 28 us per 8 MiB block of which
        17 us is the copy).  pull_streams = 2 (default of ClipLoss): blocks alternate between two copy streams so that one stream's flag
        copy and gaps hide behind the other stream's block copy (K1's correctness does not depend on the arrival
        order: it waits per block)."""
        b, s, W, r = self.b, self.slot, self.W, self.r
        cur = torch.cuda.current_stream()
        unified = self.gg is not None and out_all.data_ptr() == self.gg.data_ptr()
        if x.data_ptr() != out_all[r * b:(r + 1) * b].data_ptr():
            out_all[r * b:(r + 1) * b].copy_(x)             # (skipped when the features were produced in the slot)
        self.s1.wait_stream(cur)
        order = [(r + j) % W for j in range(W)]
        src = (lambda p: self.peer_gg[p][p * b:(p + 1) * b]) if unified else (lambda p: self.peer_feat[p][s])
        with torch.cuda.stream(self.s1):
            if not unified:
                self.feat[s].copy_(x)
            self.h_feat.barrier(channel=0)                  # every rank's block of this step is in place
        streams = [self.s1]
        if pull_streams >= 2 and W > 2:
            ready = torch.cuda.Event()
            ready.record(self.s1)
            self.s2.wait_event(ready)
            streams.append(self.s2)
        for k, p in enumerate(order[1:]):
            with torch.cuda.stream(streams[k % len(streams)]):
                out_all[p * b:(p + 1) * b].copy_(src(p), non_blocking=True)
                self.flags[p:p + 1].copy_(self.epoch, non_blocking=True)
        self._streamed_two = len(streams) == 2
        return order, [False] + [True] * (W - 1)

    # ---- forward: PUSH variant of the streamed gather (experimental, not yet validated on hardware) -------------
    def _push_buffers(self):
        """Peer-writable gather buffers [2][B, D] and per-slot ready flags [2][W] (lazily: only the push exchange
        needs them)."""
        if self.gbuf is None:
            import torch.distributed._symmetric_memory as symm_mem
            grp = self._grp
            dev = self.feat.device
            self.gbuf = symm_mem.empty((2, self.B, self.D), dtype=self.feat_dtype, device=dev)
            self.h_gbuf = symm_mem.rendezvous(self.gbuf, grp)
            self.pflags = symm_mem.empty((2, self.W), dtype=torch.int32, device=dev)
            self.pflags.zero_()
            self.h_pflags = symm_mem.rendezvous(self.pflags, grp)
            self.peer_gbuf = [self.h_gbuf.get_buffer(p, (2, self.B, self.D), self.feat_dtype) for p in range(self.W)]
            self.peer_pflags = [self.h_pflags.get_buffer(p, (2, self.W), torch.int32) for p in range(self.W)]
            self.h_pflags.barrier(channel=0)                # every rank's flags are zeroed before anyone pushes
        return self.gbuf, self.pflags

    def gather_pushed(self, x: torch.Tensor, streams: int = 2):
        """Every rank PUSHES its block into all peers' gather buffers (copy-engine writes over NVLink, ring order
        r+1, r+2, ...) and writes the step's epoch into the peer's ready flag behind it.  The consumer is the
        flag-gated persistent K1 launch, so no barrier is needed before the exchange: a rank that is late only delays
        its own block at its peers (the pull exchange pays the start-of-step barrier skew, measured 68 us at 8 GPUs).
        streams = 2: the blocks alternate between two copy streams, so one stream's flag copy and scheduling gaps
        (11 us per block, measured) hide behind the other stream's 17 us block copy -- the blocks then arrive faster
        than K1 consumes them (one [b, B/W] column block of config 5 is ~19 us of tensor time).
        Slot safety: the previous backward ended with a release barrier on s3 (`push_step_done`): no peer still reads
        the slot this step's pushes overwrite.  Returns (gather buffer [B, D] of this step, order, wait, flags [W])."""
        gbuf, pflags = self._push_buffers()
        b, s, W, r = self.b, self.slot, self.W, self.r
        cur = torch.cuda.current_stream()
        gbuf[s][r * b:(r + 1) * b].copy_(x)
        self.s1.wait_stream(cur)
        if not torch.cuda.is_current_stream_capturing():    # (a captured backward joins s3 itself before it ends)
            self.s1.wait_stream(self.s3)                    # release barrier of the previous backward
        sts = [self.s1]
        if streams >= 2 and W > 2:
            ready = torch.cuda.Event()
            ready.record(self.s1)
            self.s2.wait_event(ready)
            sts.append(self.s2)
        for j in range(1, W):
            p = (r + j) % W
            with torch.cuda.stream(sts[(j - 1) % len(sts)]):
                self.peer_gbuf[p][s][r * b:(r + 1) * b].copy_(x, non_blocking=True)
                self.peer_pflags[p][s][r:r + 1].copy_(self.epoch, non_blocking=True)
        self._streamed_two = len(sts) == 2
        self._pushed = True
        # the block of rank q is q's ((r - q) % W)-th push: r-1 lands first, then r-2, ...
        order = [(r - j) % W for j in range(W)]
        return gbuf[s], order, [False] + [True] * (W - 1), pflags[s]

    def push_step_done(self):
        """End of a backward in push mode: no rank may start pushing the next step's blocks into a peer's gather buffer
        while that peer's gradient GEMMs still read it.  Eager steps alternate between two slots and would be safe
        without this, but a captured step replays ONE slot.  The barrier runs on its own stream behind everything the
        backward has enqueued so far; only the next step's pushes wait for it (`gather_pushed`), not the compute stream
        -- except under stream capture, where every forked stream must be joined before the capture ends."""
        cur = torch.cuda.current_stream()
        self.s3.wait_stream(cur)
        with torch.cuda.stream(self.s3):
            self.h_gbuf.barrier(channel=1)
        if torch.cuda.is_current_stream_capturing():
            cur.wait_stream(self.s3)
        self.bwd_pending = False

    def end_gather(self, streamed: bool = False):
        """the pull streams must be drained before the next barrier on s1 (orders the next overwrite of a slot).
        The streamed gather never forks s2 (waiting on it would pull an un-captured stream into a graph capture)."""
        if not streamed or self._streamed_two:
            self.s1.wait_stream(self.s2)
        torch.cuda.current_stream().wait_stream(self.s1)

    # ---- forward: column-LSE exchange -------------------------------------------------------------------------
    def col_buffer(self) -> torch.Tensor:
        return self.col[self.slot]

    def combine_cols(self, K) -> torch.Tensor:
        self.h_col.barrier(channel=1)                       # on the compute stream: all K1 launches precede it
        return K.lse_combine_ptrs(self.col_ptrs[self.slot], self.W, self.B)

    def combine_cols_loss(self, K, row_lse: torch.Tensor, diag: torch.Tensor, label_offset: int, streamed: bool):
        """barrier + ONE kernel: combined column LSEs [B], this rank's loss, and (flag-gated forward) the bump of the
        exchange epoch for the next step.  -> (col_lse, loss)"""
        self.h_col.barrier(channel=1)
        if hasattr(K, "lse_combine_ptrs_loss"):
            if self._loss_scratch is None:
                self._loss_scratch = torch.zeros(K.combine_loss_scratch_bytes(), dtype=torch.uint8,
                                                 device=self.feat.device)
            return K.lse_combine_ptrs_loss(self.col_ptrs[self.slot], self.W, self.B, row_lse, diag, label_offset,
                                           self.epoch if streamed else None, self._loss_scratch)
        col = K.lse_combine_ptrs(self.col_ptrs[self.slot], self.W, self.B)
        if streamed:
            self.epoch.add_(1)
        return col, K.clip_loss(row_lse, diag, col, label_offset)

    # ---- backward: reduce-scatter of the text-gradient partials -----------------------------------------------
    def dT_buffer(self) -> torch.Tensor:
        return self.dT[self.slot]

    def reduce_scatter_begin(self):
        """call after the dB GEMM was enqueued on the current stream; pulls run on the copy streams"""
        b, s, r = self.b, self.slot, self.r
        cur = torch.cuda.current_stream()
        self.s1.wait_stream(cur)
        with torch.cuda.stream(self.s1):
            self.h_dT.barrier(channel=2)                    # every rank's partial is complete
        ready = torch.cuda.Event()
        ready.record(self.s1)
        self.s2.wait_event(ready)
        k = 0
        for j in range(1, self.W):
            p = (r + j) % self.W
            st = self.s1 if (k & 1) == 0 else self.s2
            k += 1
            with torch.cuda.stream(st):
                self.pull_tmp[p].copy_(self.peer_dT[p][s][r * b:(r + 1) * b], non_blocking=True)

    def reduce_scatter_end(self, K) -> torch.Tensor:
        cur = torch.cuda.current_stream()
        self.s1.wait_stream(self.s2)
        cur.wait_stream(self.s1)
        return K.sum_ptrs_bf16(self.sum_ptrs[self.slot], self.W, (self.b, self.D))
