"""Particle-sharded filter update: one process per GPU, `torch.distributed` for the plumbing.

Particles are independent in the motion and sensor stages, so a set of N particles is split
into contiguous blocks over the ranks with the map and the scan replicated.
One update needs exactly three exchanges:

  1. all-reduce(max) of the largest raw (log-)weight       -> common fixed-point scale
  2. all-gather of the per-rank fixed-point weight totals   -> every rank derives the global
     total T, its exclusive prefix and, in closed form, the contiguous range of output slots
     whose systematic thresholds fall into its CDF interval (amclj_plan_systematic)
  3. the particles themselves, in one of two ways:
     "pull" (default on GPUs): every rank fills its own output slots by reading the owning
         rank's CDF and particles through NVLink peer pointers inside one kernel
         (amclj_shard_pull_async; CUDA IPC handles are exchanged once) — no staging, no
         all-to-all, no host round trip;
     "a2a": each rank generates the offspring of its CDF interval and an NCCL all-to-all moves
         the rows to the ranks that own the slots (slot m lives on rank m / (N/G)).

     "mailbox" (default on GPUs): exchanges 1 and 2 are done by the library's own kernels as well —
         every rank stores its {max, total} into every peer's mailbox over NVLink and polls its
         own (12 bytes per rank and update) — so an update is ONE library call per rank
         (amclj_shard_update_async): no NCCL, no Python between the kernels, nothing but the
         one-time exchange of CUDA IPC handles goes through torch.distributed.

The resulting global index list is identical to the single-GPU one (integer CDF).  The
reference has no counterpart (single-threaded Clojure, pf.clj:222-240); this is the
data-parallel form of pf.clj:229-233.

The device work sits behind a small backend protocol so the choreography can be exercised
on CPU (gloo, world_size 2) with a stand-in backend in the tests.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from . import _lib


class CudaShard:
    """Backend over one ``amclj_ctx`` (libamclj_b200.so)."""

    def __init__(self, ctx: _lib.Context, device: torch.device):
        self.ctx, self.device = ctx, device

    def motion_sensor(self, curr, prev, seed):
        self.ctx.shard_motion_sensor_async(curr, prev, seed)

    def raw_max(self) -> torch.Tensor:   # float32[1], updated in place by the all-reduce
        return self.ctx.device_tensor(_lib.BUF_RAW_MAX, (1,), "<f4")

    def cdf(self) -> torch.Tensor:       # int64[1] view of the u64 local total
        self.ctx.shard_cdf_async()
        return self.ctx.device_tensor(_lib.BUF_TOTAL, (1,), "<i8")

    def reserve(self, rows):
        self.ctx.reserve_stage(rows)

    def generate(self, total, offset, r, m_lo, count) -> torch.Tensor:
        if count == 0:
            return self.empty(0, 4, dtype=torch.float32)
        self.ctx.shard_generate_async(total, offset, r, m_lo, count)
        return self.ctx.device_tensor(_lib.BUF_STAGE_OUT, (count, 4), "<f4")

    def inbox(self, rows) -> torch.Tensor:
        return self.ctx.device_tensor(_lib.BUF_STAGE_IN, (rows, 4), "<f4")

    def adopt(self, rows):
        self.ctx.shard_adopt_async(rows)

    def set_shard(self, rank, world, offset, n_global):
        self.ctx.set_shard(rank, world, offset, n_global)

    def empty(self, *shape, dtype):
        return torch.empty(*shape, dtype=dtype, device=self.device)

    # -- peer-memory ("pull") variant ---------------------------------------------------------
    def connect_peers(self, rank, world, n_global, group=None):
        """Exchange CUDA IPC handles once; afterwards every rank can read every shard."""
        mine = self.ctx.shard_export()
        handles = [None] * world
        dist.all_gather_object(handles, mine, group=group)
        for g in range(world):
            lo, cnt = slots_of(g, world, n_global)
            if g == rank:
                self.ctx.shard_attach_local(g, self.ctx)
            else:
                self.ctx.shard_import(g, handles[g], cnt, lo)

    def totals_all(self, world) -> torch.Tensor:
        return self.ctx.device_tensor(_lib.BUF_TOTALS_ALL, (16,), "<i8")[:world]

    def pull(self, u0):
        self.ctx.shard_pull_async(u0)

    def update_mailbox(self, curr, prev, seed, u0, scan=None):
        if scan is not None:  # stage the scan and update in one library call
            self.ctx.shard_update_scan_async(curr, prev, seed, u0, scan.ranges, scan.angle_min, scan.angle_inc, scan.range_max)
        else:
            self.ctx.shard_update_async(curr, prev, seed, u0)


def slots_of(rank, world, n_global):
    """Output slots owned by ``rank``: n_global / world each, the remainder to the low ranks."""
    q, rem = divmod(n_global, world)
    start = rank * q + min(rank, rem)
    return start, q + (1 if rank < rem else 0)


class ShardedFilter:
    def __init__(self, ctx, rank, world, n_global, device, stream=None, backend=None, group=None,
                 mode=None):
        self.rank, self.world, self.n_global, self.group = rank, world, int(n_global), group
        self.backend = backend if backend is not None else CudaShard(ctx, device)
        self.ctx = ctx
        # The NCCL collectives run on torch's current stream and read / write library buffers in place,
        # so the library's kernels must be ordered on that very stream: bind the ctx to it here and run
        # every update under it (the race-freedom argument of pull mode needs collectives and kernels on
        # ONE stream as well).
        self.stream = None
        if backend is None:
            self.stream = stream if stream is not None else torch.cuda.current_stream(device)
            ctx.set_stream(self.stream.cuda_stream)
        self.slot0, self.n_slots = slots_of(rank, world, self.n_global)
        self.backend.set_shard(rank, world, self.slot0, self.n_global)
        self.mode = mode or ("mailbox" if backend is None and world <= 16 else "a2a")
        self.connected = False
        if self.mode == "a2a":
            self.backend.reserve(max(self.n_slots, 1))
        self.last_plan = None

    def connect(self):
        """Pull mode: map the peers' buffers (call once, after the particles are in place)."""
        if self.mode in ("pull", "mailbox") and not self.connected:
            self.backend.connect_peers(self.rank, self.world, self.n_global, self.group)
            self.connected = True

    def stage_scan(self, scan):
        self.ctx.stage_scan(scan.ranges, scan.angle_min, scan.angle_inc, scan.range_max)

    def reset(self, n_local):
        self.backend.set_shard(self.rank, self.world, self.slot0, self.n_global)

    def update(self, curr, prev, seed, u0, scan=None):
        if self.stream is not None:
            with torch.cuda.stream(self.stream):
                return self._update(curr, prev, seed, u0, scan)
        return self._update(curr, prev, seed, u0, scan)

    def _update(self, curr, prev, seed, u0, scan=None):
        b = self.backend
        if self.mode == "mailbox":
            self.connect()
            b.update_mailbox(curr, prev, seed, u0, scan)  # one library call: the kernels do the exchanges
            return None
        if scan is not None:
            self.stage_scan(scan)
        b.motion_sensor(curr, prev, seed)
        mx = b.raw_max()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=self.group)
        tot = b.cdf()
        if self.mode == "pull":
            self.connect()
            dist.all_gather_into_tensor(b.totals_all(self.world), tot, group=self.group)
            b.pull(u0)  # reads the peers' CDFs and particles over NVLink; no host round trip
            return None
        totals = b.empty(self.world, dtype=torch.int64)
        dist.all_gather_into_tensor(totals, tot, group=self.group)
        h = totals.cpu().numpy().view(np.uint64)  # the one host round trip of the update
        r, off, m_lo, m_cnt, send = _lib.plan_systematic(h, self.n_global, u0)
        self.last_plan = (r, off, m_lo, m_cnt, send)
        T = int(h.sum(dtype=np.uint64))
        if T == 0:
            raise _lib.AmcljError(_lib.ERR_STATE, "sharded update: all weights are zero")
        cnt = int(m_cnt[self.rank])
        b.reserve(max(cnt, self.n_slots))
        out = b.generate(T, int(off[self.rank]), r, int(m_lo[self.rank]), cnt)
        inbox = b.inbox(self.n_slots)
        dist.all_to_all_single(inbox, out, output_split_sizes=[int(v) for v in send[:, self.rank]],
                               input_split_sizes=[int(v) for v in send[self.rank, :]], group=self.group)
        b.adopt(self.n_slots)
        return T
