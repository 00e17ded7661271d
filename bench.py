#!/usr/bin/env python
"""bench.py — the hot path of one filter update (motion -> ray-cast + beam mixture ->
normalise -> systematic resample) on synthetic inputs of BASELINE.json's configurations.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

The SAME workload at every N (weak scaling): the C4 configuration's per-GPU share — 1.25M particles
uniform over lgfloor's free space x 360 beams per GPU, i.e. 10M particles at N = 8.  A step = the scan
staged (host beam table + upload) + one full update, timed on the device; the particle set is restored
(untimed) and L2 flushed (untimed) before every step so that every step sees the named configuration.
N > 1 (torchrun, one rank per GPU): one library call per rank and update (amclj_shard_update_async) —
the ranks exchange {max, total} through peer-memory mailboxes written and polled by the library's own
kernels and pull their resampled slots over NVLink; no NCCL call and no Python between the kernels of an
update.  Before timing, the N > 1 run checks its particles against the same update of ALL particles on
one GPU (`parity`).  The N = 1 line also carries C2, C3 and C5 as sub-records (`configs`).

Prints ONE JSON line (rank 0).  `value` = beam evaluations per second, particle set resident in HBM;
`e2e` = the same through the host-buffer C-ABI call (upload, update, download; pinned and pageable).
`--impl reference` times the CPU oracle — the stand-in for the Clojure reference, which has no JVM to run
on here — on a bounded particle sample of the SAME workload with every host core.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

METRIC = "beam_evals_per_s"
UNIT = "beam-evals/s"
PER_GPU = 1_250_000  # C4: 10M particles over 8 GPUs
SENSOR_PATHS = {-1: "unknown", 0: "walk kernel, particles in map-tile order (spread-out cloud)",
                1: "walk kernel, particles in place (compact cloud)",
                2: "per-update ray tables (dense cloud)",
                3: "persistent map-wide ray tables (one table per free cell, filled once; looked up from shared memory)"}
KERNELS = {3: "amclj::pt_lookup_kernel", 2: "amclj::rt_lookup_kernel"}


def host_threads():
    """Every core this process may run on — NOT what OMP_NUM_THREADS says (torchrun sets it to 1)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock + throttle reasons during the timed region (B200_PROFILING.md): NVML every few
    milliseconds (nvidia-smi as a fallback)."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    BITS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, index):
        self.index, self.sm, self.reasons, self.max_mhz, self._stop = index, [], set(), None, threading.Event()
        self._t = threading.Thread(target=self._run, daemon=True)
        self._nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nvml = pynvml
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and vis.split(",")[index].isdigit() else index
            self._h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self._nvml = None

    def _run(self):
        while not self._stop.is_set():
            try:
                if self._nvml is not None:
                    n = self._nvml
                    self.sm.append(int(n.nvmlDeviceGetClockInfo(self._h, n.NVML_CLOCK_SM)))
                    mask = int(n.nvmlDeviceGetCurrentClocksEventReasons(self._h)) if hasattr(
                        n, "nvmlDeviceGetCurrentClocksEventReasons") else int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self._h))
                    self.reasons |= {k for k, b in self.BITS.items() if mask & b}
                    self._stop.wait(0.004)
                    continue
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout
                r = [c.strip() for c in out.strip().split(",")]
                if r and r[0].isdigit():
                    self.sm.append(int(r[0]))
                    self.max_mhz = int(r[1]) if r[1].isdigit() else self.max_mhz
                    self.reasons |= {k for j, k in enumerate(["hw_slowdown", "hw_thermal_slowdown",
                                                              "sw_thermal_slowdown", "sw_power_cap"])
                                     if r[2 + j].lower().startswith("active")}
            except Exception:
                pass
            self._stop.wait(0.05)

    def __enter__(self):
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["clock sampling unavailable"]}
        sm = sorted(self.sm)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(sm), "source": "nvml" if self._nvml is not None else "nvidia-smi"}


# ---------------------------------------------------------------------------------------------
# workloads
# ---------------------------------------------------------------------------------------------
def headline_workload(world, rank=None):
    """C4 at its per-GPU share.  rank=None: the whole N x 1.25M set."""
    from amclj_b200 import workloads
    w = workloads.c3(PER_GPU * world)
    if rank is not None:
        lo = rank * PER_GPU
        w.x, w.y, w.theta = (a[lo:lo + PER_GPU].copy() for a in (w.x, w.y, w.theta))
    w.name = f"C4 per-GPU share: lgfloor, {PER_GPU} particles/GPU uniform over free space x 360 beams (10M at 8 GPUs)"
    return w


def config_of(w, world):
    """Identical for the GPU arm and the reference arm (sampling notes live in cpu_baseline.sample)."""
    return {"workload": w.name,
            "map": f"{w.grid.info.width}x{w.grid.info.height}@{float(w.grid.info.resolution):.2f}m",
            "particles_per_gpu": PER_GPU, "particles_total": PER_GPU * world, "beams": len(w.scan.ranges),
            "range_max_m": w.scan.range_max,
            "preset": "NS (north-star semantics: heading-correct Bresenham, all beams, log-weights, systematic)",
            "step": "scan staged (host beam table + upload) + motion + sensor + normalise + systematic resample, one library call",
            "l2": "flushed between timed steps (256 MiB memset, untimed); particle set restored before each step (untimed)",
            "barrier": "per step: dist.barrier + synchronize on both sides; N > 1: its release aligned on the device time line "
                       "by amclj_shard_rendezvous_async (a one-warp kernel over the peer-memory mailboxes) before the start event",
            "parallelism": f"particles sharded x{world}, map and scan replicated"}


# ---------------------------------------------------------------------------------------------
# CPU legs (the oracle: CPU port of the Clojure path, double precision, one cell per Bresenham step)
# ---------------------------------------------------------------------------------------------
def cpu_pass(w, sample, threads):
    """Motion + sensor + normalise + systematic resample over the first `sample` particles of `w` on
    `threads` OpenMP threads.  Returns seconds."""
    import oracle

    po = oracle.params("NS")
    sel = slice(0, sample)
    nrm = np.random.Generator(np.random.PCG64(3)).standard_normal((sample, 3)).astype(np.float32)
    s = w.scan
    t0 = time.perf_counter()
    mx, my, mt = oracle.motion_f64(po, w.curr, w.prev, nrm, w.x[sel], w.y[sel], w.theta[sel])
    raw, _ = oracle.sensor_f64(po, w.grid, s.ranges, s.angle_min, s.angle_inc, s.range_max, mx, my, mt,
                               threads=threads)
    lin, _ = oracle.linear_weights(raw.astype(np.float32), True)
    oracle.systematic_resample(lin, w.u0)
    return time.perf_counter() - t0


def cpu_baseline(w, budget_s=10.0, sample=100_000):
    """About `budget_s` seconds of CPU work on a bounded sample of the workload, every host core."""
    import oracle
    oracle.build()
    nth = host_threads()
    sample = min(sample, w.n)
    cpu_pass(w, min(sample, 2000), nth)  # warm the library and the page cache
    secs = []
    while sum(secs) < budget_s and len(secs) < 64:
        secs.append(cpu_pass(w, sample, nth))
    nb = len(w.scan.ranges)
    return {"value": sample * nb * len(secs) / sum(secs), "unit": UNIT, "cores": nth, "kind": "port",
            "sample": f"{len(secs)} passes over the first {sample} particles of the workload x {nb} beams "
                      f"({sum(secs):.1f} s of CPU work), oracle f64 (C port of the Clojure path) with OpenMP on {nth} "
                      f"threads; the Clojure reference itself cannot run here (no JVM)"}


def run_reference(args):
    """`--impl reference`: the reference's CPU implementation of the path on the host cores.  The
    Clojure/JVM original cannot run here (no JVM in the image), so this is the oracle port on every host
    core; a step is a bounded particle sample of the SAME workload as the GPU arm."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle

    oracle.build()
    world = max(1, args.gpus)
    w = headline_workload(1)  # the per-GPU share; every share is the same distribution
    nth = host_threads()
    assert nth > 1 or (os.cpu_count() or 1) == 1, "the reference arm must use every host core"
    sample = min(w.n, args.cpu_sample)
    nb = len(w.scan.ranges)
    for _ in range(args.warmup):
        cpu_pass(w, min(sample, 2000), nth)
    secs = [cpu_pass(w, sample, nth) for _ in range(args.steps)]
    value = sample * nb * len(secs) / sum(secs)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(secs) / len(secs),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_of(w, world),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": nth, "kind": "port",
                         "sample": f"{sample} of the workload's particles x {nb} beams per step, oracle f64 (C port of "
                                   f"the Clojure path), OpenMP {nth} threads; the Clojure reference cannot run: no JVM"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
def time_single(ctx, w, stream, flush, steps, warmup, torch, df_mode=0, stage_in_step=True):
    """K timed steps of workload `w` on one ctx (N = 1 path); returns per-step ms and the last stats."""
    s = w.scan
    ms, launches = [], []
    with torch.cuda.stream(stream):
        for k in range(warmup + steps):
            ctx.set_particles(w.x, w.y, w.theta)
            flush.fill_(k & 0xff)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            if stage_in_step:  # a real filter gets a new scan every update: staging it is part of the step
                ctx.filter_update_scan_async(w.curr, w.prev, 100 + k, w.u0, s.ranges, s.angle_min, s.angle_inc, s.range_max)
            else:
                ctx.filter_update_async(w.curr, w.prev, 100 + k, w.u0)
            e1.record(stream)
            torch.cuda.synchronize()
            if k >= warmup:
                ms.append(e0.elapsed_time(e1))
                launches.append(ctx.stats().kernels_launched)
    return ms, launches, ctx.stats()


def sub_record(name, w, dev_index, stream, flush, steps, torch):
    """A named configuration as a sub-record of the N = 1 line: its own ctx, its own map."""
    from amclj_b200 import _lib
    with _lib.Context(dev_index, "NS") as c:
        c.set_stream(stream.cuda_stream)
        c.set_map(w.grid)
        c.set_particles(w.x, w.y, w.theta)
        c.stage_scan(w.scan.ranges, w.scan.angle_min, w.scan.angle_inc, w.scan.range_max)
        ms, launches, st = time_single(c, w, stream, flush, steps, 3, torch)
        nb = len(w.scan.ranges)
        m = float(np.mean(ms))
        rec = {"workload": w.name, "particles": w.n, "beams": nb, "ms_per_step": m, "value": w.n * nb / (m * 1e-3),
               "unit": UNIT, "updates_per_s": 1e3 / m, "steps": len(ms), "kernels_per_update": int(launches[-1]),
               "sensor_path": SENSOR_PATHS.get(int(st.sensor_path), str(st.sensor_path)),
               "table_fill_ms_once": float(st.ms_table_fill), "table_bytes": int(st.table_bytes)}
        c.set_stream(None)
    return rec


def run_b200(args):
    import torch
    import torch.distributed as dist

    from amclj_b200 import _lib, workloads
    from amclj_b200.sharded import ShardedFilter

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if args.gpus > 1 and world == 1:
        raise SystemExit("N > 1 runs one rank per GPU: launch with python -m torch.distributed.run --nproc-per-node N "
                         "(one process driving several GPUs is amclj_create_multi: tests/c/mgpu_harness.c, tools/mgpu_bench.py)")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)  # set-up and the timing max only: not inside an update

    if args.workload == "c4share":
        w = headline_workload(world, rank)
        w_all = headline_workload(world) if (world > 1 and rank == 0) else None
    else:
        if world > 1:
            raise SystemExit("--workload other than c4share is an N = 1 experiment")
        w = getattr(workloads, args.workload)()
        w_all = None
    s = w.scan
    nb = len(s.ranges)
    ctx = _lib.Context(local, "NS")
    ctx.set_params(_lib.params("NS", df_mode=args.df_mode))
    stream = torch.cuda.Stream(device=dev)
    ctx.set_stream(stream.cuda_stream)
    ctx.set_map(w.grid)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    filt = None
    ctx.set_particles(w.x, w.y, w.theta)
    ctx.stage_scan(s.ranges, s.angle_min, s.angle_inc, s.range_max)
    if world > 1:
        filt = ShardedFilter(ctx, rank, world, w.n * world, dev, stream=stream, mode=args.shard_mode)
        filt.connect()

    def one_step(k):
        # a real filter gets a new scan every update: staging it (host beam table + upload) is part of the step.  One
        # library call does both; with the geometry of the previous scan it enqueues the motion kernel first and builds
        # the table while that runs (amclj_filter_update_scan_async / amclj_shard_update_scan_async)
        if filt and args.shard_mode == "mailbox":
            filt.update(w.curr, w.prev, seed=100 + k, u0=w.u0, scan=s)
        elif filt:
            ctx.stage_scan(s.ranges, s.angle_min, s.angle_inc, s.range_max)
            filt.update(w.curr, w.prev, seed=100 + k, u0=w.u0)
        elif args.two_calls:
            ctx.stage_scan(s.ranges, s.angle_min, s.angle_inc, s.range_max)
            ctx.filter_update_async(w.curr, w.prev, 100 + k, w.u0)
        else:
            ctx.filter_update_scan_async(w.curr, w.prev, 100 + k, w.u0, s.ranges, s.angle_min, s.angle_inc, s.range_max)

    def restore():
        ctx.set_particles(w.x, w.y, w.theta)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- parity of the N > 1 update, driver-visible: all ranks' particles after one sharded update vs the
    # same update of ALL particles on one GPU (rank 0, a second ctx) --------------------------------------
    parity = None
    if world > 1:
        with torch.cuda.stream(stream):
            restore()
            one_step(-7)
        barrier()
        mine = torch.from_numpy(np.stack(ctx.get_particles()[:3], 1)).to(dev)
        got = [torch.empty_like(mine) for _ in range(world)] if rank == 0 else None
        dist.gather(mine, got, dst=0)
        if rank == 0:
            full = torch.cat(got).cpu().numpy()
            with _lib.Context(local, "NS") as ref:
                ref.set_map(w.grid)
                ref.set_particles(w_all.x, w_all.y, w_all.theta)
                ref.stage_scan(s.ranges, s.angle_min, s.angle_inc, s.range_max)
                ref.filter_update_async(w.curr, w.prev, 100 - 7, w.u0)
                ref.synchronize()
                rx, ry, rt, _ = ref.get_particles()
            bad = int(((full[:, 0] != rx) | (full[:, 1] != ry) | (full[:, 2] != rt)).sum())
            parity = {"mismatching_slots": bad, "n": int(len(rx)),
                      "check": f"particles of all {world} ranks after one sharded update (this run's entry point, seed 93) "
                               f"vs the same update of all {len(rx)} particles on one GPU: x, y, theta compared bit for bit"}
        del got, mine
        barrier()

    with torch.cuda.stream(stream):
        for k in range(args.warmup):
            restore()
            one_step(k)
        barrier()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        total_ms_lib, launches, paths = [], [], []
        with ClockSampler(local) as clocks:
            t_wall0 = time.perf_counter()
            for k in range(args.steps):
                restore()
                flush.fill_(k & 0xff)
                barrier()
                if filt and args.shard_mode == "mailbox":
                    # the barrier's release made precise on the DEVICE time line: the ranks' hosts leave the NCCL
                    # barrier tens of microseconds apart, and with one barrier per step that skew would be timed as
                    # part of every step (a rank that starts early waits inside the update for the one that starts late)
                    ctx.shard_rendezvous_async()
                ev[k][0].record(stream)
                one_step(args.warmup + k)
                ev[k][1].record(stream)
                barrier()
                st = ctx.stats()
                total_ms_lib.append(st.ms_total)
                launches.append(st.kernels_launched)
                paths.append(st.sensor_path)
            t_wall = time.perf_counter() - t_wall0
        step_ms = [a.elapsed_time(b) for a, b in ev]
    dev_ms = float(np.sum(step_ms))
    if world > 1:
        t = torch.tensor([dev_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms = float(t.item())
    ms_per_step = dev_ms / args.steps
    n_total = w.n * world
    value = n_total * nb / (ms_per_step * 1e-3)

    # ---- stage and dominant-kernel times: K more steps with event records at the stage boundaries and around
    # the dominant sensor kernel (collect_stats bit 1; the headline loop above runs without them) -------------
    sensor_ms, main_ms, motion_ms, resample_ms = [], [], [], []
    ctx.set_params(_lib.params("NS", collect_stats=2, df_mode=args.df_mode))
    ctx.stage_scan(s.ranges, s.angle_min, s.angle_inc, s.range_max)
    with torch.cuda.stream(stream):
        for k in range(min(args.steps, 10)):
            restore()
            flush.fill_(k & 0xff)
            barrier()
            one_step(args.warmup + k)
            barrier()
            st2 = ctx.stats()
            main_ms.append(st2.ms_sensor_main)
            sensor_ms.append(st2.ms_sensor)
            motion_ms.append(st2.ms_motion)
            resample_ms.append(st2.ms_resample)
    ctx.set_params(_lib.params("NS", df_mode=args.df_mode))
    ctx.stage_scan(s.ranges, s.angle_min, s.angle_inc, s.range_max)

    # ---- e2e: host PoseArray in, resampled PoseArray out, through the C ABI (N > 1: every rank its shard) ----
    def e2e_leg(kind):
        """kind: "pinned" (cudaHostAlloc'ed buffers), "pageable" (plain malloc), "registered" (plain malloc'ed buffers
        page-locked once with amclj_host_register: what the JNA shim does with its Memory objects)."""
        bufs = [torch.empty(w.n, dtype=torch.float32) for _ in range(4)]
        if kind == "pinned":
            bufs = [b.pin_memory() for b in bufs]
        nx, ny, nt, nw = (b.numpy() for b in bufs)
        if kind == "registered":
            for a in (nx, ny, nt, nw):
                ctx.host_register(a)
        e_ms = []
        for k in range(3 + min(args.steps, 10)):
            nx[:], ny[:], nt[:] = w.x, w.y, w.theta
            flush.fill_(k & 0xff)
            barrier()
            t0 = time.perf_counter()
            if filt:
                ctx.shard_update_host(nx, ny, nt, nw, w.curr, w.prev, s.ranges, s.angle_min, s.angle_inc, s.range_max,
                                      w.u0, seed=500 + k)
            else:
                ctx.filter_update_host(nx, ny, nt, nw, w.curr, w.prev, s.ranges, s.angle_min, s.angle_inc,
                                       s.range_max, w.u0, seed=500 + k)
            barrier()
            e_ms.append(1e3 * (time.perf_counter() - t0))
        if kind == "registered":
            for a in (nx, ny, nt, nw):
                ctx.host_unregister(a)
        m = float(np.mean(e_ms[3:]))
        if world > 1:
            t = torch.tensor([m], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            m = float(t.item())
        return m

    e2e = None
    if not (filt and args.shard_mode != "mailbox"):
        m_pin, m_page, m_reg = e2e_leg("pinned"), e2e_leg("pageable"), e2e_leg("registered")
        e2e = {"value": n_total * nb / (m_pin * 1e-3), "unit": UNIT,
               "h2d_bytes_per_step": int(world * (3 * 4 * w.n + 6 * 4 * nb)), "d2h_bytes_per_step": int(world * 4 * 4 * w.n),
               "ms_per_step": m_pin,
               "pageable": {"value": n_total * nb / (m_page * 1e-3), "ms_per_step": m_page,
                            "note": "the same call with plain malloc'ed host buffers (what a JNA Memory is)"},
               "pageable_registered": {"value": n_total * nb / (m_reg * 1e-3), "ms_per_step": m_reg,
                                       "note": "plain malloc'ed host buffers page-locked once with amclj_host_register "
                                               "(what the JNA shim does with its Memory objects)"},
               "call": ("amclj_shard_update_host on every rank (its shard)" if filt else "amclj_filter_update_host")
                       + ": pinned host buffers, wall clock around the synchronous call, max over ranks"}

    # ---- diagnostic pass for the reference-cell / probe counters (untimed; the walking path visits the cells
    # the counters count; N > 1: local, no exchange) ------------------------------------------------------------
    ctx.set_params(_lib.params("NS", collect_stats=1, df_mode=args.df_mode))
    restore()
    ctx.stage_scan(s.ranges, s.angle_min, s.angle_inc, s.range_max)
    ctx.sensor_update(s.ranges, s.angle_min, s.angle_inc, s.range_max)
    sd = ctx.stats()
    ctx.set_params(_lib.params("NS", df_mode=args.df_mode))
    ctx.stage_scan(s.ranges, s.angle_min, s.angle_inc, s.range_max)

    line = None
    if rank == 0:
        peak, peak_src = peaks()
        props = torch.cuda.get_device_properties(dev)
        sm_count = props.multi_processor_count
        cl = clocks.summary()
        sm_hz = 1e6 * float(cl.get("sm_mhz") or 1965)
        stage_ms = float(np.mean(sensor_ms))
        k_ms = float(np.mean(main_ms)) or stage_ms
        path = int(paths[-1]) if paths else -1
        kernel = KERNELS.get(path, "amclj::sensor_kernel")
        n_rays = w.n * nb
        # committed ncu capture of this kernel on this workload (profiles/): warp instructions and DRAM bytes per launch
        prof, pfile = None, ROOT / "profiles" / "dominant_kernel_metrics.json"
        if pfile.exists():
            for rec in json.loads(pfile.read_text()).get("kernels", []):
                if rec.get("kernel") == kernel and rec.get("particles") == w.n and rec.get("beams") == nb:
                    prof = rec
        issue_peak = sm_count * 4 * sm_hz
        st_last = ctx.stats()
        tbl_bytes = int(st_last.table_bytes) if path == 3 else 0
        # algorithmic HBM bytes of one launch of the dominant kernel: persistent tables: every
        # table touched once + per particle 16 B look-up state + 4 B permutation + 8 B accumulator; walk kernel: 16 B
        alg_bytes = float(tbl_bytes + 28 * w.n) if path == 3 else 16.0 * w.n
        hbm = {"achieved": alg_bytes / (k_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
               "frac": alg_bytes / (k_ms * 1e-3) / 1e9 / peak, "peak_source": peak_src,
               "algorithmic_bytes_per_launch": alg_bytes,
               "traffic": None if not prof else prof.get("dram_bytes_read", 0) + prof.get("dram_bytes_write", 0)}
        roofline = {"bound": "issue", "unit": "warp-inst/s", "peak": issue_peak,
                    "peak_source": f"{sm_count} SMs x 4 schedulers x {sm_hz / 1e6:.0f} MHz (median SM clock sampled during the timed region)",
                    "achieved": None, "frac": None, "traffic": hbm["traffic"],
                    "kernel": kernel, "kernel_ms": k_ms, "kernel_share_of_step": k_ms / float(np.mean(total_ms_lib)),
                    "sensor_stage_ms": stage_ms, "sensor_path": SENSOR_PATHS.get(path, str(path)),
                    "rays_per_launch": n_rays, "rays_per_s": n_rays / (k_ms * 1e-3),
                    "scheduler_cycles_per_ray": (k_ms * 1e-3) * issue_peak / n_rays,
                    "reference_cells_per_ray": sd.cells_visited / max(sd.rays, 1),
                    "field_probes_per_ray": sd.probes / max(sd.rays, 1),
                    "reference_cells_per_s": (sd.cells_visited / max(sd.rays, 1)) * n_rays / (k_ms * 1e-3),
                    "hit_fraction": sd.hits / max(sd.rays, 1),
                    "hbm": hbm,
                    "note": "instruction issue binds this kernel, HBM is secondary: frac = warp instructions per launch "
                            "(ncu smsp__inst_executed.sum of the committed capture) / (SMs x 4 x clock x kernel time measured here)"}
        if prof:
            wi = float(prof["warp_instructions"])
            roofline.update({"achieved": wi / (k_ms * 1e-3), "frac": wi / (k_ms * 1e-3) / issue_peak,
                             "warp_instructions_per_launch": wi, "warp_instructions_per_ray": wi / n_rays,
                             "profile": prof.get("source"),
                             # which pipe those instructions queue for (ncu, same capture): on B200 integer / logic /
                             # compare / select instructions issue every second cycle on the ALU pipe (tools/pipe_rates.cu)
                             "pipes_under_ncu": {"issue_active_pct": prof.get("issue_active_pct"),
                                                 "alu_pipe_pct": prof.get("alu_pipe_pct"), "fma_pipe_pct": prof.get("fma_pipe_pct")}})
        cpu = None if args.quick else cpu_baseline(headline_workload(1) if args.workload == "c4share" else w)
        configs = None
        if world == 1 and not args.quick and args.workload == "c4share":
            sub_steps = min(args.steps, 10)
            configs = {"c2": sub_record("c2", workloads.c2(), local, stream, flush, sub_steps, torch),
                       "c3": sub_record("c3", workloads.c3(), local, stream, flush, sub_steps, torch)}
            if not args.skip_c5:
                configs["c5"] = sub_record("c5", workloads.c5(), local, stream, flush, min(sub_steps, 5), torch)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_of(w, world) if args.workload == "c4share" else {"workload": w.name, "particles_per_gpu": w.n, "beams": nb},
            "updates_per_s": 1e3 / ms_per_step, "particles_total": n_total,
            "e2e": e2e, "gpu_launches": int(np.sum(launches)), "kernels_per_update": int(launches[-1]),
            "roofline": roofline, "cpu_baseline": cpu, "clocks": cl, "parity": parity,
            "exchange": None if world == 1 else (
                "peer-memory mailboxes inside the library's kernels (amclj_shard_update_async): 0 NCCL calls per update"
                if args.shard_mode == "mailbox" else f"torch.distributed NCCL ({args.shard_mode})"),
            "table_fill_ms_once": float(st_last.ms_table_fill), "table_bytes": int(st_last.table_bytes),
            "wall_s_timed_loop": t_wall,
            "stage_ms": {"motion": float(np.mean(motion_ms)), "sensor": stage_ms, "sensor_main_kernel": k_ms,
                         "normalise_resample_exchange": float(np.mean(resample_ms)),
                         "total_lib_events": float(np.mean(total_ms_lib)),
                         "note": "library events on rank 0; motion / sensor / resample from a separate pass with stage events"},
            "configs": configs,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    ctx.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c4share", choices=["c4share", "c1", "c2", "c3", "c5"])
    ap.add_argument("--cpu-sample", type=int, default=50_000, help="particles per step of the reference arm")
    ap.add_argument("--shard-mode", default="mailbox", choices=["mailbox", "pull", "a2a"])
    ap.add_argument("--df-mode", type=int, default=0, help="amclj_params.df_mode (0 auto)")
    ap.add_argument("--quick", action="store_true", help="skip the CPU baseline leg and the sub-records (tuning sweeps)")
    ap.add_argument("--two-calls", action="store_true", help="N = 1: amclj_stage_scan + amclj_filter_update_async instead of the fused call (A/B)")
    ap.add_argument("--skip-c5", action="store_true", help="leave the C5 sub-record out (it builds an 8192^2 map)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
