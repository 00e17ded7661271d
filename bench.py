#!/usr/bin/env python
"""bench.py — the hot path of one filter update (motion -> ray-cast + beam mixture ->
normalise -> systematic resample) on synthetic inputs of BASELINE.json's configurations.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

N = 1: config C2 (lgfloor, 100k particles x 360 beams, pose tracking).  N > 1 (torchrun, one
rank per GPU): the C4 workload (10M particles x 360 beams over 8 GPUs) at its per-GPU share,
1.25M particles per rank, weak scaling; per update an NCCL all-reduce(max) and all-gather(totals),
then every rank pulls its resampled slots from the owning rank over NVLink peer memory
(--shard-mode a2a: generate + NCCL all-to-all instead).

Prints ONE JSON line (rank 0).  `value` = beam evaluations per second with the particle set
resident in HBM; `e2e` = the same through the host-buffer C-ABI call
(amclj_filter_update_host: upload, update, download).  `--impl reference` times the CPU
oracle (the stand-in for the Clojure reference, which has no JVM to run on here) on a bounded
particle sample with all host threads.
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
PER_GPU_SHARDED = 1_250_000  # C4: 10M particles over 8 GPUs
# kernels per update are counted by the library (amclj_stats.kernels_launched): dense cloud on the ray
# tables = motion, histogram+plan, scatter+fill, look-up, finish, scan, gather = 7; walk kernel = motion,
# sensor, combine, scan, gather = 5 (+ 3 tile-order kernels for a spread-out cloud); sharded: + decode-max,
# pull instead of gather, and 2 NCCL collectives that are not counted
SENSOR_PATHS = {-1: "unknown", 0: "walk kernel, particles in map-tile order (spread-out cloud)",
                1: "walk kernel, particles in place (compact cloud)",
                2: "ray tables (dense cloud: one walk per (start cell, end cell), looked up per particle-beam)"}


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
            uuid = None
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


def cpu_pass(w, sample, threads=0):
    """One pass of the oracle (CPU port of the Clojure path: double precision, one cell per
    Bresenham iteration) over `sample` particles of workload `w`: motion + sensor + normalise +
    systematic resample.  Returns (seconds, threads)."""
    import oracle

    po = oracle.params("NS")
    sel = slice(0, sample)
    nrm = w.normals()[sel]
    s = w.scan
    nthreads = oracle.threads() if threads <= 0 else threads
    t0 = time.perf_counter()
    mx, my, mt = oracle.motion_f64(po, w.curr, w.prev, nrm, w.x[sel], w.y[sel], w.theta[sel])
    raw, st = oracle.sensor_f64(po, w.grid, s.ranges, s.angle_min, s.angle_inc, s.range_max, mx, my, mt,
                                threads=nthreads)
    lin, wmax = oracle.linear_weights(raw.astype(np.float32), True)
    oracle.systematic_resample(lin, w.u0)
    return time.perf_counter() - t0, nthreads


def cpu_baseline(w, budget_s=12.0, max_passes=64):
    """Whole-workload passes until about `budget_s` seconds of CPU work are spent."""
    cpu_pass(w, min(w.n, 2000))  # warm the library and the page cache
    secs, nth = [], 1
    while sum(secs) < budget_s and len(secs) < max_passes:
        dt, nth = cpu_pass(w, w.n)
        secs.append(dt)
    return w.n * len(w.scan.ranges) * len(secs) / sum(secs), sum(secs), len(secs), nth


def run_reference(args):
    """`--impl reference`: the reference's CPU implementation of the path on the host cores.
    The Clojure/JVM original cannot run here (no JVM in the image), so this is the oracle port
    with every host thread; each step is a bounded particle sample of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from amclj_b200 import workloads
    import oracle

    oracle.build()
    w = workloads.c2() if args.gpus == 1 else workloads.c3(200_000)
    sample = min(w.n, args.cpu_sample)
    for _ in range(args.warmup):
        cpu_pass(w, min(sample, 2000))
    secs = []
    for _ in range(args.steps):
        dt, nth = cpu_pass(w, sample)
        secs.append(dt)
    value = sample * len(w.scan.ranges) * len(secs) / sum(secs)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(secs) / len(secs),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_of(args, w, sample_note=f"{sample} of {w.n} particles per step"),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": nth, "kind": "port",
                         "sample": f"{sample} particles x {len(w.scan.ranges)} beams per step, oracle f64, "
                                   f"OpenMP {nth} threads (Clojure reference cannot run: no JVM)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def ctx_ring_entries(ctx):
    """Ray-table entries per start cell for the staged scan geometry (0 when the path is unavailable)."""
    return int(ctx.ray_table_entries())


def config_of(args, w, sample_note=None, extra=None):
    c = {"workload": w.name, "map": f"{w.grid.info.width}x{w.grid.info.height}@{float(w.grid.info.resolution):.2f}m",
         "particles_per_gpu": w.n, "beams": len(w.scan.ranges), "range_max_m": w.scan.range_max,
         "preset": "NS (north-star semantics: heading-correct Bresenham, all beams, log-weights, systematic)",
         "l2": "flushed between timed steps (256 MiB memset, untimed); particle set restored before each step (untimed)"}
    if sample_note:
        c["sample"] = sample_note
    if extra:
        c.update(extra)
    return c


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
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    if world == 1:
        w = workloads.c2() if args.workload == "c2" else getattr(workloads, args.workload)()
        if args.sd_xy is not None and args.workload == "c2":
            w.x, w.y, w.theta = workloads.gaussian_particles(w.n, sd_xy=args.sd_xy)
            w.name += f" (sd_xy {args.sd_xy} m)"
        if args.random_heading and args.workload == "c2":
            w.theta = np.random.Generator(np.random.PCG64(9)).uniform(-math.pi, math.pi, w.n).astype(np.float32)
            w.name += " (random headings)"
    else:
        w = workloads.c3(PER_GPU_SHARDED * world)
        lo = rank * PER_GPU_SHARDED
        w.x, w.y, w.theta = (a[lo:lo + PER_GPU_SHARDED].copy() for a in (w.x, w.y, w.theta))
        w.name = f"C4-share lgfloor {PER_GPU_SHARDED} particles/GPU x 360 (10M at 8 GPUs)"
    s = w.scan
    nb = len(s.ranges)
    ctx = _lib.Context(local, "NS")
    ctx.set_params(_lib.params("NS", df_mode=args.df_mode))
    stream = torch.cuda.Stream(device=dev)
    ctx.set_stream(stream.cuda_stream)
    ctx.set_map(w.grid)
    ctx.stage_scan(s.ranges, s.angle_min, s.angle_inc, s.range_max) if world == 1 else None
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    filt = None
    if world > 1:
        ctx.set_particles(w.x, w.y, w.theta)
        filt = ShardedFilter(ctx, rank, world, w.n * world, dev, mode=args.shard_mode)
        filt.stage_scan(s)
        filt.connect()

    def one_step(k):
        if filt:
            filt.update(w.curr, w.prev, seed=100 + k, u0=w.u0)
        else:
            ctx.filter_update_async(w.curr, w.prev, 100 + k, w.u0)

    def restore():
        ctx.set_particles(w.x, w.y, w.theta)
        if filt:
            filt.reset(w.n)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.cuda.stream(stream):
        for k in range(args.warmup):
            restore()
            one_step(k)
        barrier()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        sensor_ms, total_ms_lib, main_ms, launches, paths = [], [], [], [], []
        with ClockSampler(local) as clocks:
            t_wall0 = time.perf_counter()
            for k in range(args.steps):
                restore()
                flush.fill_(k & 0xff)
                barrier()
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

    # ---- stage and dominant-kernel times: the same K steps again with event records at the stage
    # boundaries and around the dominant sensor kernel (collect_stats bit 1; an event between two
    # kernels keeps the second from being staged behind the first, ~2.5 us each, so the headline
    # loop above runs without them) -------------------------------------------------------------
    ctx.set_params(_lib.params("NS", collect_stats=2, df_mode=args.df_mode))
    if filt:
        filt.stage_scan(s)
    else:
        ctx.stage_scan(s.ranges, s.angle_min, s.angle_inc, s.range_max)
    with torch.cuda.stream(stream):
        for k in range(args.steps):
            restore()
            flush.fill_(k & 0xff)
            barrier()
            one_step(args.warmup + k)
            barrier()
            st2 = ctx.stats()
            main_ms.append(st2.ms_sensor_main)
            sensor_ms.append(st2.ms_sensor)
    ctx.set_params(_lib.params("NS", df_mode=args.df_mode))
    if filt:
        filt.stage_scan(s)
    else:
        ctx.stage_scan(s.ranges, s.angle_min, s.angle_inc, s.range_max)

    # ---- e2e: host PoseArray in, resampled PoseArray out, through the C ABI ------------
    e2e = None
    if world == 1:
        hx, hy, ht, hw = (torch.empty(w.n, dtype=torch.float32).pin_memory() for _ in range(4))
        nx, ny, nt, nw = (t.numpy() for t in (hx, hy, ht, hw))
        e_ms = []
        for k in range(args.warmup + args.steps):
            nx[:], ny[:], nt[:] = w.x, w.y, w.theta
            flush.fill_(k & 0xff)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            ctx.filter_update_host(nx, ny, nt, nw, w.curr, w.prev, s.ranges, s.angle_min, s.angle_inc,
                                   s.range_max, w.u0, seed=500 + k)
            e_ms.append(1e3 * (time.perf_counter() - t0))
        e_ms = e_ms[args.warmup:]
        e2e = {"value": w.n * nb / (np.mean(e_ms) * 1e-3), "unit": UNIT,
               "h2d_bytes_per_step": int(3 * 4 * w.n + 4 * nb + 5 * 4 * nb),
               "d2h_bytes_per_step": int(4 * 4 * w.n), "ms_per_step": float(np.mean(e_ms)),
               "call": "amclj_filter_update_host (pinned host buffers; wall clock around the synchronous call)"}
    else:
        e2e = {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
               "note": "sharded particle sets stay device-resident between updates; see the N=1 line for the host-buffer call"}

    # ---- diagnostic pass for the cell / probe counters (untimed; every rank takes part in
    # the collectives of the sharded update) -------------------------------------------
    ctx.set_params(_lib.params("NS", collect_stats=1, df_mode=args.df_mode))
    restore()
    if filt:
        filt.stage_scan(s)
        with torch.cuda.stream(stream):
            filt.update(w.curr, w.prev, seed=1, u0=w.u0)
    else:
        ctx.stage_scan(s.ranges, s.angle_min, s.angle_inc, s.range_max)
        ctx.filter_update_async(w.curr, w.prev, 1, w.u0)
    ctx.synchronize()
    st = ctx.stats()
    same_share = None
    if world > 1:
        # the same per-GPU share as a plain single-GPU update (no collectives), so that the sharded
        # line can be read against an N = 1 run of the SAME workload (the N = 1 bench line is C2)
        ctx.set_params(_lib.params("NS", df_mode=args.df_mode))
        ctx.stage_scan(s.ranges, s.angle_min, s.angle_inc, s.range_max)
        ss = []
        for k in range(3 + min(args.steps, 10)):
            ctx.set_particles(w.x, w.y, w.theta)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(stream):
                e0.record(stream)
                ctx.filter_update_async(w.curr, w.prev, 900 + k, w.u0)
                e1.record(stream)
            torch.cuda.synchronize()
            ss.append(e0.elapsed_time(e1))
        t = torch.tensor([float(np.mean(ss[3:]))], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        same_share = float(t.item())
    line = None
    if rank == 0:
        # ---- roofline of the dominant kernel (sensor: ray-cast + beam mixture) ---------
        peak, peak_src = peaks()
        props = torch.cuda.get_device_properties(dev)
        sm_count = props.multi_processor_count
        sm_hz = 1e6 * float((clocks.summary().get("sm_mhz") or 1965))
        stage_ms = float(np.mean(sensor_ms))          # the whole sensor stage (sort + fill + look-up + finish, or walk + combine)
        k_ms = float(np.mean(main_ms)) or stage_ms      # its dominant kernel alone (events around that launch)
        path = int(paths[-1]) if paths else -1
        tables = path == 2
        # measured ceiling of dependent lookups on this GPU, placed like what the kernel reads:
        # shared memory, L1 (directional planes of a compact cloud, per-cell ray tables) or L2 (large maps)
        placement = 0 if st.map_in_smem else (2 if (tables or w.grid.info.width * w.grid.info.height < (1 << 22)) else 1)
        lookup_peak = ctx.microbench(placement)
        traffic = None
        tfile = ROOT / "profiles" / "sensor_traffic.json"
        if tfile.exists():  # dram bytes per launch from the committed ncu --set full capture of this workload
            t = json.loads(tfile.read_text())
            if t.get("workload") == w.name and world == 1 and t.get("kernel", "").endswith("rt_lookup_kernel") == tables:
                traffic = t["dram_bytes_read"] + t["dram_bytes_write"]
        # sensor stage = read x,y,theta (12 B) + write raw (4 B) per particle;
        # map, tables and beam constants are served on chip
        alg_bytes = 16.0 * w.n
        achieved = alg_bytes / (k_ms * 1e-3) / 1e9
        n_rays = st.rays if st.rays else w.n * nb
        onchip = {"rays_per_s": n_rays / (k_ms * 1e-3),
                  "reference_cells_per_s": st.cells_visited / (k_ms * 1e-3),
                  "cells_per_ray": st.cells_visited / max(st.rays, 1),
                  "hit_fraction": st.hits / max(st.rays, 1),
                  "map_in_smem": int(st.map_in_smem),
                  "issue_peak_warp_inst_per_s": sm_count * 4 * sm_hz,
                  "scheduler_cycles_per_ray": (k_ms * 1e-3) * sm_count * 4 * sm_hz / max(n_rays, 1),
                  "note": "scheduler_cycles_per_ray = kernel time x SMs x 4 schedulers x SM clock / rays: "
                          "the issue-slot budget each ray consumed"}
        if tables:
            onchip.update({
                "table_cells": int(st.table_cells), "table_particles": int(st.table_particles),
                "table_misses": int(st.table_misses),
                "table_entries_per_cell": ctx_ring_entries(ctx),
                "rays_walked_for_tables_per_update": int(st.table_cells) * ctx_ring_entries(ctx),
                "walk_probes_per_update": int(st.probes),
                "table_lookups_per_s": n_rays / (k_ms * 1e-3),
                "lookup_peak_per_s": lookup_peak,
                "lookup_frac": (n_rays / (k_ms * 1e-3)) / lookup_peak,
                "lookup_peak_source": "amclj_microbench: dependent byte lookups, all SMs x 32 warps, 64 KiB table in L1 "
                                      "(each table look-up here carries ~55 more instructions: end point, ring index, mixture)"})
        else:
            onchip.update({
                "probes_per_s": st.probes / (k_ms * 1e-3), "probes_per_ray": st.probes / max(st.rays, 1),
                "lookup_peak_per_s": lookup_peak, "lookup_frac": (st.probes / (k_ms * 1e-3)) / lookup_peak,
                "lookup_peak_source": "amclj_microbench: dependent lookups, all SMs x 32 warps, "
                                      + ["128 KiB nibble table in shared memory", "64 MiB byte table in L2",
                                         "64 KiB byte table in L1"][placement]})
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": traffic, "peak_source": peak_src,
                    "kernel": "amclj::rt_lookup_kernel" if tables else "amclj::sensor_kernel",
                    "kernel_ms": k_ms, "kernel_share_of_step": k_ms / float(np.mean(total_ms_lib)),
                    "sensor_stage_ms": stage_ms, "sensor_path": SENSOR_PATHS.get(path, str(path)),
                    "algorithmic_bytes_per_launch": alg_bytes,
                    "note": "HBM is not what binds this kernel (16 B/particle): it is instruction-issue bound "
                            "(ncu, profiles/: 71-85 % of scheduler cycles issue); see onchip",
                    "onchip": onchip}
        cpu = None
        if world == 1 and not args.quick:
            from amclj_b200 import workloads as wl
            import oracle
            oracle.build()
            wc = wl.c2()
            v, dt, passes, nth = cpu_baseline(wc)
            cpu = {"value": v, "unit": UNIT, "cores": nth, "kind": "port",
                   "sample": f"{passes} passes over all {wc.n} particles x {nb} beams ({dt:.1f} s of CPU work), oracle f64 "
                             f"with OpenMP on {nth} threads (the Clojure reference cannot run here: no JVM)"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_of(args, w, extra={"parallelism": f"particles sharded x{world}, map replicated"}),
            "updates_per_s": 1e3 / ms_per_step, "particles_total": n_total,
            "e2e": e2e, "gpu_launches": int(np.sum(launches)),
            "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks.summary(),
            "wall_s_timed_loop": t_wall,
            "same_share_single_gpu": None if same_share is None else {
                "ms_per_step": same_share, "value_per_gpu": w.n * nb / (same_share * 1e-3),
                "sharded_over_single": same_share / ms_per_step,
                "note": "each rank's share run as a plain 1-GPU update (no collectives); the ratio is the "
                        "efficiency of the sharded update against the same per-GPU work"},
            "stage_ms": {"motion": None, "sensor": stage_ms, "sensor_main_kernel": k_ms,
                         "total_lib_events": float(np.mean(total_ms_lib))},
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    ctx.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=["c1", "c2", "c3", "c4", "c5"])
    ap.add_argument("--cpu-sample", type=int, default=100_000)
    ap.add_argument("--shard-mode", default="pull", choices=["pull", "a2a"])
    ap.add_argument("--sd-xy", type=float, default=None, help="C2 only: std-dev of the particle cloud in metres (policy experiments)")
    ap.add_argument("--random-heading", action="store_true", help="C2 only: uniform random headings (locality experiments)")
    ap.add_argument("--df-mode", type=int, default=0, help="amclj_params.df_mode (0 auto)")
    ap.add_argument("--quick", action="store_true", help="skip the CPU baseline leg (tuning sweeps)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
