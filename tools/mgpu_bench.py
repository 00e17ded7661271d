"""One process, one caller thread, several GPUs: times amclj_mfilter_update_async (amclj_create_multi,
SURVEY 8b) on the C4 per-GPU share — what a JVM host would drive through JNA — next to bench.py's one-rank-
per-GPU numbers.  Device time per rank from the library's own events (max over ranks).

  python tools/mgpu_bench.py [--gpus G] [--steps K]
"""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from amclj_b200 import _lib, workloads  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=2)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--per-gpu", type=int, default=1_250_000)
    a = ap.parse_args()
    w = workloads.c3(a.per_gpu * a.gpus)
    s = w.scan
    nb = len(s.ranges)
    with _lib.MultiContext(list(range(a.gpus)), "NS") as m:
        m.set_map(w.grid)
        m.set_particles(w.x, w.y, w.theta)
        m.stage_scan(s.ranges, s.angle_min, s.angle_inc, s.range_max)
        dev_ms, wall_ms = [], []
        for k in range(3 + a.steps):
            m.set_particles(w.x, w.y, w.theta)
            m.synchronize()
            t0 = time.perf_counter()
            m.stage_scan(s.ranges, s.angle_min, s.angle_inc, s.range_max)
            m.filter_update_async(w.curr, w.prev, 100 + k, w.u0)
            t1 = time.perf_counter()
            m.synchronize()
            t2 = time.perf_counter()
            if k >= 3:
                dev_ms.append(max(m.rank(r).stats().ms_total for r in range(a.gpus)))
                wall_ms.append((1e3 * (t1 - t0), 1e3 * (t2 - t0)))
        ms = float(np.mean(dev_ms))
        print(json.dumps({"entry": "amclj_mfilter_update_async (one process, one thread)", "gpus": a.gpus,
                          "particles_total": w.n, "beams": nb, "ms_per_step_device_max_over_ranks": ms,
                          "value": w.n * nb / (ms * 1e-3), "unit": "beam-evals/s",
                          "host_enqueue_ms": float(np.mean([x[0] for x in wall_ms])),
                          "wall_ms_enqueue_to_sync": float(np.mean([x[1] for x in wall_ms]))}))


if __name__ == "__main__":
    main()
