"""Reported (not gated) statistic of SURVEY.md section 7: how often the fp32 contract and the
double restatement of the Clojure disagree on a ray (start/end cell rounding at .5 boundaries)."""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
import oracle  # noqa: E402
from amclj_b200 import workloads  # noqa: E402

for name, w in (("C2 (gaussian about the true pose)", workloads.c2(3000)), ("C3 (uniform over free space)", workloads.c3(3000))):
    s = w.scan
    p = oracle.params("NS")
    _, _, d32 = oracle.sensor_f32(p, w.grid, s.ranges, s.angle_min, s.angle_inc, s.range_max, w.x, w.y, w.theta,
                                  debug=True, want_raw=False)
    L = oracle.lib()
    import ctypes as C
    n, nb = len(w.x), len(s.ranges)
    diff = 0
    steps64 = np.zeros((n, nb), np.int32)
    hit64 = np.zeros((n, nb), np.uint8)
    for i in range(n):
        for k in range(nb):
            r = oracle.map_range_f64(p, w.grid, float(w.x[i]), float(w.y[i]), float(w.theta[i]),
                                     s.angle_min + k * s.angle_inc, s.range_max)
            steps64[i, k], hit64[i, k] = r.steps, r.hit
    bad = (steps64 != d32["steps"]) | (hit64 != d32["hit"])
    raw32, _, _ = oracle.sensor_f32(p, w.grid, s.ranges, s.angle_min, s.angle_inc, s.range_max, w.x, w.y, w.theta)
    raw64, _ = oracle.sensor_f64(p, w.grid, s.ranges, s.angle_min, s.angle_inc, s.range_max, w.x, w.y, w.theta)
    rel = np.abs(raw32 - raw64) / np.abs(raw64)
    print(f"{name}: rays {n * nb}, (hit, steps) differ on {bad.mean():.2e} of rays; "
          f"weights within 1e-5 for {(rel <= 1e-5).mean():.4f} of particles (median rel err {np.median(rel):.1e})")
