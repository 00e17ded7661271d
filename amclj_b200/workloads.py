"""Synthetic inputs of the BASELINE.json configurations C1-C5, seeded with
numpy ``Generator(PCG64(seed))``: map 5, particles 1, scan 2, motion normals 3, resample u0 4.
Host-side numpy only; the same arrays feed the CUDA path and the CPU oracle in the tests.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

from . import maps

SEED_PARTICLES, SEED_SCAN, SEED_NORMALS, SEED_U0, SEED_MAP, SEED_ROULETTE = 1, 2, 3, 4, 5, 6

# core.clj:56-63 initial pose: (16.3, 15.8), q.z 0.85 q.w 0.519
TRUE_POSE = (16.3, 15.8, 2.0 * math.atan2(0.85, 0.519))


@dataclass
class Scan:
    """sensor_msgs/LaserScan fields amclj reads (sensor_msgs.clj:8-12)."""
    ranges: np.ndarray
    angle_min: float
    angle_inc: float
    range_max: float


@dataclass
class Workload:
    name: str
    grid: maps.OccupancyGrid
    x: np.ndarray
    y: np.ndarray
    theta: np.ndarray
    scan: Scan
    curr: tuple = (0.2, 0.0, 0.1)   # C2 control: prev = (0,0,0) -> curr = (0.2, 0, 0.1)
    prev: tuple = (0.0, 0.0, 0.0)
    u0: float = 0.0
    meta: dict = field(default_factory=dict)

    @property
    def n(self):
        return len(self.x)

    def normals(self):
        rng = np.random.Generator(np.random.PCG64(SEED_NORMALS))
        return rng.standard_normal((self.n, 3)).astype(np.float32)


def march_ranges(grid, pose, angles, range_max, step_cells=0.25):
    """Plausible laser ranges from ``pose``: march each beam in quarter-cell steps until a
    non-free cell (data synthesis only, not a checker)."""
    res = float(grid.info.resolution)
    W, H = grid.info.width, grid.info.height
    blocked = grid.data != 0
    x0, y0, th = pose
    a = th + np.asarray(angles, np.float64)
    out = np.full(len(a), float(range_max))
    alive = np.ones(len(a), bool)
    nsteps = int(range_max / (res * step_cells)) + 1
    for k in range(nsteps):
        d = k * res * step_cells
        cx = np.floor((x0 + d * np.cos(a)) / res + 0.5).astype(np.int64)
        cy = np.floor((y0 + d * np.sin(a)) / res + 0.5).astype(np.int64)
        oob = (cx < 0) | (cx >= W) | (cy < 0) | (cy >= H)
        hit = alive & (oob | blocked[np.clip(cy, 0, H - 1), np.clip(cx, 0, W - 1)])
        out[hit] = d
        alive &= ~hit
        if not alive.any():
            break
    return out


def make_scan(grid, pose, n, angle_min, angle_inc, range_max, noise_sd=0.05):
    angle_min, angle_inc, range_max = (float(np.float32(v)) for v in (angle_min, angle_inc, range_max))
    ang = angle_min + np.arange(n) * angle_inc
    r = march_ranges(grid, pose, ang, range_max)
    rng = np.random.Generator(np.random.PCG64(SEED_SCAN))
    noisy = r + rng.normal(0.0, noise_sd, n)
    noisy = np.where(r >= range_max, range_max, np.clip(noisy, 0.0, range_max))
    return Scan(noisy.astype(np.float32), angle_min, angle_inc, range_max)


def gaussian_particles(n, pose=TRUE_POSE, sd_xy=0.1, sd_theta=0.0, seed=SEED_PARTICLES):
    """map.clj:87-111 gaussian-initialization ([0 0.1] translation, heading sd 0)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    x = (pose[0] + sd_xy * rng.standard_normal(n)).astype(np.float32)
    y = (pose[1] + sd_xy * rng.standard_normal(n)).astype(np.float32)
    t = (pose[2] + sd_theta * rng.standard_normal(n)).astype(np.float32)
    return x, y, t


def uniform_particles(grid, n, seed=SEED_PARTICLES, ref_heading=False):
    """map.clj:59-84 uniform-initialization: cell-uniform real coordinates accepted when
    cell (int x, int y) is free, times the resolution; heading U(-pi, pi) (NS) or N(0,1)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    W, H, res = grid.info.width, grid.info.height, float(grid.info.resolution)
    free = grid.data == 0
    xs, ys = [], []
    need = n
    while need > 0:
        m = int(need * W * H / max(int(free.sum()), 1) * 1.2) + 1024
        px, py = rng.uniform(0, W, m), rng.uniform(0, H, m)
        ok = free[py.astype(np.int64), px.astype(np.int64)]
        xs.append(px[ok][:need])
        ys.append(py[ok][:need])
        need -= len(xs[-1])
    x = (np.concatenate(xs) * res).astype(np.float32)
    y = (np.concatenate(ys) * res).astype(np.float32)
    t = (rng.standard_normal(n) if ref_heading else rng.uniform(-math.pi, math.pi, n)).astype(np.float32)
    return x, y, t


def _u0():
    return float(np.random.Generator(np.random.PCG64(SEED_U0)).random())


def c1(n=1000):
    g = maps.lgfloor()
    x, y, t = gaussian_particles(n)
    scan = make_scan(g, TRUE_POSE, 180, -math.pi / 2, math.pi / 179, 5.6)
    return Workload("C1 lgfloor 1k x 180", g, x, y, t, scan, u0=_u0())


def c2(n=100_000):
    g = maps.lgfloor()
    x, y, t = gaussian_particles(n)
    scan = make_scan(g, TRUE_POSE, 360, -math.pi, 2 * math.pi / 360, 5.6)
    return Workload("C2 lgfloor 100k x 360 pose tracking", g, x, y, t, scan, u0=_u0())


def c3(n=1_000_000):
    g = maps.lgfloor()
    x, y, t = uniform_particles(g, n)
    scan = make_scan(g, TRUE_POSE, 360, -math.pi, 2 * math.pi / 360, 5.6)
    return Workload("C3 lgfloor 1M x 360 global localisation", g, x, y, t, scan, u0=_u0())


def c4(n=10_000_000):
    w = c3(n)
    w.name = "C4 lgfloor 10M x 360 sharded"
    return w


_C5_GRID = {}


def c5_grid(size=8192):
    if size not in _C5_GRID:
        _C5_GRID[size] = maps.synthetic_map(size, size, 0.05, SEED_MAP)
    return _C5_GRID[size]


def c5(n=4_000_000, size=8192):
    g = c5_grid(size)
    x, y, t = uniform_particles(g, n)
    free = np.argwhere(g.data == 0)
    cy, cx = free[len(free) // 2]
    pose = (float(cx) * 0.05, float(cy) * 0.05, 0.3)
    scan = make_scan(g, pose, 1080, -3 * math.pi / 4, math.radians(0.25), 30.0)
    return Workload(f"C5 synthetic {size}^2 {n} x 1080", g, x, y, t, scan, u0=_u0())
