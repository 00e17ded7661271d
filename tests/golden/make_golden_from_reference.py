"""Golden vectors produced by EXECUTING THE REFERENCE'S OWN SOURCE TEXT.

  python tests/golden/make_golden_from_reference.py        (needs /root/reference; writes
                                                            tests/golden/amclj_reference_golden.json)

jvia/amclj cannot run here (Clojure on the JVM, no JVM in the image) and its tests pin nothing on the
hot path.  tests/golden/clj_eval.py is a small evaluator for the part of Clojure the hot path is written
in; this script loads src/amclj/{quaternions,map,motion,sensor,pf}.clj with it — verbatim, nothing
retyped — and calls the reference's functions on the reference's own dev fixtures (sensor.clj:26-74: the
500-range scan and the particle at (15.9, 19.5); core.clj:56-63: the initial pose) and on seeded random
inputs, with sample-normal / sample-uniform / Math/random bound to injected streams and
incanter `sel` bound to data[y][x].  Every expected value in the output file therefore comes out of the
reference's text.  tests/test_reference_text.py holds the CPU oracle to these vectors (cells exact,
doubles to 1e-12) and, when /root/reference is present, re-runs a slice of this script to show the file
is reproducible.

LaserScan and OccupancyGrid fields are float32 on the ROS wire (sensor_msgs/LaserScan, nav_msgs/MapMetaData):
the scan literals of sensor.clj:26-67 are widened from float32 here, as rosjava would deliver them.
"""
from __future__ import annotations

import json
import math
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(HERE))
import clj_eval as ce  # noqa: E402
from amclj_b200 import maps  # noqa: E402  (host-side map ingest only: launch/lgfloor.pgm/.yaml -> OccupancyGrid)

K = ce.K
f32 = lambda v: float(np.float32(v))  # noqa: E731


def occupancy_grid(lg):
    """nav_msgs.clj:25-38: :data = matrix of doubles [height][width], :info = MapMetaData (resolution float32)."""
    return {K("info"): {K("resolution"): float(lg.info.resolution), K("width"): int(lg.info.width),
                        K("height"): int(lg.info.height)},
            K("data"): lg.data.astype(np.float64).tolist()}


def laser_scan(ranges, angle_min, angle_inc, range_max):
    return {K("angleMin"): f32(angle_min), K("angleIncrement"): f32(angle_inc), K("rangeMax"): f32(range_max),
            K("ranges"): ce.VecForm(f32(r) for r in ranges)}


def quat(x, y, z, w):
    return {K("x"): x, K("y"): y, K("z"): z, K("w"): w}


def pose(x, y, q):
    return {K("position"): {K("x"): x, K("y"): y, K("z"): 0}, K("orientation"): q}


def qd(q):
    return {k: q[K(k)] for k in "xyzw"}


class Tracer:
    """Wraps amclj.map/uninhabitable? (what extract-line calls per cell, sensor.clj:117-118) and
    amclj.sensor/map-range to record, per ray: cells tested, the last cell tested, the returned range."""

    def __init__(self, it):
        self.it = it
        self.mapns, self.sns = it.namespaces["amclj.map"].vars, it.namespaces["amclj.sensor"].vars
        self.orig_un, self.orig_mr = self.mapns["uninhabitable?"], self.sns["map-range"]
        self.rays, self.cur = [], None
        self.mapns["uninhabitable?"] = self._un
        self.sns["map-range"] = self._mr

    def _un(self, m, x, y):
        r = self.orig_un(m, x, y)
        if self.cur is not None:
            self.cur["cells"] += 1
            self.cur["last"] = [int(x), int(y)]
            self.cur["blocked"] = bool(r)
        return r

    def _mr(self, *a):
        self.cur = {"cells": 0, "last": None, "blocked": False}
        rng = self.orig_mr(*a)
        self.cur["range"] = float(rng)
        self.rays.append(self.cur)
        self.cur = None
        return rng

    def take(self):
        out, self.rays = self.rays, []
        return out

    def close(self):
        self.mapns["uninhabitable?"], self.sns["map-range"] = self.orig_un, self.orig_mr


def sensor_cases(it, lg, grid, quick=False):
    S = it.namespaces["amclj.sensor"].vars
    dev = S["scan"]
    dev_scan = laser_scan(dev[K("ranges")], dev[K("angleMin")], dev[K("angleIncrement")], dev[K("rangeMax")])
    rng = np.random.Generator(np.random.PCG64(2024))
    scans = {"dev500": dev_scan,
             "syn180": laser_scan(rng.uniform(0.2, 5.6, 180), -math.pi / 2, math.pi / 179, 5.6),
             "syn360": laser_scan(np.where(rng.random(360) < 0.15, 5.6, rng.uniform(0.0, 5.6, 360)), -math.pi,
                                  2 * math.pi / 360, 5.6)}
    res = float(lg.info.resolution)
    free = np.argwhere(lg.data == 0)
    poses = [("fixture particle sensor.clj:69-74", S["particle"]),
             ("initial pose core.clj:56-63", pose(16.3, 15.8, quat(0, 0, 0.85, 0.519)))]
    n_rand = 12 if quick else 220
    for k in range(n_rand):
        cy, cx = free[rng.integers(len(free))]
        h = float(rng.uniform(-math.pi, math.pi))
        x, y = f32((cx + rng.uniform(-0.45, 0.45)) * res), f32((cy + rng.uniform(-0.45, 0.45)) * res)
        poses.append((f"uniform over free space #{k}", pose(x, y, quat(0.0, 0.0, math.sin(h / 2), math.cos(h / 2)))))
    occ, unk = np.argwhere(lg.data == 100), np.argwhere(lg.data == -1)
    W, H = lg.info.width, lg.info.height
    specials = [("on an occupied cell", occ[77][1] * res, occ[77][0] * res), ("on an unknown cell", unk[5000][1] * res, unk[5000][0] * res),
                ("left of the map", -0.3, 5.0), ("below the map", 7.0, -0.2), ("cell (1,1)", 0.05, 0.05),
                ("near the far corner", (W - 2) * res, (H - 2) * res), ("half-cell boundary", 16.325, 15.825),
                ("non-unit quaternion", 14.2, 17.7)]
    for name, x, y in specials:
        q = quat(0.0, 0.0, 0.6, 0.9) if name.startswith("non-unit") else quat(0.0, 0.0, math.sin(0.35), math.cos(0.35))
        poses.append((name, pose(f32(x), f32(y), q)))
    tr = Tracer(it)
    cases = []
    try:
        for name, p in poses:
            for sname, sc in scans.items():
                if sname != "dev500" and not (name.startswith("fixture") or name.startswith("initial") or name.endswith(("#0", "#1", "#2", "#3"))
                                              or not name.startswith("uniform")):
                    continue
                out = ce.call(it, "amclj.sensor/likelihood-field-range-finder-model", sc, p, grid)
                rays = tr.take()
                theta = ce.call(it, "amclj.quaternions/to-heading", p[K("orientation")])
                cases.append({"name": name, "scan": sname, "x": p[K("position")][K("x")], "y": p[K("position")][K("y")],
                              "q": qd(p[K("orientation")]), "theta": theta, "weight": out[K("weight")], "rays": rays})
    finally:
        tr.close()
    scans_out = {k: {"angleMin": v[K("angleMin")], "angleIncrement": v[K("angleIncrement")], "rangeMax": v[K("rangeMax")],
                     "ranges": list(v[K("ranges")])} for k, v in scans.items()}
    return scans_out, cases


def quaternion_cases(it):
    rng = np.random.Generator(np.random.PCG64(7))
    qs = [rng.standard_normal(4).tolist() for _ in range(24)]
    qs += [(np.asarray(q) / np.linalg.norm(q)).tolist() for q in qs[:12]]
    for t in (0.4999, -0.499):  # the pole guards, both sides
        for d in (-1e-9, 1e-9, 1e-4, -1e-4):
            a = math.sqrt(abs(t + d))
            qs.append([a, math.copysign(a, t), 0.0, 0.3])
            qs.append([0.1, 0.0, a, math.copysign(a, t)])
    qs += [[0, 0, 0.21, 0.97], [0, 0, 0.85, 0.519], [0, 0, 0, 1], [0, 0, 1, 0]]
    out = []
    for k, q in enumerate(qs):
        qq = quat(*q)
        h = float(rng.uniform(-3.5, 3.5))
        rot = ce.call(it, "amclj.quaternions/rotate", qq, h)
        out.append({"q": qd(qq), "to_heading": ce.call(it, "amclj.quaternions/to-heading", qq), "heading": h,
                    "rotate": qd(rot), "to_heading_rotated": ce.call(it, "amclj.quaternions/to-heading", rot),
                    "mult_self": qd(ce.call(it, "amclj.quaternions/mult", qq, rot))})
    fh = []
    for h in np.linspace(-6.5, 6.5, 53):
        q = ce.call(it, "amclj.quaternions/from-heading", float(h))
        fh.append({"heading": float(h), "q": qd(q), "back": ce.call(it, "amclj.quaternions/to-heading", q)})
    return out, fh


def map_cases(it, lg, grid):
    res = float(lg.info.resolution)
    rng = np.random.Generator(np.random.PCG64(8))
    w2m = []
    for v in list(rng.uniform(-2, 35, 60)) + [0.025, 0.0249999, 0.075, 16.325, 15.825, -0.025, -0.0250001, 33.1, 31.95]:
        x, y = f32(v), f32(v * 0.77 + 0.01)
        cx, cy = ce.call(it, "amclj.map/world->map", grid, ce.VecForm([x, y]))
        w2m.append({"x": x, "y": y, "cell": [cx, cy]})
    cells = []
    W, H = lg.info.width, lg.info.height
    probes = [(int(rng.integers(0, W)), int(rng.integers(0, H))) for _ in range(80)]
    probes += [(-1, 5), (5, -1), (W + 1, 5), (5, H + 1), (0, 0), (W - 1, H - 1), (326, 316), (318, 390)]
    for x, y in probes:
        cells.append({"x": x, "y": y, "get_cell": ce.call(it, "amclj.map/get-cell", grid, x, y),
                      "uninhabitable": bool(ce.call(it, "amclj.map/uninhabitable?", grid, x, y))})
    # the reference's own bounds test is `(> x width)`: x == width falls through to `sel` (map.clj:22-24)
    edge = []
    for x, y in [(W, 5), (5, H)]:
        try:
            v = ce.call(it, "amclj.map/get-cell", grid, x, y)
            edge.append({"x": x, "y": y, "result": v})
        except IndexError:
            edge.append({"x": x, "y": y, "result": "index out of bounds in sel"})
    return w2m, cells, edge, res


def transform(x, y, yaw):
    return {K("translation"): {K("x"): x, K("y"): y, K("z"): 0}, K("rotation"): quat(0.0, 0.0, math.sin(yaw / 2), math.cos(yaw / 2))}


def motion_cases(it):
    rng = np.random.Generator(np.random.PCG64(9))
    out = []
    controls = [((0.2, 0.0, 0.1), (0.0, 0.0, 0.0)), ((1.0, 2.0, 0.3), (1.0, 2.0, -0.2)), ((0.5, -0.25, 1.0), (0.75, 0.5, 0.8)),
                ((3.0, 1.0, 0.5), (3.0, 1.0, 0.5)), ((-0.4, 0.9, -1.2), (0.1, 0.3, -0.7))]
    for curr, prev in controls:
        poses, normals, results = [], [], []
        for k in range(24):
            h = float(rng.uniform(-1.3, 1.3)) + (math.pi if k % 2 else 0.0)  # away from the misfiring pole guards
            poses.append((f32(rng.uniform(2, 30)), f32(rng.uniform(2, 30)), h))
        st = ce.Streams()
        it.streams = st
        ct, pt = transform(*curr), transform(*prev)
        static = curr == prev
        trans = math.hypot(prev[0] - curr[0], prev[1] - curr[1])
        for (x, y, h) in poses:
            z = [float(np.float32(v)) for v in rng.standard_normal(3)]
            normals.append(z)
            st.normals = [] if static else (z if trans != 0 else [z[0], z[2]])  # motion.clj:36 skips the trans draw
            st.used["normals"] = 0
            p = pose(x, y, quat(0.0, 0.0, math.sin(h / 2), math.cos(h / 2)))
            r = ce.call(it, "amclj.motion/sample-motion-model-odometry", p, ct, pt)
            assert st.used["normals"] == len(st.normals)
            results.append({"x": r[K("position")][K("x")], "y": r[K("position")][K("y")], "q": qd(r[K("orientation")]),
                            "theta": ce.call(it, "amclj.quaternions/to-heading", r[K("orientation")])})
        out.append({"curr": list(curr), "prev": list(prev),
                    "curr_yaw_seen": ce.call(it, "amclj.quaternions/to-heading", ct[K("rotation")]),
                    "prev_yaw_seen": ce.call(it, "amclj.quaternions/to-heading", pt[K("rotation")]),
                    "poses": [list(p) for p in poses], "normals": normals, "out": results})
    # the pole-guard misfire (quaternions.clj:90-93) on a planar heading near +pi/2: the reference reads the
    # particle's heading as 0 there.  Recorded, not gated: DESIGN.md section 4 lists it as a deviation.
    st = ce.Streams(normals=[0.3, -0.2, 0.1])
    it.streams = st
    h = math.pi / 2 - 0.005
    p = pose(10.0, 10.0, quat(0.0, 0.0, math.sin(h / 2), math.cos(h / 2)))
    r = ce.call(it, "amclj.motion/sample-motion-model-odometry", p, transform(0.2, 0.0, 0.1), transform(0.0, 0.0, 0.0))
    misfire = {"heading": h, "heading_seen": ce.call(it, "amclj.quaternions/to-heading", p[K("orientation")]),
               "normals": [0.3, -0.2, 0.1], "x": r[K("position")][K("x")], "y": r[K("position")][K("y")]}
    return out, misfire


def tagged_particles(weights):
    return {K("header"): None,
            K("poses"): ce.VecForm({K("id"): k, K("position"): {K("x"): float(k), K("y"): 0.0, K("z"): 0},
                                    K("orientation"): quat(0, 0, 0, 1), K("weight"): w} for k, w in enumerate(weights))}


def resample_cases(it):
    rng = np.random.Generator(np.random.PCG64(10))
    n = 48
    w = [float(v) for v in (rng.random(n).astype(np.float32) ** 3)]
    total = math.fsum(w)
    # attempts (idx, u = u32 / 2^32); an attempt whose u lies within 1e-6 (relative) of the normalised weight is
    # redrawn, so that the float compare of pf.clj:165 and the oracle's integer compare cannot disagree by rounding
    idx, u32 = [], []
    while len(idx) < 20000:
        i, u = int(rng.integers(0, n)), int(rng.integers(0, 2 ** 32))
        if abs(u / 2.0 ** 32 - w[i] / total) > 1e-6 * max(w[i] / total, 1e-9):
            idx.append(i)
            u32.append(u)
    st = ce.Streams(uniform_ints=idx, randoms=[v / 2.0 ** 32 for v in u32])
    it.streams = st
    out = ce.call(it, "amclj.pf/roulette-selection", tagged_particles(w))
    used = st.used["uniform_ints"]
    assert used == st.used["randoms"]
    roulette = {"weights": w, "attempt_idx": idx[:used], "attempt_u32": u32[:used],
                "selected": [p[K("id")] for p in out[K("poses")]], "weights_out": [p[K("weight")] for p in out[K("poses")]]}
    # tournament-selection pf.clj:182-195
    pairs = [int(v) for v in rng.integers(0, n, 2 * n)]
    st = ce.Streams(uniform_ints=pairs)
    it.streams = st
    out = ce.call(it, "amclj.pf/tournament-selection", tagged_particles(w))
    tournament = {"weights": w, "draws": pairs, "selected": [p[K("id")] for p in out[K("poses")]]}
    # roulette-step pf.clj:169-180 (un-normalised weights)
    steps = []
    for _ in range(12):
        idx2 = [int(v) for v in rng.integers(0, n, 400)]
        us = [float(v) for v in rng.random(400)]
        st = ce.Streams(uniform_ints=idx2, randoms=us)
        it.streams = st
        p = ce.call(it, "amclj.pf/roulette-step", tagged_particles(w))
        k = st.used["uniform_ints"]
        steps.append({"attempt_idx": idx2[:k], "attempt_u": us[:k], "selected": p[K("id")]})
    return roulette, tournament, {"weights": w, "cases": steps}


def pose_estimate_case(it):
    rng = np.random.Generator(np.random.PCG64(11))
    ps, rows = ce.VecForm(), []
    for _ in range(40):
        x, y, h = f32(rng.uniform(0, 30)), f32(rng.uniform(0, 30)), f32(rng.uniform(-3, 3))  # float32-representable: the SoA the kernels hold
        ps.append(pose(x, y, quat(0.0, 0.0, math.sin(h / 2), math.cos(h / 2))))
        rows.append([x, y, h])
    e = ce.call(it, "amclj.pf/pose-estimate", ps)[K("pose")]
    return {"poses": rows, "mean": {"x": e[K("position")][K("x")], "y": e[K("position")][K("y")],
                                    "qz": e[K("orientation")][K("z")], "qw": e[K("orientation")][K("w")],
                                    "qx": e[K("orientation")][K("x")], "qy": e[K("orientation")][K("y")]}}


def init_cases(it, lg, grid):
    """map.clj:59-84 uniform-initialization with injected draws: per proposal x ~ U(0, W), y ~ U(0, H) (cell units),
    heading = rotate (0 0 0 1) by a standard normal; kept when cell (int x, int y) is free; times the resolution."""
    rng = np.random.Generator(np.random.PCG64(12))
    n = 30
    us = [float(v) for v in rng.random(4000)]
    zs = [float(v) for v in rng.standard_normal(2000)]
    st = ce.Streams(uniforms=us, normals=zs)
    it.streams = st
    ps = ce.call(it, "amclj.map/uniform-initialization", grid, n)
    ku, kz = st.used["uniforms"], st.used["normals"]
    out = [{"x": p[K("position")][K("x")], "y": p[K("position")][K("y")], "q": qd(p[K("orientation")]),
            "theta": ce.call(it, "amclj.quaternions/to-heading", p[K("orientation")])} for p in ps]
    return {"n": n, "uniforms": us[:ku], "normals": zs[:kz], "out": out}


def filter_step_case(it, lg, grid, scans):
    """One mcl* iteration's hot path composed the reference's way (pf.clj:229-233): apply-motion-model ->
    apply-sensor-model -> roulette-selection, 40 particles, injected streams."""
    rng = np.random.Generator(np.random.PCG64(13))
    n = 40
    res = float(lg.info.resolution)
    free = np.argwhere(lg.data == 0)
    rows, poses = [], ce.VecForm()
    for _ in range(n):
        cy, cx = free[rng.integers(len(free))]
        h = float(rng.uniform(-1.2, 1.2))
        x, y = f32(cx * res), f32(cy * res)
        rows.append([x, y, h])
        poses.append(pose(x, y, quat(0.0, 0.0, math.sin(h / 2), math.cos(h / 2))))
    curr, prev = (0.2, 0.0, 0.1), (0.0, 0.0, 0.0)
    z = [float(np.float32(v)) for v in rng.standard_normal(3 * n)]
    sc = scans["dev500"]
    scan = laser_scan(sc["ranges"], sc["angleMin"], sc["angleIncrement"], sc["rangeMax"])
    st = ce.Streams(normals=z)
    it.streams = st
    moved = ce.call(it, "amclj.pf/apply-motion-model", ce.VecForm([transform(*curr), transform(*prev)]),
                    {K("header"): None, K("poses"): poses})
    weighted = ce.call(it, "amclj.pf/apply-sensor-model", scan, grid, moved)
    for k, p in enumerate(weighted[K("poses")]):
        p[K("id")] = k
    w = [p[K("weight")] for p in weighted[K("poses")]]
    total = math.fsum(w)
    idx, u32 = [], []
    while len(idx) < 20000:
        i, u = int(rng.integers(0, n)), int(rng.integers(0, 2 ** 32))
        if abs(u / 2.0 ** 32 - w[i] / total) > 1e-6 * max(w[i] / total, 1e-9):
            idx.append(i)
            u32.append(u)
    st.uniform_ints, st.randoms = idx, [v / 2.0 ** 32 for v in u32]
    sel = ce.call(it, "amclj.pf/roulette-selection", weighted)
    used = st.used["uniform_ints"]
    return {"poses": rows, "curr": list(curr), "prev": list(prev), "normals": z, "scan": "dev500",
            "moved": [[p[K("position")][K("x")], p[K("position")][K("y")],
                       ce.call(it, "amclj.quaternions/to-heading", p[K("orientation")])] for p in moved[K("poses")]],
            "weights": w, "attempt_idx": idx[:used], "attempt_u32": u32[:used],
            "selected": [p[K("id")] for p in sel[K("poses")]]}


def generate(quick=False, src="/root/reference/src"):
    it = ce.load_reference(src)
    lg = maps.lgfloor()
    grid = occupancy_grid(lg)
    scans, sensor = sensor_cases(it, lg, grid, quick)
    quats, from_heading = quaternion_cases(it)
    w2m, cells, edge, res = map_cases(it, lg, grid)
    motion, misfire = motion_cases(it)
    roulette, tournament, rstep = resample_cases(it)
    consts = {k: it.namespaces[ns].vars[k] for ns, ks in
              {"amclj.sensor": ["max-range", "max-beams", "z-hit", "z-short", "z-max", "z-rand", "sigma-hit", "lambda-short"],
               "amclj.motion": ["alpha1", "alpha2", "alpha3", "alpha4"], "amclj.pf": ["num-particles", "α-slow", "α-fast"]}.items()
              for k in ks}
    return {"generator": "tests/golden/make_golden_from_reference.py: values computed by evaluating /root/reference/src/amclj/*.clj",
            "constants": consts, "resolution": res, "scans": scans, "sensor": sensor, "quaternion": quats,
            "from_heading": from_heading, "world_to_map": w2m, "cells": cells, "cells_at_the_edge": edge,
            "motion": motion, "motion_pole_guard_misfire": misfire, "roulette": roulette, "tournament": tournament,
            "roulette_step": rstep, "pose_estimate": pose_estimate_case(it), "uniform_initialization": init_cases(it, lg, grid),
            "filter_step": filter_step_case(it, lg, grid, scans)}


if __name__ == "__main__":
    g = generate()
    out = HERE / "amclj_reference_golden.json"
    out.write_text(json.dumps(g))
    print(out, out.stat().st_size, "bytes;", len(g["sensor"]), "sensor cases,", sum(len(c["rays"]) for c in g["sensor"]), "rays")
    print("fixture particle REF weight:", g["sensor"][0]["weight"])
