"""Golden-vector generator: an independent, literal Python transliteration of the
reference's hot-path Clojure, run on the reference's own dev fixtures.

The reference (jvia/amclj) cannot run here (no JVM) and ships no expected outputs for this
path, so these vectors are the best available pin for oracle/amclj_oracle.c: a second
restatement, written function-for-function from the Clojure source with Python floats
(IEEE doubles, libm) standing in for java.lang.Math.  It reads the 500-range LaserScan
fixture out of the reference source at generation time (sensor.clj:26-67) and needs
/root/reference; the produced tests/golden/amclj_golden.json is committed and is what the
tests read.  Usage:  python tests/golden/make_golden.py

Every function below cites the Clojure it transliterates.  `ns=True` arguments select the
north-star reading of the same function (switches S1-S7).
"""
from __future__ import annotations

import json
import math
import re
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from amclj_b200 import maps  # noqa: E402  (host-side map ingest only)

REF = Path("/root/reference")
f32 = lambda v: float(np.float32(v))  # a ROS float32 field, widened to double

# sensor.clj:11-21
MAX_RANGE = -1.0
MAX_BEAMS = 30
Z_HIT, Z_SHORT, Z_MAX, Z_RAND = 0.95, 0.1, 0.05, 0.05
SIGMA_HIT, LAMBDA_SHORT = 0.2, 0.1
# motion.clj:12-15
ALPHA1, ALPHA2, ALPHA3, ALPHA4 = 0.002, 0.002, 0.02, 0.02
# (the Clojure double literals themselves: the oracle widens amclj_params' floats back to them, decimal_exact)


def java_round(v):  # Math/round
    return math.floor(v + 0.5)


# ---- map.clj ---------------------------------------------------------------------------
def get_cell(grid, x, y):  # map.clj:19-24
    width, height = grid.info.width, grid.info.height
    if x > width or x < 0 or y > height or y < 0:
        return -1
    if x == width or y == height:
        raise IndexError("reference falls through to incanter sel here (unpinned)")
    return int(grid.data[y, x])  # (sel data y x)


def inhabitable(grid, x, y):  # map.clj:48-51
    return get_cell(grid, x, y) == 0


def uninhabitable(grid, x, y):  # map.clj:57
    return not inhabitable(grid, x, y)


def world_to_map(grid, xy):  # map.clj:36-42
    res = float(grid.info.resolution)
    return [java_round(xy[0] / res), java_round(xy[1] / res)]


# ---- quaternions.clj -------------------------------------------------------------------
def to_heading(q):  # quaternions.clj:82-97, q = dict x y z w
    qx, qy, qz, qw = q["x"], q["y"], q["z"], q["w"]
    test = qx * qy + qz * qw
    if test > 0.4999:
        return 2 * math.atan2(qx, qw)
    if test < -0.499:
        return -2 * math.atan2(qx, qw)
    return math.atan2(2 * qz * qw - 2 * qx * qy, 1 - 2 * qz * qz - 2 * qy * qy)


def mult(qa, qb):  # quaternions.clj:7-16
    qaw, qax, qay, qaz = qa["w"], qa["x"], qa["y"], qa["z"]
    qbw, qbx, qby, qbz = qb["w"], qb["x"], qb["y"], qb["z"]
    return {
        "x": qax * qbw + qaw * qbx + qay * qbz + (-(qaz * qby)),
        "y": qaw * qby + (-(qax * qbz)) + qay * qbw + qaz * qbx,
        "z": qaw * qbz + qax * qby + (-(qay * qbx)) + qaz * qbw,
        "w": qaw * qbw - qax * qbx - qay * qby - qaz * qbz,
    }


def rotate(quat, heading):  # quaternions.clj:19-36
    pitch = bank = 0
    c1, s1 = math.cos(heading / 2), math.sin(heading / 2)
    c2, s2 = math.cos(pitch / 2), math.sin(pitch / 2)
    c3, s3 = math.cos(bank / 2), math.sin(bank / 2)
    c1c2, s1s2 = c1 * c2, s1 * s2
    return mult({"w": c1c2 * c3 - s1s2 * s3, "x": c1c2 * s3 + s1s2 * c3,
                 "y": c1 * s2 * c3 - s1 * c2 * s3, "z": s1 * c2 * c3 + c1 * s2 * s3}, quat)


# ---- sensor.clj ------------------------------------------------------------------------
def steep_p(a, b):  # sensor.clj:87-91
    return abs(b[1] - a[1]) > abs(b[0] - a[0])


def bearing(k, scan):  # sensor.clj:93-97
    return scan["angleMin"] + k * scan["angleIncrement"]


def project_pose(xy, brg, max_range):  # sensor.clj:99-103
    return [xy[0] + max_range * math.cos(brg), xy[1] + max_range * math.sin(brg)]


def calc_step(a, b):  # sensor.clj:108-110
    return [1 if a[0] < b[0] else -1, 1 if a[1] < b[1] else -1]


def extract_line(grid, p0, pmax, step, steep, miss_value, ns_termination=False):
    """sensor.clj:112-128.  Returns (range, hit?, tested cell (un-swapped), steps, cells)."""
    x0, y0 = p0
    max_x, max_y = pmax
    x_step, y_step = step
    delta_x, delta_y = abs(x0 - max_x), abs(y0 - max_y)
    delta_err = delta_y
    x, y, error, cells = x0, y0, 0, 0
    while True:
        cells += 1
        cell = (y, x) if steep else (x, y)
        if uninhabitable(grid, *cell):
            return (math.hypot(x - x0, y - y0) * float(grid.info.resolution), True, cell,
                    abs(x - x0), cells)
        if (x == max_x) if ns_termination else (x > max_x):
            return miss_value, False, cell, abs(x - x0), cells
        x = x + x_step
        error = error + delta_err
        if 2 * error >= delta_x:
            y, error = y + y_step, error - delta_x


def map_range(pose3, obs_bearing, grid, range_max, ns=False):  # sensor.clj:130-139
    ox, oy, theta = pose3
    brg = theta + obs_bearing  # computed and, in the reference, unused
    direction = brg if ns else obs_bearing
    length = range_max if ns else MAX_RANGE
    x0, y0 = world_to_map(grid, [ox, oy])
    x1, y1 = world_to_map(grid, project_pose([ox, oy], direction, length))
    steep = steep_p([x1, y1], [x0, y0])
    x0, y0 = (y0, x0) if steep else (x0, y0)
    if steep:
        x1, y1 = y1, x1
    elif not ns:
        x1, y1 = x0, y0  # `(if steep [y1 x1] [x0 y0])`
    xstep, ystep = calc_step([x0, y0], [x1, y1])
    return extract_line(grid, [x0, y0], [x1, y1], [xstep, ystep], steep, length,
                        ns_termination=ns)


def likelihood_star(scan, pose3, grid, ns=False):  # sensor.clj:141-162
    num_scans = len(scan["ranges"])
    step = 1 if ns else int((num_scans - 1) / (MAX_BEAMS - 1))
    sigma = SIGMA_HIT if ns else Z_HIT
    rmax = scan["rangeMax"]
    terms, rays = [], []
    for i in range(0, num_scans, step):
        obs_range = scan["ranges"][i]
        obs_bearing = bearing(i, scan)
        rng, hit, cell, steps, cells = map_range(pose3, obs_bearing, grid, rmax, ns=ns)
        z = obs_range - rng
        pz1 = Z_HIT * math.exp((-(z * z)) / (2 * sigma * sigma))
        pz2 = Z_SHORT * LAMBDA_SHORT * math.exp((-LAMBDA_SHORT) * obs_range) if z < 0 else 0
        pz3 = 1.0 * Z_MAX if obs_range == rmax else 0
        pz4 = Z_RAND * (1.0 / rmax) if obs_range < rmax else 0
        s = pz1 + pz2 + pz3 + pz4
        terms.append(math.log(s) if ns else math.pow(s, 3))
        rays.append({"beam": i, "range": rng, "hit": int(hit), "cell": [int(cell[0]), int(cell[1])],
                     "steps": int(steps), "cells": int(cells)})
    return terms, rays


def likelihood(scan, pose, grid, ns=False):  # sensor.clj:165-173
    theta = to_heading(pose["orientation"])
    terms, rays = likelihood_star(scan, [pose["position"]["x"], pose["position"]["y"], theta],
                                  grid, ns=ns)
    weight = 0.0
    for t in terms:  # (reduce + ...)
        weight = weight + t
    return weight, rays, theta


# ---- motion.clj ------------------------------------------------------------------------
def sample_motion_model_odometry_star(pose, curr, prev, normals):  # motion.clj:24-45
    xn, yn, thn = curr
    x, y, th = prev
    if (x - xn) == 0 and (y - yn) == 0 and (th - thn) == 0:
        return pose
    n = iter(normals)
    sample = lambda var: math.sqrt(var) * next(n)  # motion.clj:21-22, injected draw
    p_theta = to_heading(pose["orientation"])
    rot1 = math.atan2(yn - y, xn - x) - th
    trans = math.hypot(x - xn, y - yn)
    rot2 = thn - th - rot1
    rot1s = rot1 - sample(ALPHA1 * math.pow(rot1, 2) + ALPHA2 * math.pow(rot2, 2))
    if trans == 0:
        transs = 0
    else:
        transs = trans - sample(ALPHA3 * math.pow(trans, 2) + ALPHA4 * math.pow(rot1, 2)
                                + ALPHA4 * math.pow(rot2, 2))
    rot2s = rot2 - sample(ALPHA1 * math.pow(rot2, 2) + ALPHA2 * math.pow(trans, 2))
    return {
        "position": {"x": pose["position"]["x"] + transs * math.cos(p_theta + rot1s),
                     "y": pose["position"]["y"] + transs * math.sin(p_theta + rot1s)},
        "orientation": rotate(pose["orientation"], rot1s + rot2s),
    }


# ---- pf.clj ----------------------------------------------------------------------------
def roulette_selection(weights, attempts):  # pf.clj:156-167 with an injected stream
    mx = len(weights)
    normalizer = 0.0
    for w in weights:
        normalizer = normalizer + w
    n_w = [w / normalizer for w in weights]
    out, used = [], 0
    for idx, u in attempts:
        if len(out) == mx:
            break
        used += 1
        if n_w[idx] > u:
            out.append(idx)
    return out, used, n_w


def pose_estimate(poses):  # pf.clj:67-84 (x, y and the quaternion sums over N)
    n = len(poses)
    sx = sum(p["position"]["x"] for p in poses)
    sy = sum(p["position"]["y"] for p in poses)
    sz = sum(p["orientation"]["z"] for p in poses)
    sw = sum(p["orientation"]["w"] for p in poses)
    return [sx / n, sy / n, sz / n, sw / n]


def parse_fixture_scan():
    src = (REF / "src/amclj/sensor.clj").read_text()
    body = src[src.index("(def scan"):src.index("(def particle")]
    num = lambda key: float(re.search(r":%s\s+(-?[0-9.eE+-]+)" % key, body).group(1))
    ranges = [float(t) for t in re.search(r":ranges\s*\[([^\]]*)\]", body).group(1).split()]
    return {"angleMin": f32(num("angleMin")), "angleIncrement": f32(num("angleIncrement")),
            "rangeMax": f32(num("rangeMax")), "ranges": [f32(r) for r in ranges]}


def main():
    grid = maps.load_map(REF / "launch/lgfloor.yaml")
    assert (grid.data == maps.lgfloor().data).all()
    scan = parse_fixture_scan()
    assert len(scan["ranges"]) == 500
    rng = np.random.Generator(np.random.PCG64(20141))
    free = np.argwhere(grid.data == 0)
    res = float(grid.info.resolution)

    def planar(x, y, z, w):
        return {"position": {"x": x, "y": y}, "orientation": {"x": 0.0, "y": 0.0, "z": z, "w": w}}

    poses = [planar(15.9, 19.5, 0.21, 0.97),   # sensor.clj:69-74 (non-unit quaternion)
             planar(16.3, 15.8, 0.85, 0.519)]  # core.clj:56-63
    for _ in range(10):
        cy, cx = free[rng.integers(len(free))]
        th = float(rng.uniform(-math.pi, math.pi))
        poses.append(planar(f32((cx + rng.uniform(-0.4, 0.4)) * res),
                            f32((cy + rng.uniform(-0.4, 0.4)) * res),
                            math.sin(th / 2), math.cos(th / 2)))
    sensor_cases = []
    for pose in poses:
        entry = {"x": pose["position"]["x"], "y": pose["position"]["y"], "q": pose["orientation"]}
        for name, ns in (("ref", False), ("ns", True)):
            weight, rays, theta = likelihood(scan, pose, grid, ns=ns)
            entry["theta"] = theta
            entry[name] = {"weight": weight, "rays": rays}
        sensor_cases.append(entry)

    # motion: BASELINE config C2 control prev=(0,0,0) -> curr=(0.2,0,0.1), plus edge cases
    motion_cases = []
    for curr, prev in (((0.2, 0.0, 0.1), (0.0, 0.0, 0.0)), ((1.0, 2.0, 0.3), (1.0, 2.0, -0.2)),
                       ((0.5, -0.25, 1.0), (0.75, 0.5, 2.5)), ((3.0, 1.0, 0.5), (3.0, 1.0, 0.5))):
        normals = rng.standard_normal((len(poses), 3)).astype(np.float32)
        outs = []
        for pose, nrm in zip(poses, normals):
            draws = [float(v) for v in nrm]
            trans = math.hypot(prev[0] - curr[0], prev[1] - curr[1])
            use = draws if trans != 0 else [draws[0], draws[2]]  # motion.clj:36 skips a draw
            np_ = sample_motion_model_odometry_star(pose, curr, prev, use)
            outs.append({"x": np_["position"]["x"], "y": np_["position"]["y"],
                         "theta": to_heading(np_["orientation"]), "q": np_["orientation"]})
        motion_cases.append({"curr": curr, "prev": prev, "normals": normals.tolist(), "out": outs})

    # roulette: 64 weights, float semantics (pf.clj:159-166)
    w = rng.random(64).astype(np.float32) ** 3
    att = [(int(rng.integers(0, 64)), float(rng.integers(0, 2 ** 32)) / 2.0 ** 32)
           for _ in range(40000)]
    sel, used, n_w = roulette_selection([float(v) for v in w], att)
    assert len(sel) == 64
    roulette = {"weights": [float(v) for v in w], "attempt_idx": [a[0] for a in att[:used]],
                "attempt_u32": [int(a[1] * 2 ** 32) for a in att[:used]], "selected": sel,
                "attempts_used": used, "normalized": n_w}

    quat_cases = []
    for _ in range(12):
        q = rng.standard_normal(4)
        q[:2] *= 0.05
        q /= np.linalg.norm(q)
        qd = dict(zip("xyzw", (float(v) for v in q)))
        h = float(rng.uniform(-3, 3))
        quat_cases.append({"q": qd, "to_heading": to_heading(qd), "heading": h,
                           "rotated": rotate(qd, h)})
    quat_cases.append({"q": {"x": 0.5, "y": 0.5, "z": 0.5, "w": 0.5},
                       "to_heading": to_heading({"x": 0.5, "y": 0.5, "z": 0.5, "w": 0.5}),
                       "heading": 0.0, "rotated": rotate({"x": 0.5, "y": 0.5, "z": 0.5, "w": 0.5}, 0.0)})

    est = pose_estimate(poses)
    out = {"generator": "tests/golden/make_golden.py", "scan": scan, "sensor": sensor_cases,
           "motion": motion_cases, "roulette": roulette, "quaternion": quat_cases,
           "pose_estimate": {"poses": [[p["position"]["x"], p["position"]["y"],
                                        p["orientation"]["z"], p["orientation"]["w"]] for p in poses],
                             "mean": est},
           "map_counts": {"occupied": int((grid.data == 100).sum()), "free": int((grid.data == 0).sum()),
                          "unknown": int((grid.data == -1).sum())}}
    path = Path(__file__).resolve().parent / "amclj_golden.json"
    path.write_text(json.dumps(out))
    print(path, path.stat().st_size, "bytes")
    print("fixture particle REF weight:", sensor_cases[0]["ref"]["weight"])
    print("fixture particle NS log-weight:", sensor_cases[0]["ns"]["weight"])


if __name__ == "__main__":
    main()
