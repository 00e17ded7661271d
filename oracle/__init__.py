"""ctypes binding of the CPU oracle (oracle/amclj_oracle.c).

TEST INFRASTRUCTURE ONLY — see the header of amclj_oracle.c.  The product package
``amclj_b200`` never imports this module; it is used by tests/, by
``__graft_entry__.smoke()`` and by ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs, as the checker and as the timed CPU stand-in for the (un-runnable) Clojure reference.
Pinned to the reference's own source text: tests/golden/clj_eval.py evaluates the reference's .clj files
and tests/test_reference_text.py holds the oracle to the vectors that produced (the reference ships no
golden vectors of its own on this path and cannot run here: no JVM).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / "libamclj_oracle.so"

f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
i8p = np.ctypeslib.ndpointer(np.int8, flags="C_CONTIGUOUS")
u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
u64p = np.ctypeslib.ndpointer(np.uint64, flags="C_CONTIGUOUS")


class Params(C.Structure):
    """Mirror of ``amclj_params`` (include/amclj_b200.h)."""

    _fields_ = [
        ("s1_use_heading", C.c_int32),
        ("s2_use_scan_range_max", C.c_int32),
        ("s3_proper_endpoint", C.c_int32),
        ("s4_endpoint_termination", C.c_int32),
        ("s5_use_sigma_hit", C.c_int32),
        ("s6_max_beams", C.c_int32),
        ("s7_log_weights", C.c_int32),
        ("s8_resampler", C.c_int32),
        ("s9_pr_rot1_variance", C.c_int32),
        ("s10_apply_laser_offset", C.c_int32),
        ("collect_stats", C.c_int32),
        ("df_mode", C.c_int32),
        ("max_range", C.c_float),
        ("z_hit", C.c_float),
        ("z_short", C.c_float),
        ("z_max", C.c_float),
        ("z_rand", C.c_float),
        ("sigma_hit", C.c_float),
        ("lambda_short", C.c_float),
        ("laser_x", C.c_float),
        ("laser_y", C.c_float),
        ("alpha1", C.c_float),
        ("alpha2", C.c_float),
        ("alpha3", C.c_float),
        ("alpha4", C.c_float),
        ("reserved1", C.c_float * 3),
    ]


def params(preset: str = "NS", **overrides) -> Params:
    """The oracle's own statement of the presets: REF = the Clojure constants and quirks
    (sensor.clj:11-21, motion.clj:12-16, SURVEY section 8 switches); NS = north star."""
    p = Params()
    p.max_range = -1.0
    p.z_hit, p.z_short, p.z_max, p.z_rand = 0.95, 0.1, 0.05, 0.05
    p.sigma_hit, p.lambda_short = 0.2, 0.1
    p.laser_x, p.laser_y = 0.135, 0.0
    p.alpha1, p.alpha2, p.alpha3, p.alpha4 = 0.002, 0.002, 0.02, 0.02
    if preset.upper() == "REF":
        p.s6_max_beams = 30
        p.s8_resampler = 0
    elif preset.upper() == "NS":
        p.s1_use_heading = 1
        p.s2_use_scan_range_max = 1
        p.s3_proper_endpoint = 1
        p.s4_endpoint_termination = 1
        p.s5_use_sigma_hit = 1
        p.s6_max_beams = 0
        p.s7_log_weights = 1
        p.s8_resampler = 1
    else:
        raise ValueError(preset)
    for k, v in overrides.items():
        if not hasattr(p, k):
            raise AttributeError(k)
        setattr(p, k, v)
    return p


class Ray(C.Structure):
    _fields_ = [
        ("x0", C.c_int32), ("y0", C.c_int32), ("x1", C.c_int32), ("y1", C.c_int32),
        ("steep", C.c_int32), ("hit", C.c_int32), ("hit_x", C.c_int32), ("hit_y", C.c_int32),
        ("steps", C.c_int32), ("reserved", C.c_int32), ("cells", C.c_int64),
        ("range", C.c_double),
    ]


class SensorStats(C.Structure):
    _fields_ = [("rays", C.c_uint64), ("cells", C.c_uint64), ("hits", C.c_uint64)]


class Control(C.Structure):
    _fields_ = [("is_static", C.c_int32), ("reserved", C.c_int32), ("rot1", C.c_double),
                ("trans", C.c_double), ("rot2", C.c_double), ("sd_rot1", C.c_double),
                ("sd_trans", C.c_double), ("sd_rot2", C.c_double)]


def build(force: bool = False) -> Path:
    """Compile the oracle with the committed recipe (oracle/Makefile)."""
    src = _HERE / "amclj_oracle.c"
    if force or not _LIB_PATH.exists() or _LIB_PATH.stat().st_mtime < src.stat().st_mtime:
        env = {k: v for k, v in os.environ.items() if k not in ("CC", "CFLAGS")}
        subprocess.run(["make", "-C", str(_HERE), "-B"], check=True, env=env,
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(str(_LIB_PATH))
    PP = C.POINTER(Params)
    L.amclj_oracle_blocked.argtypes = [i8p, C.c_int32, C.c_int32, C.c_int64, C.c_int64]
    L.amclj_oracle_blocked.restype = C.c_int
    L.amclj_oracle_to_heading.argtypes = [C.c_double] * 4
    L.amclj_oracle_to_heading.restype = C.c_double
    L.amclj_oracle_quat_mult.argtypes = [f64p, f64p, f64p]
    L.amclj_oracle_quat_rotate.argtypes = [f64p, C.c_double, f64p]
    L.amclj_oracle_beam_step.argtypes = [C.c_int32, C.c_int32]
    L.amclj_oracle_beam_step.restype = C.c_int32
    L.amclj_oracle_beams_used.argtypes = [C.c_int32, C.c_int32]
    L.amclj_oracle_beams_used.restype = C.c_int32
    L.amclj_oracle_map_range_f64.argtypes = [PP, i8p, C.c_int32, C.c_int32, C.c_float,
                                             C.c_double, C.c_double, C.c_double, C.c_double,
                                             C.c_double, C.POINTER(Ray)]
    L.amclj_oracle_sincosf.argtypes = [C.c_float, C.POINTER(C.c_float), C.POINTER(C.c_float)]
    L.amclj_oracle_beam_table.argtypes = [C.c_int32, C.c_float, C.c_float, C.c_int32, f32p,
                                          f32p, i32p]
    L.amclj_oracle_map_range_f32.argtypes = [PP, i8p, C.c_int32, C.c_int32, C.c_float,
                                             C.c_float, C.c_float, C.c_float, C.c_float,
                                             C.c_float, C.c_float, C.c_float, C.c_float, C.c_double, C.c_double,
                                             C.POINTER(Ray)]
    L.amclj_oracle_sincos_f64.argtypes = [C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.amclj_oracle_sensor_f64.argtypes = [PP, i8p, C.c_int32, C.c_int32, C.c_float, f32p,
                                          C.c_int32, C.c_float, C.c_float, C.c_float, f64p,
                                          f64p, f64p, C.c_int64, f64p, C.POINTER(SensorStats),
                                          C.c_int32]
    L.amclj_oracle_rays_f64.argtypes = [PP, i8p, C.c_int32, C.c_int32, C.c_float, C.c_int32, C.c_float,
                                        C.c_float, C.c_float, f64p, f64p, f64p, C.c_int64,
                                        np.ctypeslib.ndpointer(np.int32, flags="C"),
                                        np.ctypeslib.ndpointer(np.uint8, flags="C"),
                                        np.ctypeslib.ndpointer(np.float64, flags="C"), C.c_int32]
    L.amclj_oracle_sensor_f32.argtypes = [PP, i8p, C.c_int32, C.c_int32, C.c_float, f32p,
                                          C.c_int32, C.c_float, C.c_float, C.c_float, f32p,
                                          f32p, f32p, C.c_int64, C.c_void_p, C.c_void_p,
                                          C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                          C.POINTER(SensorStats), C.c_int32]
    L.amclj_oracle_control.argtypes = [PP, f64p, f64p, C.POINTER(Control)]
    L.amclj_oracle_motion_f64.argtypes = [PP, f64p, f64p, f32p, f64p, f64p, f64p, C.c_int64]
    L.amclj_oracle_motion_f32.argtypes = [PP, f64p, f64p, f32p, f32p, f32p, f32p, C.c_int64]
    L.amclj_oracle_philox.argtypes = [u32p, u32p, u32p]
    L.amclj_oracle_motion_normals.argtypes = [C.c_uint64, C.c_int64, C.c_int64, f32p]
    L.amclj_oracle_wmax.argtypes = [f32p, C.c_int64]
    L.amclj_oracle_wmax.restype = C.c_float
    L.amclj_oracle_fixed_weights.argtypes = [f32p, C.c_int64, C.c_float, C.c_void_p]
    L.amclj_oracle_fixed_weights.restype = C.c_uint64
    L.amclj_oracle_systematic_offset.argtypes = [C.c_double, C.c_uint64]
    L.amclj_oracle_systematic_offset.restype = C.c_uint64
    L.amclj_oracle_systematic.argtypes = [u64p, C.c_int64, C.c_uint64, C.c_uint64, C.c_uint64,
                                          C.c_int64, C.c_int64, C.c_int64, i64p]
    L.amclj_oracle_plan.argtypes = [u64p, C.c_int32, C.c_int64, C.c_double,
                                    C.POINTER(C.c_uint64), u64p, i64p, i64p]
    L.amclj_oracle_roulette_stream.argtypes = [u64p, C.c_int64, C.c_uint64, i32p, u32p,
                                               C.c_int64, i32p]
    L.amclj_oracle_roulette_stream.restype = C.c_int64
    L.amclj_oracle_roulette_attempts.argtypes = [C.c_uint64, C.c_int64, C.c_int64, C.c_int64,
                                                 i32p, u32p]
    L.amclj_oracle_roulette_philox.argtypes = [u64p, C.c_int64, C.c_uint64, C.c_uint64, i32p,
                                               C.c_int32]
    L.amclj_oracle_roulette_philox.restype = C.c_int64
    L.amclj_oracle_tournament_stream.argtypes = [f32p, C.c_int64, i32p, i32p]
    L.amclj_oracle_tournament_philox.argtypes = [f32p, C.c_int64, C.c_uint64, i32p]
    L.amclj_oracle_pose_estimate.argtypes = [f32p, f32p, f32p, C.c_int64, f64p]
    L.amclj_oracle_uniform_init.argtypes = [i8p, C.c_int32, C.c_int32, C.c_float, C.c_int64,
                                            C.c_int64, C.c_int32, C.c_uint64, f32p, f32p, f32p]
    L.amclj_oracle_threads.restype = C.c_int32
    L.amclj_oracle_median.argtypes = [f32p, C.c_int64]
    L.amclj_oracle_median.restype = C.c_double
    L.amclj_oracle_augmented.argtypes = [f32p, f32p, f32p, f32p, C.c_int64, C.POINTER(C.c_double),
                                         C.POINTER(C.c_double), C.c_double, C.c_double, i8p, C.c_int32,
                                         C.c_int32, C.c_float, C.c_uint64, f32p, f32p, f32p, f32p, i32p]
    L.amclj_oracle_expf_contract.argtypes = [C.c_float]
    L.amclj_oracle_expf_contract.restype = C.c_float
    L.amclj_oracle_linear_weights.argtypes = [f32p, C.c_int64, C.c_float, f32p]
    L.amclj_oracle_max.argtypes = [f32p, C.c_int64]
    L.amclj_oracle_max.restype = C.c_float
    L.amclj_oracle_gaussian_init.argtypes = [C.c_float] * 5 + [C.c_int64, C.c_int64, C.c_uint64,
                                                               f32p, f32p, f32p]
    _lib = L
    return L


# ---------------------------------------------------------------------------- helpers
def _grid(grid):
    """(cells int8 [H,W] contiguous, W, H, res float32) from an OccupancyGrid-like."""
    cells = np.ascontiguousarray(grid.data, dtype=np.int8)
    return cells, int(grid.info.width), int(grid.info.height), float(np.float32(grid.info.resolution))


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def map_range_f64(p, grid, ox, oy, theta, obs_bearing, scan_range_max) -> Ray:
    cells, W, H, res = _grid(grid)
    r = Ray()
    lib().amclj_oracle_map_range_f64(C.byref(p), cells, W, H, res, ox, oy, theta, obs_bearing,
                                     scan_range_max, C.byref(r))
    return r


def sincos_f64(a):
    s, c = C.c_double(), C.c_double()
    lib().amclj_oracle_sincos_f64(float(a), C.byref(s), C.byref(c))
    return s.value, c.value


def sincosf(a):
    s, c = C.c_float(), C.c_float()
    lib().amclj_oracle_sincosf(float(np.float32(a)), C.byref(s), C.byref(c))
    return np.float32(s.value), np.float32(c.value)


def beam_table(n, angle_min, angle_inc, max_beams):
    nb = lib().amclj_oracle_beams_used(n, max_beams)
    cb, sb = np.zeros(nb, np.float32), np.zeros(nb, np.float32)
    bi = np.zeros(nb, np.int32)
    lib().amclj_oracle_beam_table(n, angle_min, angle_inc, max_beams, cb, sb, bi)
    return cb, sb, bi


def sensor_f64(p, grid, ranges, angle_min, angle_inc, range_max, x, y, theta, threads=0):
    cells, W, H, res = _grid(grid)
    ranges = np.ascontiguousarray(ranges, np.float32)
    x, y, theta = (np.ascontiguousarray(a, np.float64) for a in (x, y, theta))
    raw = np.zeros(len(x), np.float64)
    st = SensorStats()
    lib().amclj_oracle_sensor_f64(C.byref(p), cells, W, H, res, ranges, len(ranges), angle_min,
                                  angle_inc, range_max, x, y, theta, len(x), raw, C.byref(st),
                                  threads)
    return raw, st


def rays_f64(p, grid, n_scan, angle_min, angle_inc, range_max, x, y, theta, threads=0):
    """Per-ray view of the double restatement: (steps, hit, range_m), each [len(x)][beams_used]."""
    cells, W, H, res = _grid(grid)
    x, y, theta = (np.ascontiguousarray(a, np.float64) for a in (x, y, theta))
    nb = lib().amclj_oracle_beams_used(int(n_scan), p.s6_max_beams)
    steps = np.zeros((len(x), nb), np.int32)
    hit = np.zeros((len(x), nb), np.uint8)
    rng = np.zeros((len(x), nb), np.float64)
    lib().amclj_oracle_rays_f64(C.byref(p), cells, W, H, res, int(n_scan), angle_min, angle_inc, range_max,
                                x, y, theta, len(x), steps, hit, rng, threads)
    return steps, hit, rng


def sensor_f32(p, grid, ranges, angle_min, angle_inc, range_max, x, y, theta, debug=False,
               threads=0, want_raw=True):
    cells, W, H, res = _grid(grid)
    ranges = np.ascontiguousarray(ranges, np.float32)
    x, y, theta = (np.ascontiguousarray(a, np.float32) for a in (x, y, theta))
    n = len(x)
    nb = lib().amclj_oracle_beams_used(len(ranges), p.s6_max_beams)
    raw = np.zeros(n, np.float32) if want_raw else None
    out = {}
    if debug:
        out = dict(hit_x=np.zeros((n, nb), np.int32), hit_y=np.zeros((n, nb), np.int32),
                   steps=np.zeros((n, nb), np.int32), hit=np.zeros((n, nb), np.uint8),
                   range_m=np.zeros((n, nb), np.float32))
    st = SensorStats()
    lib().amclj_oracle_sensor_f32(C.byref(p), cells, W, H, res, ranges, len(ranges), angle_min,
                                  angle_inc, range_max, x, y, theta, n, _ptr(raw),
                                  _ptr(out.get("hit_x")), _ptr(out.get("hit_y")),
                                  _ptr(out.get("steps")), _ptr(out.get("hit")),
                                  _ptr(out.get("range_m")), C.byref(st), threads)
    return raw, st, out


def control(p, curr, prev) -> Control:
    c = Control()
    lib().amclj_oracle_control(C.byref(p), np.asarray(curr, np.float64),
                               np.asarray(prev, np.float64), C.byref(c))
    return c


def motion_f64(p, curr, prev, normals, x, y, theta):
    x, y, theta = (np.array(a, np.float64) for a in (x, y, theta))
    lib().amclj_oracle_motion_f64(C.byref(p), np.asarray(curr, np.float64),
                                  np.asarray(prev, np.float64),
                                  np.ascontiguousarray(normals, np.float32), x, y, theta, len(x))
    return x, y, theta


def motion_f32(p, curr, prev, normals, x, y, theta):
    x, y, theta = (np.array(a, np.float32) for a in (x, y, theta))
    lib().amclj_oracle_motion_f32(C.byref(p), np.asarray(curr, np.float64),
                                  np.asarray(prev, np.float64),
                                  np.ascontiguousarray(normals, np.float32), x, y, theta, len(x))
    return x, y, theta


def philox(ctr, key):
    out = np.zeros(4, np.uint32)
    lib().amclj_oracle_philox(np.asarray(ctr, np.uint32), np.asarray(key, np.uint32), out)
    return out


def motion_normals(seed, gid_first, n):
    out = np.zeros((n, 3), np.float32)
    lib().amclj_oracle_motion_normals(seed, gid_first, n, out)
    return out


def wmax(w):
    return np.float32(lib().amclj_oracle_wmax(np.ascontiguousarray(w, np.float32), len(w)))


def fixed_weights(w, wmax_):
    w = np.ascontiguousarray(w, np.float32)
    q = np.zeros(len(w), np.uint64)
    total = lib().amclj_oracle_fixed_weights(w, len(w), float(wmax_), _ptr(q))
    return q, int(total)


def systematic_offset(u0, total):
    return int(lib().amclj_oracle_systematic_offset(float(u0), int(total)))


def systematic(q, total, cdf_offset, r, n_global, m_lo, count):
    q = np.ascontiguousarray(q, np.uint64)
    idx = np.zeros(count, np.int64)
    lib().amclj_oracle_systematic(q, len(q), int(total), int(cdf_offset), int(r), int(n_global),
                                  int(m_lo), int(count), idx)
    return idx


def systematic_resample(w, u0):
    """Single-rank convenience: linear float32 weights -> int64 index list."""
    wm = wmax(w)
    q, total = fixed_weights(w, wm)
    r = systematic_offset(u0, total)
    return systematic(q, total, 0, r, len(w), 0, len(w))


def plan(totals, n_global, u0):
    totals = np.ascontiguousarray(totals, np.uint64)
    g = len(totals)
    r = C.c_uint64()
    off = np.zeros(g, np.uint64)
    m_lo, m_cnt = np.zeros(g, np.int64), np.zeros(g, np.int64)
    lib().amclj_oracle_plan(totals, g, n_global, u0, C.byref(r), off, m_lo, m_cnt)
    return int(r.value), off, m_lo, m_cnt


def roulette_stream(q, total, sidx, su):
    q = np.ascontiguousarray(q, np.uint64)
    idx = np.full(len(q), -1, np.int32)
    used = lib().amclj_oracle_roulette_stream(q, len(q), int(total),
                                              np.ascontiguousarray(sidx, np.int32),
                                              np.ascontiguousarray(su, np.uint32), len(sidx), idx)
    return idx, int(used)


def roulette_attempts(seed, t0, length, n):
    sidx, su = np.zeros(length, np.int32), np.zeros(length, np.uint32)
    lib().amclj_oracle_roulette_attempts(seed, t0, length, n, sidx, su)
    return sidx, su


def roulette_philox(q, total, seed, threads=0):
    q = np.ascontiguousarray(q, np.uint64)
    idx = np.full(len(q), -1, np.int32)
    used = lib().amclj_oracle_roulette_philox(q, len(q), int(total), seed, idx, threads)
    return idx, int(used)


def tournament_stream(w, sidx):
    w = np.ascontiguousarray(w, np.float32)
    idx = np.zeros(len(w), np.int32)
    lib().amclj_oracle_tournament_stream(w, len(w), np.ascontiguousarray(sidx, np.int32), idx)
    return idx


def tournament_philox(w, seed):
    w = np.ascontiguousarray(w, np.float32)
    idx = np.zeros(len(w), np.int32)
    lib().amclj_oracle_tournament_philox(w, len(w), seed, idx)
    return idx


def pose_estimate(x, y, theta):
    out = np.zeros(4, np.float64)
    lib().amclj_oracle_pose_estimate(np.ascontiguousarray(x, np.float32),
                                     np.ascontiguousarray(y, np.float32),
                                     np.ascontiguousarray(theta, np.float32), len(x), out)
    return out


def uniform_init(grid, gid_first, n, heading_mode, seed):
    cells, W, H, res = _grid(grid)
    x, y, t = (np.zeros(n, np.float32) for _ in range(3))
    lib().amclj_oracle_uniform_init(cells, W, H, res, gid_first, n, heading_mode, seed, x, y, t)
    return x, y, t


def expf_contract(x):
    return np.float32(lib().amclj_oracle_expf_contract(float(np.float32(x))))


def linear_weights(raw, log_weights):
    """The normaliser's input: raw itself (sum p^3) or exp(raw - max) under S7.  Returns
    (w float32, wmax float32)."""
    raw = np.ascontiguousarray(raw, np.float32)
    m = np.float32(lib().amclj_oracle_max(raw, len(raw)))
    if not log_weights:
        return raw, m
    w = np.zeros(len(raw), np.float32)
    lib().amclj_oracle_linear_weights(raw, len(raw), float(m), w)
    return w, np.float32(1.0)


def gaussian_init(mx, my, mth, sd_xy, sd_th, gid_first, n, seed):
    x, y, t = (np.zeros(n, np.float32) for _ in range(3))
    lib().amclj_oracle_gaussian_init(mx, my, mth, sd_xy, sd_th, gid_first, n, seed, x, y, t)
    return x, y, t


def median(w):
    w = np.ascontiguousarray(w, np.float32)
    return float(lib().amclj_oracle_median(w, len(w)))


def augmented(w, x, y, th, grid, slow, fast, alpha_slow, alpha_fast, seed):
    """pf.clj:246-258.  Returns (x, y, theta, w, idx, slow', fast')."""
    cells, W, H, res = _grid(grid)
    w, x, y, th = (np.ascontiguousarray(a, np.float32) for a in (w, x, y, th))
    n = len(w)
    ox, oy, ot, ow = (np.zeros(n, np.float32) for _ in range(4))
    oi = np.zeros(n, np.int32)
    s, f = C.c_double(slow), C.c_double(fast)
    lib().amclj_oracle_augmented(w, x, y, th, n, C.byref(s), C.byref(f), alpha_slow, alpha_fast, cells, W, H,
                                 res, seed, ox, oy, ot, ow, oi)
    return ox, oy, ot, ow, oi, s.value, f.value


def to_heading(qx, qy, qz, qw):
    return lib().amclj_oracle_to_heading(qx, qy, qz, qw)


def quat_rotate(q, heading):
    out = np.zeros(4, np.float64)
    lib().amclj_oracle_quat_rotate(np.asarray(q, np.float64), heading, out)
    return out


def threads():
    return int(lib().amclj_oracle_threads())
