"""ctypes binding of libamclj_b200.so — the C ABI declared in include/amclj_b200.h.

This is the product's only native entry: there is NO CPU fallback.  If the library has not
been built (``python -c 'import __graft_entry__ as g; g.build()'`` or
``make -C amclj_b200/csrc``) loading fails loudly, and ``Context()`` raises when no CUDA
device is usable (AMCLJ_ERR_NODEVICE).
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
import os as _os

# AMCLJ_LIB: load a differently-built copy of the library (kernel tuning experiments only)
LIB_PATH = Path(_os.environ["AMCLJ_LIB"]) if _os.environ.get("AMCLJ_LIB") else _HERE / "libamclj_b200.so"

OK, ERR_INVALID, ERR_CUDA, ERR_STATE, ERR_NOMEM, ERR_NODEVICE = 0, -1, -2, -3, -4, -5
PRESET_REF, PRESET_NS = 0, 1
RESAMPLE_ROULETTE, RESAMPLE_SYSTEMATIC, RESAMPLE_TOURNAMENT = 0, 1, 2
BUF_X, BUF_Y, BUF_THETA, BUF_W, BUF_RAW, BUF_RAW_MAX, BUF_TOTAL, BUF_STAGE_OUT, BUF_STAGE_IN, BUF_CDF, BUF_TOTALS_ALL = range(11)
IPC_BYTES = 8 * 64

# every symbol include/amclj_b200.h declares (tests check the library exports all of them)
SYMBOLS = [
    "amclj_abi_version", "amclj_params_default", "amclj_create", "amclj_destroy",
    "amclj_last_error", "amclj_set_stream", "amclj_synchronize", "amclj_set_params",
    "amclj_get_params", "amclj_set_shard", "amclj_set_map", "amclj_set_particles",
    "amclj_get_particles", "amclj_num_particles", "amclj_set_poses_quat", "amclj_get_poses_quat", "amclj_init_uniform", "amclj_init_gaussian",
    "amclj_motion_update", "amclj_sensor_update", "amclj_get_raw_weights", "amclj_raycast_debug",
    "amclj_beams_used", "amclj_ray_table_entries", "amclj_ray_ring", "amclj_normalize", "amclj_resample", "amclj_get_resample_indices",
    "amclj_filter_update", "amclj_stage_scan", "amclj_filter_update_async",
    "amclj_filter_update_host", "amclj_pose_estimate", "amclj_get_stats", "amclj_microbench", "amclj_median_weight",
    "amclj_augmented_resample", "amclj_augment_state",
    "amclj_shard_motion_sensor_async", "amclj_shard_cdf_async", "amclj_shard_generate_async",
    "amclj_shard_get_stage_idx", "amclj_shard_adopt_async", "amclj_plan_systematic",
    "amclj_device_ptr", "amclj_reserve_stage", "amclj_shard_export", "amclj_shard_import",
    "amclj_shard_attach_local", "amclj_shard_pull_async", "amclj_shard_update_async", "amclj_shard_update_host", "amclj_shard_rendezvous_async", "amclj_filter_update_scan_async", "amclj_shard_update_scan_async", "amclj_host_register", "amclj_host_unregister",
    "amclj_create_multi", "amclj_destroy_multi", "amclj_multi_size", "amclj_multi_ctx", "amclj_mlast_error",
    "amclj_mset_params", "amclj_mset_map", "amclj_mset_particles", "amclj_mget_particles", "amclj_mnum_particles",
    "amclj_minit_uniform", "amclj_mstage_scan", "amclj_mfilter_update_async", "amclj_mfilter_update_host",
    "amclj_msynchronize", "amclj_mpose_estimate",
]


class Params(C.Structure):
    """``amclj_params``: the ten semantic switches and the Clojure ``def`` constants."""

    _fields_ = [
        ("s1_use_heading", C.c_int32), ("s2_use_scan_range_max", C.c_int32),
        ("s3_proper_endpoint", C.c_int32), ("s4_endpoint_termination", C.c_int32),
        ("s5_use_sigma_hit", C.c_int32), ("s6_max_beams", C.c_int32),
        ("s7_log_weights", C.c_int32), ("s8_resampler", C.c_int32),
        ("s9_pr_rot1_variance", C.c_int32), ("s10_apply_laser_offset", C.c_int32),
        ("collect_stats", C.c_int32), ("df_mode", C.c_int32),
        ("max_range", C.c_float), ("z_hit", C.c_float), ("z_short", C.c_float),
        ("z_max", C.c_float), ("z_rand", C.c_float), ("sigma_hit", C.c_float),
        ("lambda_short", C.c_float), ("laser_x", C.c_float), ("laser_y", C.c_float),
        ("alpha1", C.c_float), ("alpha2", C.c_float), ("alpha3", C.c_float),
        ("alpha4", C.c_float), ("reserved1", C.c_float * 3),
    ]


class Stats(C.Structure):
    _fields_ = [
        ("ms_motion", C.c_double), ("ms_sensor", C.c_double), ("ms_resample", C.c_double),
        ("ms_total", C.c_double), ("rays", C.c_uint64), ("cells_visited", C.c_uint64),
        ("probes", C.c_uint64), ("hits", C.c_uint64), ("n_particles", C.c_int64),
        ("beams_used", C.c_int32), ("map_in_smem", C.c_int32), ("oob_probes", C.c_uint64),
        ("table_cells", C.c_uint64), ("table_particles", C.c_uint64), ("table_misses", C.c_uint64),
        ("ms_sensor_main", C.c_double), ("sensor_path", C.c_int32), ("kernels_launched", C.c_int32),
        ("zero_weight_updates", C.c_uint64), ("ms_table_fill", C.c_double), ("table_bytes", C.c_uint64),
    ]


class AmcljError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"amclj_b200 error {code}: {msg}")
        self.code = code


def build(force: bool = False) -> Path:
    """Compile the library in-tree for sm_100a (amclj_b200/csrc/Makefile)."""
    args = ["make", "-C", str(_HERE / "csrc"), "-j4"] + (["-B"] if force else [])
    subprocess.run(args, check=True, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    return LIB_PATH


_lib = None


def _vp(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def lib():
    """Load the shared library (never a fallback: a missing .so is an error)."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise ImportError(f"{LIB_PATH} is missing: build it with `make -C amclj_b200/csrc` "
                          "(there is no CPU fallback)")
    L = C.CDLL(str(LIB_PATH))
    vp, i32, i64, u64, f32, f64 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64, C.c_float, C.c_double
    PP, d3 = C.POINTER(Params), C.POINTER(C.c_double)
    sig = {
        "amclj_abi_version": ([], i32),
        "amclj_params_default": ([PP, i32], i32),
        "amclj_create": ([i32, C.POINTER(vp)], i32),
        "amclj_destroy": ([vp], i32),
        "amclj_last_error": ([vp], C.c_char_p),
        "amclj_set_stream": ([vp, vp], i32),
        "amclj_synchronize": ([vp], i32),
        "amclj_set_params": ([vp, PP], i32),
        "amclj_get_params": ([vp, PP], i32),
        "amclj_set_shard": ([vp, i32, i32, i64, i64], i32),
        "amclj_set_map": ([vp, vp, i32, i32, f32], i32),
        "amclj_set_particles": ([vp, vp, vp, vp, vp, i64], i32),
        "amclj_get_particles": ([vp, vp, vp, vp, vp], i32),
        "amclj_num_particles": ([vp], i64),
        "amclj_set_poses_quat": ([vp, vp, vp, i64], i32),
        "amclj_get_poses_quat": ([vp, vp, vp], i32),
        "amclj_init_uniform": ([vp, i64, i32, u64], i32),
        "amclj_init_gaussian": ([vp, i64, f64, f64, f64, f64, f64, u64], i32),
        "amclj_motion_update": ([vp, d3, d3, vp, u64], i32),
        "amclj_sensor_update": ([vp, vp, i32, f32, f32, f32], i32),
        "amclj_get_raw_weights": ([vp, vp], i32),
        "amclj_raycast_debug": ([vp, i32, f32, f32, f32, i64, i64, vp, vp, vp, vp, vp], i32),
        "amclj_beams_used": ([vp, i32], i32),
        "amclj_ray_table_entries": ([vp], i32),
        "amclj_ray_ring": ([f64, i32, vp, vp, C.POINTER(i32)], i32),
        "amclj_normalize": ([vp, i32, C.POINTER(u64), C.POINTER(f32)], i32),
        "amclj_resample": ([vp, i32, f64, vp, vp, i64, u64, vp], i32),
        "amclj_get_resample_indices": ([vp, vp], i32),
        "amclj_filter_update": ([vp, d3, d3, vp, u64, vp, i32, f32, f32, f32, f64], i32),
        "amclj_stage_scan": ([vp, vp, i32, f32, f32, f32], i32),
        "amclj_filter_update_async": ([vp, d3, d3, u64, f64], i32),
        "amclj_filter_update_host": ([vp, vp, vp, vp, vp, i64, d3, d3, vp, u64, vp, i32, f32, f32,
                                      f32, f64], i32),
        "amclj_pose_estimate": ([vp, d3], i32),
        "amclj_get_stats": ([vp, C.POINTER(Stats)], i32),
        "amclj_microbench": ([vp, i32, C.POINTER(f64)], i32),
        "amclj_median_weight": ([vp, C.POINTER(f64)], i32),
        "amclj_augmented_resample": ([vp, f64, f64, u64, vp], i32),
        "amclj_augment_state": ([vp, d3, d3], i32),
        "amclj_shard_motion_sensor_async": ([vp, d3, d3, u64], i32),
        "amclj_shard_cdf_async": ([vp], i32),
        "amclj_shard_generate_async": ([vp, u64, u64, u64, i64, i64], i32),
        "amclj_shard_get_stage_idx": ([vp, vp, i64], i32),
        "amclj_shard_adopt_async": ([vp, i64], i32),
        "amclj_plan_systematic": ([vp, i32, i64, f64, C.POINTER(u64), vp, vp, vp, vp], i32),
        "amclj_device_ptr": ([vp, i32, C.POINTER(vp), C.POINTER(i64)], i32),
        "amclj_reserve_stage": ([vp, i64], i32),
        "amclj_shard_export": ([vp, vp], i32),
        "amclj_shard_import": ([vp, i32, vp, i64, i64], i32),
        "amclj_shard_attach_local": ([vp, i32, vp], i32),
        "amclj_shard_pull_async": ([vp, f64], i32),
        "amclj_shard_update_async": ([vp, d3, d3, u64, f64], i32),
        "amclj_shard_rendezvous_async": ([vp], i32),
        "amclj_filter_update_scan_async": ([vp, d3, d3, u64, f64, vp, i32, f32, f32, f32], i32),
        "amclj_shard_update_scan_async": ([vp, d3, d3, u64, f64, vp, i32, f32, f32, f32], i32),
        "amclj_host_register": ([vp, vp, i64], i32),
        "amclj_host_unregister": ([vp, vp], i32),
        "amclj_shard_update_host": ([vp, vp, vp, vp, vp, i64, d3, d3, u64, vp, i32, f32, f32, f32, f64], i32),
        "amclj_create_multi": ([vp, i32, C.POINTER(vp)], i32),
        "amclj_destroy_multi": ([vp], i32),
        "amclj_multi_size": ([vp], i32),
        "amclj_multi_ctx": ([vp, i32], vp),
        "amclj_mlast_error": ([vp], C.c_char_p),
        "amclj_mset_params": ([vp, PP], i32),
        "amclj_mset_map": ([vp, vp, i32, i32, f32], i32),
        "amclj_mset_particles": ([vp, vp, vp, vp, vp, i64], i32),
        "amclj_mget_particles": ([vp, vp, vp, vp, vp], i32),
        "amclj_mnum_particles": ([vp], i64),
        "amclj_minit_uniform": ([vp, i64, i32, u64], i32),
        "amclj_mstage_scan": ([vp, vp, i32, f32, f32, f32], i32),
        "amclj_mfilter_update_async": ([vp, d3, d3, u64, f64], i32),
        "amclj_mfilter_update_host": ([vp, vp, vp, vp, vp, i64, d3, d3, u64, vp, i32, f32, f32, f32, f64], i32),
        "amclj_msynchronize": ([vp], i32),
        "amclj_mpose_estimate": ([vp, d3], i32),
    }
    for name, (args, res) in sig.items():
        fn = getattr(L, name)
        fn.argtypes = args
        fn.restype = res
    assert set(sig) == set(SYMBOLS)
    _lib = L
    return L


def params(preset: str = "REF", **overrides) -> Params:
    p = Params()
    rc = lib().amclj_params_default(C.byref(p), PRESET_NS if preset.upper() == "NS" else PRESET_REF)
    if rc:
        raise AmcljError(rc, "params_default")
    for k, v in overrides.items():
        if not hasattr(p, k):
            raise AttributeError(k)
        setattr(p, k, v)
    return p


def _d3(v):
    return (C.c_double * 3)(*[float(t) for t in v])


def _f32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float32)


def ray_ring(r_cells: float):
    """Host-only view of the end-cell ring of a ray of r_cells cells: (E, lo[2E+1], cnt[2E+1], entries)."""
    L = lib()
    cap = 2 * int(r_cells + 4) + 8
    lo, cnt = np.zeros(cap, np.int32), np.zeros(cap, np.int32)
    e = C.c_int32(0)
    n = L.amclj_ray_ring(float(r_cells), cap, _vp(lo), _vp(cnt), C.byref(e))
    rows = 2 * e.value + 1
    return e.value, lo[:rows], cnt[:rows], int(n)


def plan_systematic(totals, n_global, u0):
    """Host-only planner for the sharded systematic resampler (no GPU needed)."""
    totals = np.ascontiguousarray(totals, np.uint64)
    g = len(totals)
    r = C.c_uint64()
    off = np.zeros(g, np.uint64)
    m_lo, m_cnt = np.zeros(g, np.int64), np.zeros(g, np.int64)
    send = np.zeros((g, g), np.int64)
    rc = lib().amclj_plan_systematic(_vp(totals), g, int(n_global), float(u0), C.byref(r), _vp(off),
                                     _vp(m_lo), _vp(m_cnt), _vp(send))
    if rc:
        raise AmcljError(rc, "plan_systematic: invalid arguments")
    return int(r.value), off, m_lo, m_cnt, send


class Context:
    """One ``amclj_ctx``: one GPU, one shard of the particle set."""

    def __init__(self, device: int = 0, preset: str | None = None):
        self._L = lib()
        h = C.c_void_p()
        rc = self._L.amclj_create(int(device), C.byref(h))
        if rc:
            raise AmcljError(rc, "amclj_create failed (no usable CUDA device? there is no CPU fallback)")
        self._h = h
        self.device = device
        if preset is not None:
            self.set_params(params(preset))

    # -- plumbing ----------------------------------------------------------------------
    def _ck(self, rc):
        if rc:
            raise AmcljError(rc, self._L.amclj_last_error(self._h).decode())

    def close(self):
        if getattr(self, "_h", None):
            self._L.amclj_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def synchronize(self):
        self._ck(self._L.amclj_synchronize(self._h))

    def set_stream(self, cuda_stream: int | None):
        self._ck(self._L.amclj_set_stream(self._h, C.c_void_p(cuda_stream or 0)))

    def set_params(self, p: Params):
        self._ck(self._L.amclj_set_params(self._h, C.byref(p)))

    def get_params(self) -> Params:
        p = Params()
        self._ck(self._L.amclj_get_params(self._h, C.byref(p)))
        return p

    def set_shard(self, rank, world, global_offset, n_global):
        self._ck(self._L.amclj_set_shard(self._h, rank, world, int(global_offset), int(n_global)))

    # -- data --------------------------------------------------------------------------
    def set_map(self, grid):
        """``grid``: an OccupancyGrid-like (``.data`` int8 [H,W], ``.info``)."""
        cells = np.ascontiguousarray(grid.data, dtype=np.int8)
        self._ck(self._L.amclj_set_map(self._h, _vp(cells), int(grid.info.width), int(grid.info.height),
                                       float(np.float32(grid.info.resolution))))

    def set_particles(self, x, y, theta, w=None):
        x, y, theta, w = _f32(x), _f32(y), _f32(theta), _f32(w)
        self._ck(self._L.amclj_set_particles(self._h, _vp(x), _vp(y), _vp(theta), _vp(w), len(x)))

    @property
    def n(self) -> int:
        return int(self._L.amclj_num_particles(self._h))

    def get_particles(self, out=None):
        n = self.n
        x, y, t, w = out if out is not None else (np.empty(n, np.float32) for _ in range(4))
        self._ck(self._L.amclj_get_particles(self._h, _vp(x), _vp(y), _vp(t), _vp(w)))
        return x, y, t, w

    def set_poses_quat(self, poses7, w=None):
        """PoseArray rows (x, y, z, qx, qy, qz, qw) in double; theta = quat/to-heading inside the library."""
        poses7 = np.ascontiguousarray(poses7, np.float64).reshape(-1, 7)
        w = _f32(w)
        self._ck(self._L.amclj_set_poses_quat(self._h, _vp(poses7), _vp(w), len(poses7)))

    def get_poses_quat(self):
        poses7, w = np.empty((self.n, 7), np.float64), np.empty(self.n, np.float32)
        self._ck(self._L.amclj_get_poses_quat(self._h, _vp(poses7), _vp(w)))
        return poses7, w

    def init_uniform(self, n, heading_mode=1, seed=1):
        self._ck(self._L.amclj_init_uniform(self._h, int(n), int(heading_mode), int(seed)))

    def init_gaussian(self, n, x, y, theta, sd_xy=0.1, sd_theta=0.0, seed=1):
        self._ck(self._L.amclj_init_gaussian(self._h, int(n), x, y, theta, sd_xy, sd_theta, int(seed)))

    # -- stages ------------------------------------------------------------------------
    def motion_update(self, curr, prev, normals=None, seed=0):
        normals = _f32(normals)
        self._ck(self._L.amclj_motion_update(self._h, _d3(curr), _d3(prev), _vp(normals), int(seed)))

    def sensor_update(self, ranges, angle_min, angle_inc, range_max):
        ranges = _f32(ranges)
        self._ck(self._L.amclj_sensor_update(self._h, _vp(ranges), len(ranges), angle_min, angle_inc,
                                             range_max))

    def stage_scan(self, ranges, angle_min, angle_inc, range_max):
        ranges = _f32(ranges)
        self._ck(self._L.amclj_stage_scan(self._h, _vp(ranges), len(ranges), angle_min, angle_inc,
                                          range_max))

    def raw_weights(self):
        raw = np.empty(self.n, np.float32)
        self._ck(self._L.amclj_get_raw_weights(self._h, _vp(raw)))
        return raw

    def ray_table_entries(self):
        return int(self._L.amclj_ray_table_entries(self._h))

    def beams_used(self, n_scan):
        return int(self._L.amclj_beams_used(self._h, int(n_scan)))

    def raycast_debug(self, n_scan, angle_min, angle_inc, range_max, first=0, count=None):
        count = self.n - first if count is None else count
        nb = self.beams_used(n_scan)
        out = dict(hit_x=np.zeros((count, nb), np.int32), hit_y=np.zeros((count, nb), np.int32),
                   steps=np.zeros((count, nb), np.int32), hit=np.zeros((count, nb), np.uint8),
                   range_m=np.zeros((count, nb), np.float32))
        self._ck(self._L.amclj_raycast_debug(self._h, int(n_scan), angle_min, angle_inc, range_max,
                                             int(first), int(count), _vp(out["hit_x"]),
                                             _vp(out["hit_y"]), _vp(out["steps"]), _vp(out["hit"]),
                                             _vp(out["range_m"])))
        return out

    def normalize(self, from_raw=True):
        total, wmax = C.c_uint64(), C.c_float()
        self._ck(self._L.amclj_normalize(self._h, int(bool(from_raw)), C.byref(total), C.byref(wmax)))
        return int(total.value), float(wmax.value)

    def cdf(self):
        """Parity probe: the inclusive fixed-point CDF of the last normalisation."""
        import torch  # device memory plumbing only
        ptr, nbytes = self.device_ptr(BUF_CDF)
        n = self.n
        t = torch.empty(n, dtype=torch.int64, device=f"cuda:{self.device}")
        src = _DevArray(ptr, (n,), "<i8")
        t.copy_(torch.as_tensor(src, device=f"cuda:{self.device}"))
        return t.cpu().numpy().view(np.uint64)

    def resample(self, mode=RESAMPLE_SYSTEMATIC, u0=0.0, stream_idx=None, stream_u=None, seed=0,
                 want_idx=True):
        n = self.n
        idx = np.empty(n, np.int32) if want_idx else None
        si = None if stream_idx is None else np.ascontiguousarray(stream_idx, np.int32)
        su = None if stream_u is None else np.ascontiguousarray(stream_u, np.uint32)
        self._ck(self._L.amclj_resample(self._h, int(mode), float(u0), _vp(si), _vp(su),
                                        0 if si is None else len(si), int(seed), _vp(idx)))
        return idx

    def resample_indices(self):
        idx = np.empty(self.n, np.int32)
        self._ck(self._L.amclj_get_resample_indices(self._h, _vp(idx)))
        return idx

    def filter_update(self, curr, prev, ranges, angle_min, angle_inc, range_max, u0, normals=None,
                      seed=0):
        ranges, normals = _f32(ranges), _f32(normals)
        self._ck(self._L.amclj_filter_update(self._h, _d3(curr), _d3(prev), _vp(normals), int(seed),
                                             _vp(ranges), len(ranges), angle_min, angle_inc,
                                             range_max, float(u0)))

    def filter_update_async(self, curr, prev, seed, u0):
        self._ck(self._L.amclj_filter_update_async(self._h, _d3(curr), _d3(prev), int(seed), float(u0)))

    def filter_update_host(self, x, y, theta, w, curr, prev, ranges, angle_min, angle_inc, range_max,
                           u0, normals=None, seed=0):
        """Host PoseArray in, resampled PoseArray out (in place); buffers must be float32."""
        ranges, normals = _f32(ranges), _f32(normals)
        for a in (x, y, theta, w):
            assert a.dtype == np.float32 and a.flags.c_contiguous
        self._ck(self._L.amclj_filter_update_host(
            self._h, _vp(x), _vp(y), _vp(theta), _vp(w), len(x), _d3(curr), _d3(prev), _vp(normals),
            int(seed), _vp(ranges), len(ranges), angle_min, angle_inc, range_max, float(u0)))

    def pose_estimate(self):
        out = (C.c_double * 4)()
        self._ck(self._L.amclj_pose_estimate(self._h, C.cast(out, C.POINTER(C.c_double))))
        return np.array(list(out))

    def stats(self) -> Stats:
        s = Stats()
        self._ck(self._L.amclj_get_stats(self._h, C.byref(s)))
        return s

    def median_weight(self):
        v = C.c_double()
        self._ck(self._L.amclj_median_weight(self._h, C.byref(v)))
        return float(v.value)

    def augmented_resample(self, alpha_slow=0.01, alpha_fast=0.1, seed=0):
        idx = np.empty(self.n, np.int32)
        self._ck(self._L.amclj_augmented_resample(self._h, alpha_slow, alpha_fast, int(seed), _vp(idx)))
        return idx

    def augment_state(self, set_to=None):
        out = (C.c_double * 2)()
        arg = None if set_to is None else (C.c_double * 2)(*set_to)
        self._ck(self._L.amclj_augment_state(self._h, arg, out))
        return float(out[0]), float(out[1])

    def microbench(self, which=0):
        """Dependent-lookup ceiling (lookups/s): 0 = shared-memory nibble table, 1 = L2 byte table."""
        v = C.c_double()
        self._ck(self._L.amclj_microbench(self._h, int(which), C.byref(v)))
        return float(v.value)

    # -- sharded building blocks ---------------------------------------------------------
    def shard_motion_sensor_async(self, curr, prev, seed):
        self._ck(self._L.amclj_shard_motion_sensor_async(self._h, _d3(curr), _d3(prev), int(seed)))

    def shard_cdf_async(self):
        self._ck(self._L.amclj_shard_cdf_async(self._h))

    def shard_generate_async(self, total_global, cdf_offset, r, m_lo, count):
        self._ck(self._L.amclj_shard_generate_async(self._h, int(total_global), int(cdf_offset), int(r),
                                                    int(m_lo), int(count)))

    def shard_stage_idx(self, count):
        idx = np.empty(count, np.int64)
        self._ck(self._L.amclj_shard_get_stage_idx(self._h, _vp(idx), int(count)))
        return idx

    def shard_adopt_async(self, count):
        self._ck(self._L.amclj_shard_adopt_async(self._h, int(count)))

    def shard_export(self) -> bytes:
        buf = (C.c_uint8 * IPC_BYTES)()
        self._ck(self._L.amclj_shard_export(self._h, buf))
        return bytes(buf)

    def shard_import(self, rank, handles: bytes, n_peer, peer_global_offset):
        buf = (C.c_uint8 * IPC_BYTES).from_buffer_copy(handles)
        self._ck(self._L.amclj_shard_import(self._h, int(rank), buf, int(n_peer), int(peer_global_offset)))

    def shard_attach_local(self, rank, peer: "Context"):
        self._ck(self._L.amclj_shard_attach_local(self._h, int(rank), peer._h))

    def shard_pull_async(self, u0):
        self._ck(self._L.amclj_shard_pull_async(self._h, float(u0)))

    def shard_update_async(self, curr, prev, seed, u0):
        """The whole sharded update in one call: exchanges through peer-memory mailboxes, no collective library."""
        self._ck(self._L.amclj_shard_update_async(self._h, _d3(curr), _d3(prev), int(seed), float(u0)))

    def filter_update_scan_async(self, curr, prev, seed, u0, ranges, angle_min, angle_inc, range_max):
        """stage_scan + filter_update_async in one call (motion first when the scan's geometry is staged)."""
        ranges = _f32(ranges)
        self._ck(self._L.amclj_filter_update_scan_async(self._h, _d3(curr), _d3(prev), int(seed), float(u0), _vp(ranges),
                                                        len(ranges), angle_min, angle_inc, range_max))

    def shard_update_scan_async(self, curr, prev, seed, u0, ranges, angle_min, angle_inc, range_max):
        ranges = _f32(ranges)
        self._ck(self._L.amclj_shard_update_scan_async(self._h, _d3(curr), _d3(prev), int(seed), float(u0), _vp(ranges),
                                                       len(ranges), angle_min, angle_inc, range_max))

    def host_register(self, arr):
        """Page-lock a caller-owned numpy buffer for the *_host calls (what the JNA shim does with its Memory)."""
        self._ck(self._L.amclj_host_register(self._h, _vp(arr), int(arr.nbytes)))

    def host_unregister(self, arr):
        self._ck(self._L.amclj_host_unregister(self._h, _vp(arr)))

    def shard_rendezvous_async(self):
        """Device-side barrier among the connected shards (one GPU each): what follows starts together."""
        self._ck(self._L.amclj_shard_rendezvous_async(self._h))

    def shard_update_host(self, x, y, theta, w, curr, prev, ranges, angle_min, angle_inc, range_max, u0, seed=0):
        ranges = _f32(ranges)
        for a in (x, y, theta, w):
            assert a.dtype == np.float32 and a.flags.c_contiguous
        self._ck(self._L.amclj_shard_update_host(self._h, _vp(x), _vp(y), _vp(theta), _vp(w), len(x), _d3(curr),
                                                 _d3(prev), int(seed), _vp(ranges), len(ranges), angle_min, angle_inc,
                                                 range_max, float(u0)))

    def reserve_stage(self, rows):
        self._ck(self._L.amclj_reserve_stage(self._h, int(rows)))

    def device_ptr(self, which):
        p, b = C.c_void_p(), C.c_int64()
        self._ck(self._L.amclj_device_ptr(self._h, int(which), C.byref(p), C.byref(b)))
        return int(p.value or 0), int(b.value)

    def device_tensor(self, which, shape, typestr):
        """Zero-copy torch view of a library-owned device buffer (for torch.distributed)."""
        import torch
        ptr, nbytes = self.device_ptr(which)
        return torch.as_tensor(_DevArray(ptr, tuple(shape), typestr), device=f"cuda:{self.device}")


class _RankView(Context):
    """A rank's ``amclj_ctx`` inside a MultiContext (borrowed: never destroyed from here)."""

    def __init__(self, handle, device):
        self._L = lib()
        self._h = handle
        self.device = device

    def close(self):
        self._h = None


class MultiContext:
    """``amclj_mctx``: one process, one caller thread, several GPUs."""

    def __init__(self, devices, preset: str | None = None):
        self._L = lib()
        devs = (C.c_int32 * len(devices))(*[int(d) for d in devices])
        h = C.c_void_p()
        rc = self._L.amclj_create_multi(devs, len(devices), C.byref(h))
        if rc:
            raise AmcljError(rc, "amclj_create_multi failed (no usable CUDA device? there is no CPU fallback)")
        self._h = h
        self.devices = list(devices)
        if preset is not None:
            self.set_params(params(preset))

    def _ck(self, rc):
        if rc:
            raise AmcljError(rc, self._L.amclj_mlast_error(self._h).decode())

    def close(self):
        if getattr(self, "_h", None):
            self._L.amclj_destroy_multi(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def world(self):
        return int(self._L.amclj_multi_size(self._h))

    def rank(self, r) -> Context:
        return _RankView(C.c_void_p(self._L.amclj_multi_ctx(self._h, int(r))), self.devices[r])

    def set_params(self, p):
        self._ck(self._L.amclj_mset_params(self._h, C.byref(p)))

    def set_map(self, grid):
        cells = np.ascontiguousarray(grid.data, dtype=np.int8)
        self._ck(self._L.amclj_mset_map(self._h, _vp(cells), int(grid.info.width), int(grid.info.height),
                                        float(np.float32(grid.info.resolution))))

    def set_particles(self, x, y, theta, w=None):
        x, y, theta, w = _f32(x), _f32(y), _f32(theta), _f32(w)
        self._ck(self._L.amclj_mset_particles(self._h, _vp(x), _vp(y), _vp(theta), _vp(w), len(x)))

    @property
    def n(self):
        return int(self._L.amclj_mnum_particles(self._h))

    def get_particles(self):
        x, y, t, w = (np.empty(self.n, np.float32) for _ in range(4))
        self._ck(self._L.amclj_mget_particles(self._h, _vp(x), _vp(y), _vp(t), _vp(w)))
        return x, y, t, w

    def init_uniform(self, n, heading_mode=1, seed=1):
        self._ck(self._L.amclj_minit_uniform(self._h, int(n), int(heading_mode), int(seed)))

    def stage_scan(self, ranges, angle_min, angle_inc, range_max):
        ranges = _f32(ranges)
        self._ck(self._L.amclj_mstage_scan(self._h, _vp(ranges), len(ranges), angle_min, angle_inc, range_max))

    def filter_update_async(self, curr, prev, seed, u0):
        self._ck(self._L.amclj_mfilter_update_async(self._h, _d3(curr), _d3(prev), int(seed), float(u0)))

    def filter_update_host(self, x, y, theta, w, curr, prev, ranges, angle_min, angle_inc, range_max, u0, seed=0):
        ranges = _f32(ranges)
        for a in (x, y, theta, w):
            assert a.dtype == np.float32 and a.flags.c_contiguous
        self._ck(self._L.amclj_mfilter_update_host(self._h, _vp(x), _vp(y), _vp(theta), _vp(w), len(x), _d3(curr),
                                                   _d3(prev), int(seed), _vp(ranges), len(ranges), angle_min, angle_inc,
                                                   range_max, float(u0)))

    def synchronize(self):
        self._ck(self._L.amclj_msynchronize(self._h))

    def pose_estimate(self):
        out = (C.c_double * 4)()
        self._ck(self._L.amclj_mpose_estimate(self._h, C.cast(out, C.POINTER(C.c_double))))
        return np.array(list(out))


class _DevArray:
    """Minimal ``__cuda_array_interface__`` carrier for a raw device pointer."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"data": (ptr, False), "shape": shape, "typestr": typestr,
                                         "version": 2, "strides": None}
