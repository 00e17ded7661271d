"""Persistent map-wide ray tables (csrc/ptable.cu) and parity at the benchmarked sizes.

The tables hold, for every free cell of the map, the range of a ray to every end cell it can reach
(sensor.clj:130-139 rounds both ends to cells before anything else looks at them), filled once
with the walk kernel's own ray cast.  Checked here:
  * the PRODUCTION look-up code (not a diagnostic twin) leaves every range it reads in the parity
    probe's buffer: bit-exact against the oracle's cell-by-cell walk (df_mode 7);
  * raw weights bit-identical to the walk kernel (df_mode 4) in every semantic mode, so which path
    ran can never be seen in a result;
  * C3 (1M uniform) and the C4 per-GPU share (1.25M): weights of an oracle sample to 1e-5, rays of
    a 64-particle window bit-exact — at the sizes bench.py times.
"""
import math

import numpy as np
import pytest

import oracle
from amclj_b200 import _lib, maps, workloads

pytestmark = pytest.mark.gpu

RTOL = 1e-5


@pytest.fixture(scope="module")
def lg():
    return maps.lgfloor()


def _scan(lg, n=360):
    return workloads.make_scan(lg, workloads.TRUE_POSE, n, -math.pi, 2 * math.pi / n, 5.6)


def _edge_particles(lg):
    """Poses the look-up must get right besides free cells: on occupied / unknown cells, outside the
    map on every side, NaN, huge."""
    res = float(lg.info.resolution)
    W, H = lg.info.width, lg.info.height
    ex = np.array([-1.0, (W + 3) * res, 5.0, 5.0, 0.0, (W - 1) * res, (W - 0.6) * res, 16.3, np.nan, 1e12,
                   -0.024, 0.0251], np.float32)
    ey = np.array([5.0, 5.0, -2.0, (H + 10) * res, 0.0, (H - 1) * res, (H - 0.6) * res, np.nan, 3.0, -1e12,
                   -0.024, 0.0251], np.float32)
    et = np.linspace(-4, 4, len(ex)).astype(np.float32)
    occ = np.argwhere(lg.data == 100)[::97][:40]
    unk = np.argwhere(lg.data == -1)[::9973][:20]
    cells = np.concatenate([occ, unk])
    return (np.concatenate([ex, (cells[:, 1] * res).astype(np.float32)]),
            np.concatenate([ey, (cells[:, 0] * res).astype(np.float32)]),
            np.concatenate([et, np.zeros(len(cells), np.float32)]))


def _ranges_vs_oracle(c, lg, x, y, th, scan, preset, first=0, count=None, **kw):
    """df_mode 7 parity probe: the production look-up kernel writes every range it read."""
    po = oracle.params(preset, **{k: v for k, v in kw.items() if k != "df_mode"})
    c.set_params(_lib.params(preset, df_mode=7, **kw))
    c.set_particles(x, y, th)
    count = len(x) - first if count is None else count
    dbg = c.raycast_debug(len(scan.ranges), scan.angle_min, scan.angle_inc, scan.range_max, first, count)
    sl = slice(first, first + count)
    _, _, ref = oracle.sensor_f32(po, lg, scan.ranges, scan.angle_min, scan.angle_inc, scan.range_max,
                                  x[sl], y[sl], th[sl], debug=True, want_raw=False)
    bad = np.argwhere(dbg["range_m"] != ref["range_m"])
    assert len(bad) == 0, f"{preset} {kw}: {len(bad)} ranges differ, first at {bad[:5].tolist()}"
    s = c.stats()
    assert s.table_misses == 0  # no end cell ever left the ring
    return s


@pytest.mark.parametrize("preset", ["NS", "REF"])
def test_production_lookup_ranges_bit_exact(lg, preset):
    x, y, th = workloads.uniform_particles(lg, 3000, seed=21)
    ex, ey, et = _edge_particles(lg)
    x, y, th = np.concatenate([x, ex]), np.concatenate([y, ey]), np.concatenate([th, et])
    with _lib.Context(0) as c:
        c.set_params(_lib.params(preset, df_mode=7))
        c.set_map(lg)
        s = _ranges_vs_oracle(c, lg, x, y, th, _scan(lg), preset)
        assert s.table_bytes > 0 and s.ms_table_fill > 0


def test_production_lookup_ranges_c2_dense_cloud(lg):
    """C2's cloud (100k particles on ~300 cells): a 2,000-particle window of the full set."""
    w = workloads.c2()
    with _lib.Context(0) as c:
        c.set_params(_lib.params("NS", df_mode=7))
        c.set_map(lg)
        _ranges_vs_oracle(c, lg, w.x, w.y, w.theta, w.scan, "NS", first=40_000, count=2000)


@pytest.mark.parametrize("which", ["c2", "c3"])
def test_production_lookup_casts_the_rays_of_the_double_path(lg, which):
    """The ranges the production look-up kernel reads, against the DOUBLE restatement of the Clojure ray by ray
    (oracle.rays_f64) — 40,000 particles x 360 beams = 14.4 M rays per cloud: every hit range equal up to float
    rounding of hypot * res, every miss exact.  A ray cast to a different cell would show as a range off by about
    a cell (0.05 m) or as a hit turned miss; csrc/contract.h's error budget says there is none."""
    w = workloads.c2(40_000) if which == "c2" else workloads.c3(40_000)
    s = w.scan
    with _lib.Context(0) as c:
        c.set_params(_lib.params("NS", df_mode=7))
        c.set_map(lg)
        c.set_particles(w.x, w.y, w.theta)
        dbg = c.raycast_debug(len(s.ranges), s.angle_min, s.angle_inc, s.range_max)
        assert c.stats().table_misses == 0
    _, _, rng = oracle.rays_f64(oracle.params("NS"), lg, len(s.ranges), s.angle_min, s.angle_inc, s.range_max,
                                w.x, w.y, w.theta)
    diff = np.abs(dbg["range_m"].astype(np.float64) - rng)
    bad = np.argwhere(diff > 2e-6)
    assert len(bad) == 0, f"{len(bad)} of {diff.size} rays differ from the double path, first {bad[:4].tolist()}, max {diff.max()}"


@pytest.mark.parametrize("switches", [
    dict(s1_use_heading=0), dict(s3_proper_endpoint=0), dict(s4_endpoint_termination=0),
    dict(s2_use_scan_range_max=0), dict(s10_apply_laser_offset=1),
    dict(s3_proper_endpoint=0, s4_endpoint_termination=0), dict(s6_max_beams=30), dict(s7_log_weights=0),
])
def test_switch_matrix_ranges_and_weights(lg, switches):
    """Every semantic switch is a launch-uniform parameter: ranges against the oracle,
    raw weights bit-identical to the walk kernel."""
    x, y, th = workloads.uniform_particles(lg, 1500, seed=22)
    scan = workloads.make_scan(lg, workloads.TRUE_POSE, 180, -math.pi / 2, math.pi / 179, 5.6)
    with _lib.Context(0) as c, _lib.Context(0) as w:
        c.set_params(_lib.params("NS", df_mode=7, **switches))
        c.set_map(lg)
        _ranges_vs_oracle(c, lg, x, y, th, scan, "NS", **switches)
        c.sensor_update(scan.ranges, scan.angle_min, scan.angle_inc, scan.range_max)
        assert c.stats().sensor_path == 3
        w.set_params(_lib.params("NS", df_mode=4, **switches))
        w.set_map(lg)
        w.set_particles(x, y, th)
        w.sensor_update(scan.ranges, scan.angle_min, scan.angle_inc, scan.range_max)
        assert np.array_equal(c.raw_weights(), w.raw_weights())


@pytest.mark.parametrize("preset,n_scan", [("NS", 360), ("REF", 500), ("NS", 1080)])
def test_weights_bit_identical_to_walk_kernel(lg, preset, n_scan):
    """Spread-out and dense clouds, both presets: the persistent tables (auto, df_mode 0) and the walk
    kernel (df_mode 4) give the same raw weights bit for bit."""
    ux, uy, ut = workloads.uniform_particles(lg, 60_000, seed=23)
    gx, gy, gt = workloads.gaussian_particles(40_000)
    gt = np.random.Generator(np.random.PCG64(3)).uniform(-math.pi, math.pi, len(gt)).astype(np.float32)
    x, y, th = np.concatenate([ux, gx]), np.concatenate([uy, gy]), np.concatenate([ut, gt])
    scan = _scan(lg, n_scan)
    with _lib.Context(0) as c, _lib.Context(0) as w:
        c.set_params(_lib.params(preset))
        w.set_params(_lib.params(preset, df_mode=4))
        for q in (c, w):
            q.set_map(lg)
            q.set_particles(x, y, th)
            q.sensor_update(scan.ranges, scan.angle_min, scan.angle_inc, scan.range_max)
        assert c.stats().sensor_path == 3 and w.stats().sensor_path != 3
        assert np.array_equal(c.raw_weights(), w.raw_weights())


def test_filter_updates_agree_with_the_walk_kernel(lg):
    """Three full updates (motion -> sensor -> normalise -> systematic resample) on the persistent tables
    and on the walk kernel: identical particles and identical resample lists after every one."""
    w0 = workloads.c3(200_000)
    s = w0.scan
    with _lib.Context(0) as a, _lib.Context(0) as b:
        a.set_params(_lib.params("NS"))
        b.set_params(_lib.params("NS", df_mode=4))
        for q in (a, b):
            q.set_map(lg)
            q.set_particles(w0.x, w0.y, w0.theta)
        for k in range(3):
            for q in (a, b):
                q.filter_update((0.2 * (k + 1), 0.0, 0.1 * (k + 1)), (0.2 * k, 0.0, 0.1 * k), s.ranges, s.angle_min,
                                s.angle_inc, s.range_max, 0.37 + 0.1 * k, seed=11 + k)
            assert np.array_equal(a.resample_indices(), b.resample_indices())
            for u, v in zip(a.get_particles(), b.get_particles()):
                assert np.array_equal(u, v)
        assert a.stats().sensor_path == 3


def test_rebuild_on_new_map_and_new_ray_length(lg):
    """The tables are keyed by (map, ray length in cells, S3/S4): a new map or range_max refills them."""
    x, y, th = workloads.uniform_particles(lg, 800, seed=24)
    g2 = maps.OccupancyGrid(lg.info, lg.data.copy())
    g2.data[300:340, 300:340] = 100
    with _lib.Context(0) as c:
        c.set_params(_lib.params("NS", df_mode=7))
        c.set_map(lg)
        _ranges_vs_oracle(c, lg, x, y, th, _scan(lg), "NS")
        sc = workloads.make_scan(lg, workloads.TRUE_POSE, 90, -math.pi, 2 * math.pi / 90, 3.3)
        _ranges_vs_oracle(c, lg, x, y, th, sc, "NS")
        c.set_map(g2)
        _ranges_vs_oracle(c, g2, x, y, th, _scan(lg), "NS")


@pytest.mark.parametrize("seed", range(6))
def test_random_small_maps(seed):
    """Small random maps (odd sizes, odd resolutions): ranges from the persistent tables against the oracle."""
    rng = np.random.Generator(np.random.PCG64(100 + seed))
    W, H = int(rng.integers(9, 70)), int(rng.integers(9, 70))
    res = float(np.float32([0.05, 0.1, 0.025, 1.0, 0.3, 0.07][seed]))
    g = (rng.random((H, W)) < 0.08).astype(np.int8) * 100
    g[rng.random((H, W)) < 0.03] = -1
    grid = maps.OccupancyGrid(maps.MapMetaData(np.float32(res), W, H), g)
    n = 400
    x = (rng.uniform(-2, W + 2, n) * res).astype(np.float32)
    y = (rng.uniform(-2, H + 2, n) * res).astype(np.float32)
    th = rng.uniform(-7, 7, n).astype(np.float32)
    rmax = float(np.float32(res * rng.uniform(3, 25)))
    scan = workloads.Scan(rng.uniform(0, rmax, 77).astype(np.float32), -2.0, float(np.float32(4.0 / 77)), rmax)
    with _lib.Context(0) as c:
        c.set_params(_lib.params("NS", df_mode=7))
        c.set_map(grid)
        _ranges_vs_oracle(c, grid, x, y, th, scan, "NS")
        c.set_params(_lib.params("REF", df_mode=7, max_range=-rmax))
        _ranges_vs_oracle(c, grid, x, y, th, scan, "REF", max_range=-rmax)


def _bench_size_parity(lg, n, seed):
    """Sensor stage at a benchmarked size: oracle sample of 2,000 particles (weights to 1e-5), rays of a
    64-particle window bit-exact through the production look-up AND cell-exact through the walk kernel."""
    x, y, th = workloads.uniform_particles(lg, n, seed=seed)
    scan = _scan(lg)
    po = oracle.params("NS")
    with _lib.Context(0) as c:
        c.set_params(_lib.params("NS"))
        c.set_map(lg)
        c.set_particles(x, y, th)
        c.sensor_update(scan.ranges, scan.angle_min, scan.angle_inc, scan.range_max)
        st = c.stats()
        assert st.sensor_path == 3
        raw = c.raw_weights()
        sel = np.random.Generator(np.random.PCG64(seed)).choice(n, 2000, replace=False)
        ref, _, _ = oracle.sensor_f32(po, lg, scan.ranges, scan.angle_min, scan.angle_inc, scan.range_max,
                                      x[sel], y[sel], th[sel])
        np.testing.assert_allclose(raw[sel], ref, rtol=RTOL)
        ref64, _ = oracle.sensor_f64(po, lg, scan.ranges, scan.angle_min, scan.angle_inc, scan.range_max,
                                     x[sel], y[sel], th[sel])
        assert np.isclose(raw[sel], ref64, rtol=RTOL).mean() >= 0.9995  # near-tie refinement: the double path's rays
        first = n // 2
        _ranges_vs_oracle(c, lg, x, y, th, scan, "NS", first=first, count=64)
        # and the same window through the walk kernel's diagnostic pass: hit flag, cell, steps
        c.set_params(_lib.params("NS", df_mode=4))
        dbg = c.raycast_debug(360, scan.angle_min, scan.angle_inc, scan.range_max, first, 64)
        _, _, r = oracle.sensor_f32(po, lg, scan.ranges, scan.angle_min, scan.angle_inc, scan.range_max,
                                    x[first:first + 64], y[first:first + 64], th[first:first + 64], debug=True,
                                    want_raw=False)
        for k in ("hit", "hit_x", "hit_y", "steps", "range_m"):
            assert np.array_equal(dbg[k], r[k]), k


def test_c3_size_parity(lg):
    """BASELINE config C3: 1M particles uniform over free space x 360 beams."""
    _bench_size_parity(lg, 1_000_000, 31)


def test_c4_share_size_parity(lg):
    """BASELINE config C4 at its per-GPU share (10M over 8 GPUs): 1.25M particles x 360 beams."""
    _bench_size_parity(lg, 1_250_000, 32)


def test_zero_probability_scan_keeps_the_particles(lg):
    """an inf / NaN return under log-weights must not wipe out the cloud, and an update whose
    weights are all zero keeps the particles and says so (AMCLJ_ERR_STATE) instead of collapsing onto
    particle n-1."""
    w = workloads.c1(4000)
    s = w.scan
    bad = s.ranges.copy()
    bad[7], bad[50], bad[90] = np.inf, np.nan, 9.0  # beyond range_max
    po = oracle.params("NS")
    with _lib.Context(0) as c:
        c.set_params(_lib.params("NS"))
        c.set_map(lg)
        c.set_particles(w.x, w.y, w.theta)
        c.sensor_update(bad, s.angle_min, s.angle_inc, s.range_max)
        raw = c.raw_weights()
        assert np.isfinite(raw).all()
        ref, _, _ = oracle.sensor_f32(po, lg, bad, s.angle_min, s.angle_inc, s.range_max, w.x, w.y, w.theta)
        np.testing.assert_allclose(raw, ref, rtol=RTOL)
        # a scan that gives every particle probability zero: z_rand = z_max = 0 and an observation so far
        # from every map range that the hit Gaussian underflows
        c.set_params(_lib.params("NS", z_rand=0.0, z_max=0.0, z_short=0.0, sigma_hit=0.001))
        c.set_particles(w.x, w.y, w.theta)
        far = np.full_like(s.ranges, 5.5)
        with pytest.raises(_lib.AmcljError) as e:
            c.filter_update((0, 0, 0), (0, 0, 0), far, s.angle_min, s.angle_inc, s.range_max, 0.3)
        assert e.value.code == _lib.ERR_STATE
        x, y, t, _ = c.get_particles()
        assert np.array_equal(x, w.x) and np.array_equal(y, w.y) and np.array_equal(t, w.theta)
        assert c.stats().zero_weight_updates == 1
        # the next good update is back to normal
        c.set_params(_lib.params("NS"))
        c.filter_update(w.curr, w.prev, s.ranges, s.angle_min, s.angle_inc, s.range_max, 0.3, seed=2)


def test_fused_scan_and_update_call_equals_the_two_calls(lg):
    """amclj_filter_update_scan_async = amclj_stage_scan + amclj_filter_update_async, bit for bit: on its fast path
    (the geometry of the staged scan: motion kernel first, beam table built while it runs), with other ranges in the
    same geometry, and when the geometry changes (then it IS the two calls)."""
    w = workloads.c3(30_000)
    s = w.scan
    rng = np.random.Generator(np.random.PCG64(5))
    scans = [np.asarray(s.ranges, np.float32)]
    scans.append(np.clip(scans[0] + rng.normal(0, 0.3, len(scans[0])).astype(np.float32), 0, s.range_max).astype(np.float32))
    scans.append(scans[1][::2].copy())  # another beam count: a new geometry
    outs = []
    for fused in (False, True):
        with _lib.Context(0, "NS") as c:
            c.set_map(lg)
            c.set_particles(w.x, w.y, w.theta)
            got = []
            for it, r in enumerate(scans + [scans[0]]):
                inc = s.angle_inc * (2 if len(r) != len(scans[0]) else 1)
                if fused:
                    c.filter_update_scan_async(w.curr, w.prev, 40 + it, w.u0, r, s.angle_min, inc, s.range_max)
                else:
                    c.stage_scan(r, s.angle_min, inc, s.range_max)
                    c.filter_update_async(w.curr, w.prev, 40 + it, w.u0)
                c.synchronize()
                got.append(c.get_particles()[:3])
            assert c.stats().sensor_path == 3
            outs.append(got)
    for a, b in zip(*outs):
        for u, v in zip(a, b):
            assert np.array_equal(u, v)
    assert not np.array_equal(outs[0][0][0], outs[0][1][0])
