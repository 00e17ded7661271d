"""GPU parity of the ray-table path (amclj_b200/csrc/raytable.cu): for a dense cloud the range
of a ray is looked up in a table filled once per (start cell, end cell) — sensor.clj:130-139
rounds both ends to cells before anything else looks at them.  The table is an accelerator, never an
approximation: every ray against the oracle (hit flag, cell, steps, metres: bit-exact), weights
bit-identical to the walk kernel, and the policy that picks a path never changes a result.
"""
import math

import numpy as np
import pytest

import oracle
from amclj_b200 import _lib, maps, workloads

pytestmark = pytest.mark.gpu

RTOL = 1e-5  # north_star: fp32 log-weights within 1e-5 relative


@pytest.fixture(scope="module")
def lg():
    return maps.lgfloor()


def _scan(lg, n=360):
    return workloads.make_scan(lg, workloads.TRUE_POSE, n, -math.pi, 2 * math.pi / n, 5.6)


def _rays_vs_oracle(c, grid, x, y, th, scan, preset, **kw):
    pl, po = _lib.params(preset, **kw), oracle.params(preset, **{k: v for k, v in kw.items()})
    c.set_params(pl)
    c.set_particles(x, y, th)
    dbg = c.raycast_debug(len(scan.ranges), scan.angle_min, scan.angle_inc, scan.range_max)
    _, st, ref = oracle.sensor_f32(po, grid, scan.ranges, scan.angle_min, scan.angle_inc, scan.range_max,
                                   x, y, th, debug=True, want_raw=False)
    for k in ("hit", "hit_x", "hit_y", "steps"):
        assert np.array_equal(dbg[k], ref[k]), f"{preset}: {k} differs at {np.argwhere(dbg[k] != ref[k])[:5]}"
    assert np.array_equal(dbg["range_m"], ref["range_m"])
    s = c.stats()
    assert (s.rays, s.cells_visited, s.hits) == (st.rays, st.cells, st.hits)
    assert s.oob_probes == 0
    return s


@pytest.mark.parametrize("preset", ["NS", "REF"])
def test_dense_cloud_every_ray_through_the_tables(lg, preset):
    """Pose-tracking cloud, tables forced for every occupied cell: each ray's hit flag, cell, steps
    and range come out of a table and equal the oracle's cell-by-cell walk."""
    x, y, th = workloads.gaussian_particles(4000, sd_theta=0.3)
    with _lib.Context(0) as c:
        c.set_params(_lib.params(preset, df_mode=6))
        c.set_map(lg)
        s = _rays_vs_oracle(c, lg, x, y, th, _scan(lg), preset, df_mode=6)
        assert s.table_particles == len(x) and 50 < s.table_cells < 2000
        assert s.table_misses == 0  # every end cell was inside the ring the tables cover


@pytest.mark.parametrize("switches", [
    dict(s1_use_heading=0), dict(s3_proper_endpoint=0), dict(s4_endpoint_termination=0),
    dict(s2_use_scan_range_max=0), dict(s10_apply_laser_offset=1), dict(s6_max_beams=30),
    dict(s3_proper_endpoint=0, s4_endpoint_termination=0, s7_log_weights=0),
])
def test_dense_cloud_switch_matrix(lg, switches):
    """Every semantic switch through the tables (S10 moves the ray's start off the particle's cell:
    those rays are walked, the result is the same)."""
    x, y, th = workloads.gaussian_particles(1500, sd_theta=1.0, seed=21)
    scan = workloads.make_scan(lg, workloads.TRUE_POSE, 180, -math.pi / 2, math.pi / 179, 5.6)
    with _lib.Context(0) as c:
        c.set_params(_lib.params("NS", df_mode=6))
        c.set_map(lg)
        _rays_vs_oracle(c, lg, x, y, th, scan, "NS", df_mode=6, **switches)
        c.sensor_update(scan.ranges, scan.angle_min, scan.angle_inc, scan.range_max)
        raw = c.raw_weights()
        c.set_params(_lib.params("NS", df_mode=4, **switches))
        c.sensor_update(scan.ranges, scan.angle_min, scan.angle_inc, scan.range_max)
        assert np.array_equal(c.raw_weights(), raw)


def test_cloud_with_outliers_and_edge_particles(lg):
    """A dense core plus stragglers inside the bounding box (sparse cells: walked), particles on
    blocked cells and just outside the map."""
    rng = np.random.Generator(np.random.PCG64(31))
    x, y, th = workloads.gaussian_particles(3000, sd_theta=0.2, seed=5)
    ox = (workloads.TRUE_POSE[0] + rng.uniform(-3, 3, 200)).astype(np.float32)
    oy = (workloads.TRUE_POSE[1] + rng.uniform(-3, 3, 200)).astype(np.float32)
    ot = rng.uniform(-3.2, 3.2, 200).astype(np.float32)
    x, y, th = np.concatenate([x, ox]), np.concatenate([y, oy]), np.concatenate([th, ot])
    with _lib.Context(0) as c:
        c.set_params(_lib.params("NS", df_mode=6))
        c.set_map(lg)
        _rays_vs_oracle(c, lg, x, y, th, _scan(lg), "NS", df_mode=6)
    # a cloud hugging the map corner: start cells outside the map get tables too (immediate hits)
    res = float(lg.info.resolution)
    cx = (rng.uniform(-0.3, 0.3, 800)).astype(np.float32)
    cy = (rng.uniform(-0.3, 0.3, 800) + (lg.info.height - 1) * res).astype(np.float32)
    ct = rng.uniform(-3.2, 3.2, 800).astype(np.float32)
    with _lib.Context(0) as c:
        c.set_params(_lib.params("NS", df_mode=6))
        c.set_map(lg)
        _rays_vs_oracle(c, lg, cx, cy, ct, _scan(lg, 90), "NS", df_mode=6)


def test_auto_policy_takes_the_tables_for_a_dense_cloud_only(lg):
    """df_mode 0: a dense cloud runs on the tables, a spread-out one on the walk kernel (decided on
    the device each update); weights are bit-identical to the forced walk either way, and within
    1e-5 of the oracle."""
    scan = _scan(lg)
    po = oracle.params("NS")
    with _lib.Context(0) as c, _lib.Context(0) as w:
        c.set_params(_lib.params("NS", df_mode=0, collect_stats=1))
        w.set_params(_lib.params("NS", df_mode=4))
        c.set_map(lg)
        w.set_map(lg)
        clouds = [workloads.gaussian_particles(30_000, sd_theta=0.05), workloads.uniform_particles(lg, 5000, seed=3),
                  workloads.gaussian_particles(30_000, sd_theta=0.05, seed=8), workloads.gaussian_particles(30_000, seed=9),
                  workloads.gaussian_particles(150, seed=10)]
        took = []
        for x, y, th in clouds:
            for ctx in (c, w):
                ctx.set_particles(x, y, th)
                ctx.sensor_update(scan.ranges, scan.angle_min, scan.angle_inc, scan.range_max)
            raw = c.raw_weights()
            assert np.array_equal(raw, w.raw_weights())
            took.append(c.stats().table_cells > 0)
            if len(x) <= 5000:
                ref, _, _ = oracle.sensor_f32(po, lg, scan.ranges, scan.angle_min, scan.angle_inc, scan.range_max, x, y, th)
                np.testing.assert_allclose(raw, ref, rtol=RTOL)
        # dense, spread, dense (the hint said "spread": this update still walks), dense again, too few per cell
        assert took[0] and not took[1] and took[3] and not took[4]


def test_filter_updates_agree_with_the_walk_kernel(lg):
    """Five full updates (motion, sensor, normalise, systematic resample) of a tracking cloud on the
    auto policy against the forced walk kernel: identical particles and resample indices throughout."""
    wl = workloads.c2(40_000)
    s = wl.scan
    with _lib.Context(0) as a, _lib.Context(0) as b:
        a.set_params(_lib.params("NS", df_mode=0))
        b.set_params(_lib.params("NS", df_mode=4))
        for ctx in (a, b):
            ctx.set_map(lg)
            ctx.set_particles(wl.x, wl.y, wl.theta)
        for step in range(5):
            for ctx in (a, b):
                ctx.filter_update(wl.curr, wl.prev, s.ranges, s.angle_min, s.angle_inc, s.range_max, 0.37 + 0.1 * step, seed=50 + step)
            pa, pb = a.get_particles(), b.get_particles()
            for u, v in zip(pa, pb):
                assert np.array_equal(u, v), step
            assert np.array_equal(a.resample_indices(), b.resample_indices())


def test_small_maps_and_rings(lg):
    """Tiny maps and tiny rings (rays shorter than the ring margin), dense clouds on them."""
    rng = np.random.Generator(np.random.PCG64(41))
    for W, H, res, rmax in ((9, 7, 0.25, 0.3), (33, 20, 0.1, 1.0), (64, 64, 1.0, 40.0), (5, 1, 0.05, 0.02)):
        g = np.where(rng.random((H, W)) < 0.1, 100, 0).astype(np.int8)
        grid = maps.OccupancyGrid(maps.MapMetaData(np.float32(res), W, H), g)
        n = 700
        x = (rng.uniform(-1.0, W + 1.0, n) * res).astype(np.float32)
        y = (rng.uniform(-1.0, H + 1.0, n) * res).astype(np.float32)
        th = rng.uniform(-7, 7, n).astype(np.float32)
        scan = workloads.Scan(rng.uniform(0, rmax, 33).astype(np.float32), -1.7, 0.11, float(np.float32(rmax)))
        for preset in ("NS", "REF"):
            kw = dict(df_mode=6) if preset == "NS" else dict(df_mode=6, max_range=-float(rmax))
            with _lib.Context(0) as c:
                c.set_params(_lib.params(preset, **kw))
                c.set_map(grid)
                s = _rays_vs_oracle(c, grid, x, y, th, scan, preset, **kw)
                assert s.table_particles == n and s.table_misses == 0
