"""GPU parity: the CUDA path (through the C ABI, amclj_b200/_lib.py) against the CPU oracle
on identical seeded inputs.  Bars (BASELINE.json north_star): bit-exact for hit flag, hit
cell, steps, fixed-point weights and resample index lists; 1e-5 relative (fp32) for poses
and weights.
"""
import math

import numpy as np
import pytest

import oracle
from amclj_b200 import _lib, maps, workloads

pytestmark = pytest.mark.gpu

RTOL = 1e-5  # north_star: floating-point log-weights and poses within 1e-5 relative (fp32)


@pytest.fixture(scope="module")
def ctx():
    c = _lib.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="module")
def lg():
    return maps.lgfloor()


def _pair(preset, **kw):
    """Matching parameter blocks for the library and the oracle."""
    return _lib.params(preset, **kw), oracle.params(preset, **kw)


def _ray_parity(ctx, grid, x, y, th, scan, preset, **kw):
    pl, po = _pair(preset, **kw)
    ctx.set_params(pl)
    ctx.set_particles(x, y, th)
    dbg = ctx.raycast_debug(len(scan.ranges), scan.angle_min, scan.angle_inc, scan.range_max)
    _, st, ref = oracle.sensor_f32(po, grid, scan.ranges, scan.angle_min, scan.angle_inc,
                                   scan.range_max, x, y, th, debug=True, want_raw=False)
    for k in ("hit", "hit_x", "hit_y", "steps"):
        assert np.array_equal(dbg[k], ref[k]), f"{preset}: {k} differs at {np.argwhere(dbg[k] != ref[k])[:5]}"
    assert np.array_equal(dbg["range_m"], ref["range_m"]), "range in metres is exact in the fp32 contract"
    s = ctx.stats()
    assert (s.rays, s.cells_visited, s.hits) == (st.rays, st.cells, st.hits)
    assert s.oob_probes == 0  # every probe of every ray stayed inside the staged field
    return s


@pytest.mark.parametrize("preset", ["NS", "REF"])
def test_raycast_bit_exact_c1(ctx, lg, preset):
    w = workloads.c1()
    ctx.set_map(lg)
    s = _ray_parity(ctx, lg, w.x, w.y, w.theta, w.scan, preset)
    assert s.probes < s.cells_visited or preset == "REF"


@pytest.mark.parametrize("preset", ["NS", "REF"])
def test_raycast_bit_exact_uniform_and_edges(ctx, lg, preset):
    """Global-localisation poses plus the edge cases the walk must get right: particles on
    blocked / unknown cells, outside the map on every side, at the far corner, NaN."""
    x, y, th = workloads.uniform_particles(lg, 3000, seed=11)
    res = float(lg.info.resolution)
    W, H = lg.info.width, lg.info.height
    ex = np.array([-1.0, (W + 3) * res, 5.0, 5.0, 0.0, (W - 1) * res, (W - 0.6) * res, 16.3,
                   np.nan, 1e12, -0.024, 0.0251], np.float32)
    ey = np.array([5.0, 5.0, -2.0, (H + 10) * res, 0.0, (H - 1) * res, (H - 0.6) * res, np.nan,
                   3.0, -1e12, -0.024, 0.0251], np.float32)
    et = np.linspace(-4, 4, len(ex)).astype(np.float32)
    occ = np.argwhere(lg.data == 100)[::97][:40]
    x = np.concatenate([x, ex, (occ[:, 1] * res).astype(np.float32)])
    y = np.concatenate([y, ey, (occ[:, 0] * res).astype(np.float32)])
    th = np.concatenate([th, et, np.zeros(len(occ), np.float32)])
    ctx.set_map(lg)
    scan = workloads.make_scan(lg, workloads.TRUE_POSE, 360, -math.pi, 2 * math.pi / 360, 5.6)
    _ray_parity(ctx, lg, x, y, th, scan, preset)


@pytest.mark.parametrize("switches", [
    dict(s1_use_heading=0), dict(s3_proper_endpoint=0), dict(s4_endpoint_termination=0),
    dict(s2_use_scan_range_max=0), dict(s10_apply_laser_offset=1),
    dict(s3_proper_endpoint=0, s4_endpoint_termination=0), dict(s6_max_beams=30),
])
def test_raycast_switch_matrix(ctx, lg, switches):
    """Every semantic switch is a launch-uniform parameter of the same kernel."""
    x, y, th = workloads.uniform_particles(lg, 1500, seed=12)
    ctx.set_map(lg)
    scan = workloads.make_scan(lg, workloads.TRUE_POSE, 180, -math.pi / 2, math.pi / 179, 5.6)
    _ray_parity(ctx, lg, x, y, th, scan, "NS", **switches)


@pytest.mark.parametrize("df_mode", [1, 2, 3, 4, 5, 6])
def test_raycast_all_field_placements(ctx, lg, df_mode):
    """Shared-memory nibbles, global bytes, global nibbles and the directional (wedge) fields
    give the same answers."""
    x, y, th = workloads.uniform_particles(lg, 1500, seed=13)
    ctx.set_params(_lib.params("NS", df_mode=df_mode))
    ctx.set_map(lg)
    scan = workloads.make_scan(lg, workloads.TRUE_POSE, 360, -math.pi, 2 * math.pi / 360, 5.6)
    s = _ray_parity(ctx, lg, x, y, th, scan, "NS", df_mode=df_mode)
    assert s.map_in_smem == (1 if df_mode == 1 else 0)


def test_raycast_tiny_known_answer_map(ctx):
    """8x8 hand-checkable map (the oracle's known-answer fixture shape): a wall at x = 6."""
    g = np.zeros((8, 8), np.int8)
    g[:, 6] = 100
    g[0, :] = -1
    grid = maps.OccupancyGrid(maps.MapMetaData(np.float32(1.0), 8, 8), g)
    ctx.set_map(grid)
    ctx.set_params(_lib.params("NS"))
    ctx.set_particles(np.array([2.0], np.float32), np.array([3.0], np.float32), np.array([0.0], np.float32))
    d = ctx.raycast_debug(4, 0.0, math.pi / 2, 10.0)
    # beam 0 (+x): wall at x = 6 after 4 steps; beam 1 (+y): leaves the map at y = 8;
    # beam 2 (-x): leaves at x = -1; beam 3 (-y): unknown row y = 0
    assert d["hit"].tolist() == [[1, 1, 1, 1]]
    assert d["hit_x"].tolist() == [[6, 2, -1, 2]]
    assert d["hit_y"].tolist() == [[3, 8, 3, 0]]
    assert d["steps"].tolist() == [[4, 5, 3, 3]]
    np.testing.assert_array_equal(d["range_m"], np.array([[4, 5, 3, 3]], np.float32))


@pytest.mark.parametrize("preset", ["NS", "REF"])
def test_sensor_weights_c1(ctx, lg, preset):
    w = workloads.c1()
    pl, po = _pair(preset)
    ctx.set_map(lg)
    ctx.set_params(pl)
    ctx.set_particles(w.x, w.y, w.theta)
    ctx.sensor_update(w.scan.ranges, w.scan.angle_min, w.scan.angle_inc, w.scan.range_max)
    raw = ctx.raw_weights()
    ref, _, _ = oracle.sensor_f32(po, lg, w.scan.ranges, w.scan.angle_min, w.scan.angle_inc,
                                  w.scan.range_max, w.x, w.y, w.theta)
    np.testing.assert_allclose(raw, ref, rtol=RTOL)
    # and against the double restatement of the Clojure (same cells except where fp32
    # rounding moves a start/end cell): the bulk must agree to 1e-5 as well
    ref64, _ = oracle.sensor_f64(po, lg, w.scan.ranges, w.scan.angle_min, w.scan.angle_inc,
                                 w.scan.range_max, w.x, w.y, w.theta)
    close = np.isclose(raw, ref64, rtol=RTOL)
    # start cells are taken in double and end cells are re-taken in double next to a .5 boundary (contract.h), so
    # the fp32 path casts the rays the double path casts: every weight agrees, not "the bulk"
    assert close.mean() >= 0.999


def test_sensor_weights_c2_sample(ctx, lg):
    """BASELINE config C2 (100k x 360): full GPU pass, oracle on a 2,000-particle sample."""
    w = workloads.c2()
    pl, po = _pair("NS")
    ctx.set_map(lg)
    ctx.set_params(pl)
    ctx.set_particles(w.x, w.y, w.theta)
    ctx.sensor_update(w.scan.ranges, w.scan.angle_min, w.scan.angle_inc, w.scan.range_max)
    raw = ctx.raw_weights()
    sel = np.random.Generator(np.random.PCG64(7)).choice(w.n, 2000, replace=False)
    ref, _, _ = oracle.sensor_f32(po, lg, w.scan.ranges, w.scan.angle_min, w.scan.angle_inc,
                                  w.scan.range_max, w.x[sel], w.y[sel], w.theta[sel])
    np.testing.assert_allclose(raw[sel], ref, rtol=RTOL)


@pytest.mark.parametrize("curr,prev", [((0.2, 0.0, 0.1), (0.0, 0.0, 0.0)),
                                       ((1.0, 2.0, 0.3), (1.0, 2.0, -0.2)),  # trans == 0
                                       ((0.5, -0.25, 1.0), (0.75, 0.5, 2.5)),
                                       ((3.0, 1.0, 0.5), (3.0, 1.0, 0.5))])   # static
def test_motion_injected_normals(ctx, lg, curr, prev):
    w = workloads.c1(5000)
    nrm = w.normals()
    pl, po = _pair("REF")
    ctx.set_params(pl)
    ctx.set_particles(w.x, w.y, w.theta)
    ctx.motion_update(curr, prev, nrm)
    x, y, t, _ = ctx.get_particles()
    rx, ry, rt = oracle.motion_f32(po, curr, prev, nrm, w.x, w.y, w.theta)
    # same explicit-fma contract on both sides: identical bits, well inside 1e-5
    assert np.array_equal(x, rx) and np.array_equal(y, ry) and np.array_equal(t, rt)
    dx, dy, dt = oracle.motion_f64(po, curr, prev, nrm, w.x, w.y, w.theta)
    np.testing.assert_allclose(x, dx, rtol=RTOL)
    np.testing.assert_allclose(y, dy, rtol=RTOL)
    # the stored heading is kept in (-pi, pi] (what quat/to-heading returns); the double restatement adds up
    dt = dt - 2 * math.pi * np.round(dt / (2 * math.pi))
    np.testing.assert_allclose(t, dt, rtol=RTOL, atol=1e-6)


def test_motion_heading_stays_wrapped(ctx):
    """theta must not grow without bound over many updates (fp32 resolution, ray end cells)."""
    n = 4096
    rng = np.random.Generator(np.random.PCG64(17))
    x, y = rng.uniform(5, 25, n).astype(np.float32), rng.uniform(5, 25, n).astype(np.float32)
    th = rng.uniform(-math.pi, math.pi, n).astype(np.float32)
    pl, po = _pair("NS")
    ctx.set_params(pl)
    ctx.set_particles(x, y, th)
    rx, ry, rt = x.copy(), y.copy(), th.copy()
    for k in range(40):  # a robot turning on the spot: +0.9 rad per update
        nrm = rng.standard_normal((n, 3)).astype(np.float32)
        curr, prev = (1.0, 1.0, 0.9 * (k + 1)), (1.0, 1.0, 0.9 * k)
        ctx.motion_update(curr, prev, nrm)
        rx, ry, rt = oracle.motion_f32(po, curr, prev, nrm, rx, ry, rt)
    gx, gy, gt, _ = ctx.get_particles()
    assert np.array_equal(gx, rx) and np.array_equal(gy, ry) and np.array_equal(gt, rt)
    assert np.abs(gt).max() <= np.float32(math.pi) + 1e-6


def test_motion_philox_normals(ctx):
    w = workloads.c1(4096)
    pl, po = _pair("NS")
    ctx.set_params(pl)
    ctx.set_particles(w.x, w.y, w.theta)
    ctx.motion_update(w.curr, w.prev, None, seed=1234)
    x, y, t, _ = ctx.get_particles()
    nrm = oracle.motion_normals(1234, 0, w.n)
    rx, ry, rt = oracle.motion_f32(po, w.curr, w.prev, nrm, w.x, w.y, w.theta)
    np.testing.assert_allclose(x, rx, rtol=RTOL)
    np.testing.assert_allclose(y, ry, rtol=RTOL)
    np.testing.assert_allclose(t, rt, rtol=RTOL, atol=1e-6)
    assert abs(float(np.mean(nrm))) < 0.05 and abs(float(np.std(nrm)) - 1) < 0.05


def _weights(n, seed, spiky=False):
    rng = np.random.Generator(np.random.PCG64(seed))
    w = rng.random(n).astype(np.float32) ** (12 if spiky else 3)
    w[rng.integers(0, n, max(1, n // 50))] = 0.0
    return w


@pytest.mark.parametrize("n", [1, 2, 31, 2048, 2049, 100_000, 1_000_003])
def test_normalize_cdf_bit_exact(ctx, n):
    w = _weights(n, 40 + n % 7)
    if n == 1:
        w[:] = 0.25
    z = np.zeros(n, np.float32)
    ctx.set_particles(z, z, z, w)
    total, wmax = ctx.normalize(from_raw=False)
    q, rtotal = oracle.fixed_weights(w, oracle.wmax(w))
    assert wmax == float(oracle.wmax(w))
    assert total == rtotal
    assert np.array_equal(ctx.cdf(), np.cumsum(q, dtype=np.uint64))


@pytest.mark.parametrize("n,u0", [(1, 0.5), (1000, 0.0), (1000, 0.999999), (100_000, 0.37),
                                  (1_000_003, 0.8125)])
def test_systematic_indices_bit_exact(ctx, n, u0):
    w = _weights(n, 3, spiky=True)
    if n == 1:
        w[:] = 1.0
    x = np.arange(n, dtype=np.float32)
    ctx.set_particles(x, -x, x * 0.5, w)
    idx = ctx.resample(_lib.RESAMPLE_SYSTEMATIC, u0)
    ref = oracle.systematic_resample(w, u0)
    assert np.array_equal(idx.astype(np.int64), ref)
    assert (np.diff(idx) >= 0).all()
    gx, gy, gt, gw = ctx.get_particles()
    assert np.array_equal(gx, x[idx]) and np.array_equal(gy, -x[idx]) and np.array_equal(gt, (x * 0.5)[idx])
    np.testing.assert_allclose(gw, 1.0 / n, rtol=1e-7)


def test_systematic_from_log_weights(ctx, lg):
    """raw log-weights -> exp contract -> fixed point -> indices: exact given the raw values."""
    w = workloads.c1()
    pl, _ = _pair("NS")
    ctx.set_map(lg)
    ctx.set_params(pl)
    ctx.set_particles(w.x, w.y, w.theta)
    ctx.sensor_update(w.scan.ranges, w.scan.angle_min, w.scan.angle_inc, w.scan.range_max)
    raw = ctx.raw_weights()
    idx = ctx.resample(_lib.RESAMPLE_SYSTEMATIC, w.u0)
    lin, wmax = oracle.linear_weights(raw, True)
    q, total = oracle.fixed_weights(lin, wmax)
    r = oracle.systematic_offset(w.u0, total)
    ref = oracle.systematic(q, total, 0, r, w.n, 0, w.n)
    assert np.array_equal(idx.astype(np.int64), ref)


def test_roulette_injected_stream(ctx, golden):
    r = golden["roulette"]
    w = np.asarray(r["weights"], np.float32)
    z = np.zeros(len(w), np.float32)
    ctx.set_particles(z, z, z, w)
    idx = ctx.resample(_lib.RESAMPLE_ROULETTE, 0.0, stream_idx=r["attempt_idx"], stream_u=r["attempt_u32"])
    assert idx.tolist() == r["selected"]
    _, _, _, gw = ctx.get_particles()
    np.testing.assert_allclose(gw, np.asarray(r["normalized"])[idx], rtol=2e-6)


@pytest.mark.parametrize("n", [64, 1000, 20_000])
def test_roulette_philox_bit_exact(ctx, n):
    w = _weights(n, 5)
    z = np.zeros(n, np.float32)
    ctx.set_particles(z, z, z, w)
    idx = ctx.resample(_lib.RESAMPLE_ROULETTE, 0.0, seed=6)
    q, total = oracle.fixed_weights(w, oracle.wmax(w))
    ref, _ = oracle.roulette_philox(q, total, 6)
    assert np.array_equal(idx, ref)


def test_roulette_philox_bit_exact_c2_size(ctx):
    """pf.clj:156-167 at C2's particle count: N = 100,000 means about N^2 = 1e10 attempts
    (every attempt is accepted with probability 1/N whatever the weights); the GPU compacts them in
    attempt order, the oracle walks the same Philox stream on every host core."""
    n = 100_000
    w = _weights(n, 15)
    z = np.zeros(n, np.float32)
    ctx.set_particles(z, z, z, w)
    idx = ctx.resample(_lib.RESAMPLE_ROULETTE, 0.0, seed=16)
    q, total = oracle.fixed_weights(w, oracle.wmax(w))
    ref, attempts = oracle.roulette_philox(q, total, 16)
    assert attempts > 5e9
    assert np.array_equal(idx, ref)


def test_tournament_bit_exact(ctx):
    n = 50_000
    w = _weights(n, 8)
    z = np.zeros(n, np.float32)
    ctx.set_particles(z, z, z, w)
    idx = ctx.resample(_lib.RESAMPLE_TOURNAMENT, 0.0, seed=9)
    assert np.array_equal(idx, oracle.tournament_philox(w, 9))
    sidx = np.random.Generator(np.random.PCG64(2)).integers(0, n, 2 * n).astype(np.int32)
    ctx.set_particles(z, z, z, w)
    idx = ctx.resample(_lib.RESAMPLE_TOURNAMENT, 0.0, stream_idx=sidx)
    assert np.array_equal(idx, oracle.tournament_stream(w, sidx))


def test_pose_estimate(ctx):
    w = workloads.c1(10_000)
    ctx.set_particles(w.x, w.y, w.theta)
    np.testing.assert_allclose(ctx.pose_estimate(), oracle.pose_estimate(w.x, w.y, w.theta), rtol=1e-6)


@pytest.mark.parametrize("heading_mode", [0, 1])
def test_init_uniform(ctx, lg, heading_mode):
    ctx.set_map(lg)
    ctx.init_uniform(20_000, heading_mode, seed=77)
    x, y, t, w = ctx.get_particles()
    rx, ry, rt = oracle.uniform_init(lg, 0, 20_000, heading_mode, 77)
    assert np.array_equal(x, rx) and np.array_equal(y, ry)
    np.testing.assert_allclose(t, rt, rtol=RTOL, atol=1e-6)
    res = float(lg.info.resolution)
    assert (lg.data[(y / res).astype(int), (x / res).astype(int)] == 0).all()


def test_init_gaussian(ctx):
    ctx.init_gaussian(10_000, 16.3, 15.8, 2.0, 0.1, 0.0, seed=5)
    x, y, t, _ = ctx.get_particles()
    rx, ry, rt = oracle.gaussian_init(16.3, 15.8, 2.0, 0.1, 0.0, 0, 10_000, 5)
    np.testing.assert_allclose(x, rx, rtol=RTOL)
    np.testing.assert_allclose(y, ry, rtol=RTOL)
    assert np.array_equal(t, rt)


@pytest.mark.parametrize("preset", ["NS", "REF"])
def test_filter_update_stagewise(ctx, lg, preset):
    """One full mcl* hot-path iteration (pf.clj:229-233) through amclj_filter_update, checked
    stage by stage against the oracle fed with the same injected streams."""
    w = workloads.c1()
    nrm = w.normals()
    pl, po = _pair(preset)
    ctx.set_map(lg)
    ctx.set_params(pl)
    ctx.set_particles(w.x, w.y, w.theta)
    s = w.scan
    ctx.filter_update(w.curr, w.prev, s.ranges, s.angle_min, s.angle_inc, s.range_max, w.u0,
                      normals=nrm, seed=6)
    idx = ctx.resample_indices()
    gx, gy, gt, gw = ctx.get_particles()
    mx, my, mt = oracle.motion_f32(po, w.curr, w.prev, nrm, w.x, w.y, w.theta)
    raw, _, _ = oracle.sensor_f32(po, lg, s.ranges, s.angle_min, s.angle_inc, s.range_max, mx, my, mt)
    # resample the oracle's way from the GPU's own raw weights is covered above; here the
    # end-to-end list must agree wherever the two raw vectors round to the same fixed point
    lin, wmax = oracle.linear_weights(raw, bool(po.s7_log_weights))
    q, total = oracle.fixed_weights(lin, wmax)
    if preset == "NS":
        ref = oracle.systematic(q, total, 0, oracle.systematic_offset(w.u0, total), w.n, 0, w.n)
        assert (idx == ref).mean() > 0.98  # raw weights agree to 1e-5, not bit for bit
        np.testing.assert_allclose(gw, 1.0 / w.n, rtol=1e-7)
    assert np.array_equal(gx, mx[idx]) and np.array_equal(gy, my[idx]) and np.array_equal(gt, mt[idx])
    st = ctx.stats()
    assert st.ms_total > 0 and st.n_particles == w.n


def test_filter_update_host_roundtrip(ctx, lg):
    w = workloads.c1()
    ctx.set_map(lg)
    ctx.set_params(_lib.params("NS"))
    x, y, t = w.x.copy(), w.y.copy(), w.theta.copy()
    wt = np.zeros(w.n, np.float32)
    s = w.scan
    ctx.filter_update_host(x, y, t, wt, w.curr, w.prev, s.ranges, s.angle_min, s.angle_inc,
                           s.range_max, w.u0, seed=3)
    gx, gy, gt, gw = ctx.get_particles()
    assert np.array_equal(x, gx) and np.array_equal(y, gy) and np.array_equal(t, gt) and np.array_equal(wt, gw)
    # resampled cloud concentrates near the true pose
    assert np.hypot(x.mean() - (workloads.TRUE_POSE[0] + 0.2 * math.cos(workloads.TRUE_POSE[2])),
                    y.mean() - (workloads.TRUE_POSE[1] + 0.2 * math.sin(workloads.TRUE_POSE[2]))) < 0.5


def test_errors_are_loud(ctx):
    with pytest.raises(_lib.AmcljError):
        ctx.resample(_lib.RESAMPLE_SYSTEMATIC, 1.5)
    z = np.zeros(4, np.float32)
    ctx.set_particles(z, z, z, z)  # all weights zero
    with pytest.raises(_lib.AmcljError):
        ctx.resample(_lib.RESAMPLE_SYSTEMATIC, 0.5)
    with pytest.raises(_lib.AmcljError):
        _lib.Context(99)


def test_host_mirror_of_reference_api(lg, golden, fixture_scan):
    """amclj_b200.pf keeps the reference's function names (pf.clj, sensor.clj, motion.clj); the
    dev-fixture particle of sensor.clj:69-74 scores the survey's REF weight through it."""
    from amclj_b200 import pf
    scan = workloads.Scan(fixture_scan["ranges"], fixture_scan["angle_min"], fixture_scan["angle_inc"],
                          fixture_scan["range_max"])
    cloud = pf.ParticleCloud(0, "REF")
    e = golden["sensor"][0]
    wgt = pf.likelihood_field_range_finder_model(scan, (e["x"], e["y"], e["theta"]), lg, cloud)
    assert wgt == pytest.approx(e["ref"]["weight"], rel=RTOL)
    pf.initialize_pf(cloud, lg, 1500, pose=workloads.TRUE_POSE, seed=3)
    assert len(cloud) == pf.NUM_PARTICLES
    pf.apply_motion_model([(0.2, 0.0, 0.1), (0.0, 0.0, 0.0)], cloud, seed=4)
    pf.apply_sensor_model(scan, lg, cloud)
    assert np.isfinite(pf.weights(cloud)).all()
    idx = pf.roulette_selection(cloud, seed=5)
    assert len(idx) == 1500 and idx.min() >= 0 and idx.max() < 1500
    est = pf.pose_estimate(cloud)
    assert abs(est["position"]["x"] - workloads.TRUE_POSE[0]) < 1.0
    cloud.close()


def test_c5_large_map_global_field():
    """BASELINE config C5 shape: synthetic 8192 x 8192 grid (field too large for shared memory
    -> byte field through L2), 1080-beam 270-degree scan, 30 m rays (600 cells).  Full GPU pass
    on 20k particles; oracle on a sample: rays bit-exact, weights 1e-5."""
    w = workloads.c5(20_000)
    s = w.scan
    pl, po = _pair("NS")
    c = _lib.Context(0)
    c.set_params(pl)
    c.set_map(w.grid)
    c.set_particles(w.x, w.y, w.theta)
    dbg = c.raycast_debug(len(s.ranges), s.angle_min, s.angle_inc, s.range_max, 0, 64)
    _, _, ref = oracle.sensor_f32(po, w.grid, s.ranges, s.angle_min, s.angle_inc, s.range_max,
                                  w.x[:64], w.y[:64], w.theta[:64], debug=True, want_raw=False)
    for k in ("hit", "hit_x", "hit_y", "steps", "range_m"):
        assert np.array_equal(dbg[k], ref[k]), k
    assert c.stats().map_in_smem == 0
    c.sensor_update(s.ranges, s.angle_min, s.angle_inc, s.range_max)
    raw = c.raw_weights()
    sel = np.arange(0, w.n, 97)
    rraw, _, _ = oracle.sensor_f32(po, w.grid, s.ranges, s.angle_min, s.angle_inc, s.range_max,
                                   w.x[sel], w.y[sel], w.theta[sel])
    np.testing.assert_allclose(raw[sel], rraw, rtol=RTOL)
    c.close()


def test_c5_full_size_parity():
    """The same at BASELINE's full C5 size (4 M particles: the walk kernel then runs its particles in map-tile
    order on the L2-resident byte field, with the tile side taken from the particle density): rays of a 64-particle
    window bit-exact, weights of every 13,001st particle within 1e-5 of the oracle, no out-of-bounds probe."""
    w = workloads.c5(4_000_000)
    s = w.scan
    pl, po = _pair("NS")
    c = _lib.Context(0)
    c.set_params(pl)
    c.set_map(w.grid)
    c.set_particles(w.x, w.y, w.theta)
    first = 1_234_567
    dbg = c.raycast_debug(len(s.ranges), s.angle_min, s.angle_inc, s.range_max, first, 64)
    _, _, ref = oracle.sensor_f32(po, w.grid, s.ranges, s.angle_min, s.angle_inc, s.range_max,
                                  w.x[first:first + 64], w.y[first:first + 64], w.theta[first:first + 64], debug=True, want_raw=False)
    for k in ("hit", "hit_x", "hit_y", "steps", "range_m"):
        assert np.array_equal(dbg[k], ref[k]), k
    c.sensor_update(s.ranges, s.angle_min, s.angle_inc, s.range_max)
    st = c.stats()
    assert st.map_in_smem == 0 and st.sensor_path == 0  # walk kernel, particles in map-tile order
    raw = c.raw_weights()
    sel = np.arange(0, w.n, 13_001)
    rraw, _, _ = oracle.sensor_f32(po, w.grid, s.ranges, s.angle_min, s.angle_inc, s.range_max,
                                   w.x[sel], w.y[sel], w.theta[sel])
    np.testing.assert_allclose(raw[sel], rraw, rtol=RTOL)
    c.close()


def test_round_trip_properties_c3_size():
    """Size-independent properties at BASELINE size C3 (1M particles): systematic indices are
    non-decreasing, every source index is valid, the offspring count of a particle is within 1 of
    N * w_i / sum(w), and resampling twice with the same u0 from the same weights is idempotent."""
    n = 1_000_000
    w = _weights(n, 21)
    x = np.arange(n, dtype=np.float32)
    c = _lib.Context(0, "NS")
    c.set_particles(x, x, x, w)
    total, wmax = c.normalize(from_raw=False)
    idx = c.resample(_lib.RESAMPLE_SYSTEMATIC, 0.123456)
    assert (np.diff(idx) >= 0).all() and idx[0] >= 0 and idx[-1] < n
    counts = np.bincount(idx, minlength=n)
    q, t2 = oracle.fixed_weights(w, oracle.wmax(w))
    assert total == t2
    expect = q.astype(np.float64) * n / float(total)
    assert np.all(np.abs(counts - expect) <= 1.0 + 1e-9)
    c.set_particles(x, x, x, w)
    idx2 = c.resample(_lib.RESAMPLE_SYSTEMATIC, 0.123456)
    assert np.array_equal(idx, idx2)
    c.close()


@pytest.mark.parametrize("n", [1, 2, 1001, 100_000])
def test_median_weight(ctx, n):
    w = _weights(n, 31)
    z = np.zeros(n, np.float32)
    ctx.set_particles(z, z, z, w)
    assert ctx.median_weight() == oracle.median(w)


@pytest.mark.parametrize("scale", [1.0, 0.02])
def test_augmented_selection(ctx, lg, scale):
    """pf.clj:246-258: w-slow/w-fast update, particle injection and roulette-step, slot by slot."""
    n = 20_000
    x, y, th = workloads.uniform_particles(lg, n, seed=5)
    w = (_weights(n, 32) * scale).astype(np.float32)
    ctx.set_map(lg)
    ctx.set_particles(x, y, th, w)
    ctx.augment_state((1.0, 1.0))
    slow, fast = 1.0, 1.0
    for it in range(3):  # the averages drift apart, so later rounds inject particles
        gx0, gy0, gt0, gw0 = ctx.get_particles()
        idx = ctx.augmented_resample(0.01, 0.1, seed=40 + it)
        rx, ry, rt, rw, ridx, slow, fast = oracle.augmented(gw0, gx0, gy0, gt0, lg, slow, fast, 0.01, 0.1, 40 + it)
        assert np.array_equal(idx, ridx)
        gx, gy, gt, gw = ctx.get_particles()
        assert np.array_equal(gx, rx) and np.array_equal(gy, ry) and np.array_equal(gw, rw)
        np.testing.assert_allclose(gt, rt, rtol=RTOL, atol=1e-6)
        s2, f2 = ctx.augment_state()
        assert s2 == pytest.approx(slow, rel=1e-12) and f2 == pytest.approx(fast, rel=1e-12)
    assert (idx == -1).any() or scale == 1.0


@pytest.mark.parametrize("preset,switches", [("NS", {}), ("REF", {}), ("NS", dict(s4_endpoint_termination=0)),
                                             ("NS", dict(s3_proper_endpoint=0, s4_endpoint_termination=0)),
                                             ("NS", dict(s1_use_heading=0, s10_apply_laser_offset=1))])
def test_raycast_directional_fields_exact(ctx, lg, preset, switches):
    """df_mode 4: the walk jumps by direction-class (wedge) distances; cells, steps and ranges
    stay bit-identical to the reference's cell-by-cell Bresenham in every semantic mode."""
    x, y, th = workloads.uniform_particles(lg, 4000, seed=17)
    res = float(lg.info.resolution)
    W, H = lg.info.width, lg.info.height
    ex = np.array([-1.0, (W + 3) * res, 5.0, 0.0, (W - 0.6) * res, np.nan, 1e12], np.float32)
    ey = np.array([5.0, 5.0, -2.0, 0.0, (H - 0.6) * res, 3.0, -1e12], np.float32)
    x, y, th = np.concatenate([x, ex]), np.concatenate([y, ey]), np.concatenate([th, np.zeros(len(ex), np.float32)])
    ctx.set_map(lg)
    scan = workloads.make_scan(lg, workloads.TRUE_POSE, 360, -math.pi, 2 * math.pi / 360, 5.6)
    s = _ray_parity(ctx, lg, x, y, th, scan, preset, df_mode=4, **switches)
    assert s.map_in_smem == 0
    if preset == "NS" and not switches:
        assert s.probes < 0.7 * 7.4 * s.rays  # fewer probes than the isotropic field needs


def test_auto_field_policy_matches_every_placement(lg):
    """df_mode 0 on a small map walks the directional planes — in place for a compact particle
    cloud, in map-tile order for a spread-out one (decided on the device from the bounding box).
    Weights agree with the oracle and, bit for bit, with every forced placement."""
    c = _lib.Context(0)
    pl, po = _pair("NS", collect_stats=1)
    c.set_params(pl)
    c.set_map(lg)
    scan = workloads.make_scan(lg, workloads.TRUE_POSE, 360, -math.pi, 2 * math.pi / 360, 5.6)
    for spread in (False, True):
        x, y, th = workloads.uniform_particles(lg, 3000, seed=3) if spread else workloads.gaussian_particles(3000)
        c.set_particles(x, y, th)
        c.sensor_update(scan.ranges, scan.angle_min, scan.angle_inc, scan.range_max)
        st = c.stats()
        assert st.map_in_smem == 0 and st.probes < 5 * st.rays
        raw = c.raw_weights()
        ref, _, _ = oracle.sensor_f32(po, lg, scan.ranges, scan.angle_min, scan.angle_inc, scan.range_max, x, y, th)
        np.testing.assert_allclose(raw, ref, rtol=RTOL)
        for forced in (1, 2, 4, 5, 6):
            c2 = _lib.Context(0)
            c2.set_params(_lib.params("NS", df_mode=forced))
            c2.set_map(lg)
            c2.set_particles(x, y, th)
            c2.sensor_update(scan.ranges, scan.angle_min, scan.angle_inc, scan.range_max)
            assert np.array_equal(c2.raw_weights(), raw), forced
            c2.close()
    c.close()


@pytest.mark.parametrize("seed", range(20))
def test_raycast_random_small_maps(seed):
    """Random small maps of awkward shapes (1-cell-wide, odd widths, coarse and fine resolutions),
    random scans and switches; every field placement against the oracle, ray by ray."""
    rng = np.random.Generator(np.random.PCG64(1000 + seed))
    W = int(rng.choice([1, 2, 3, 7, 16, 33, 64, 129]))
    H = int(rng.choice([1, 2, 5, 9, 31, 64, 100]))
    res = float(rng.choice([0.05, 0.1, 0.25, 1.0, 2.5]))
    g = np.where(rng.random((H, W)) < 0.12, 100, 0).astype(np.int8)
    g[rng.random((H, W)) < 0.05] = -1
    grid = maps.OccupancyGrid(maps.MapMetaData(np.float32(res), W, H), g)
    n = 600
    x = (rng.uniform(-1.5, W + 1.5, n) * res).astype(np.float32)
    y = (rng.uniform(-1.5, H + 1.5, n) * res).astype(np.float32)
    th = rng.uniform(-7, 7, n).astype(np.float32)
    nbeams = int(rng.choice([1, 2, 17, 64, 181]))
    rmax = float(rng.choice([0.3, 2.0, 11.0])) * res * 10
    scan = workloads.Scan((rng.uniform(0, rmax, nbeams)).astype(np.float32), float(np.float32(rng.uniform(-3.2, 0))),
                          float(np.float32(rng.uniform(0.001, 6.3 / max(nbeams, 2)))), float(np.float32(rmax)))
    preset = "NS" if seed % 2 == 0 else "REF"
    switches = {} if seed % 3 else dict(s4_endpoint_termination=seed % 2, s3_proper_endpoint=(seed // 2) % 2)
    c = _lib.Context(0)
    for df_mode in (1, 2, 3, 4, 5, 6, 0):
        c.set_params(_lib.params(preset, df_mode=df_mode, **switches))
        c.set_map(grid)
        _ray_parity(c, grid, x, y, th, scan, preset, df_mode=df_mode, **switches)
    c.close()


@pytest.mark.parametrize("seed", range(6))
def test_resample_random_weight_patterns(ctx, seed):
    """Odd particle counts and hostile weight vectors (NaN, negative, infinities, one survivor,
    all equal, denormals): fixed-point CDF, systematic and tournament lists match the oracle."""
    rng = np.random.Generator(np.random.PCG64(500 + seed))
    n = int(rng.choice([1, 2, 3, 31, 33, 257, 2047, 2049, 4097, 70_001]))
    w = (rng.random(n) ** rng.choice([1, 4, 20])).astype(np.float32)
    kind = seed % 6
    if kind == 1:
        w[:] = 0.0
        w[rng.integers(0, n)] = 3.5
    elif kind == 2:
        w[:] = 0.125
    elif kind == 3 and n > 3:
        w[0], w[1], w[2] = np.nan, -1.0, 1e-42
    elif kind == 4 and n > 2:
        w[rng.integers(0, n)] = 1e30
    z = np.zeros(n, np.float32)
    ctx.set_particles(z, z, z, w)
    total, wmax = ctx.normalize(from_raw=False)
    q, rtotal = oracle.fixed_weights(w, oracle.wmax(w))
    assert total == rtotal and np.array_equal(ctx.cdf(), np.cumsum(q, dtype=np.uint64))
    if rtotal > 0:
        u0 = float(rng.random())
        idx = ctx.resample(_lib.RESAMPLE_SYSTEMATIC, u0)
        assert np.array_equal(idx.astype(np.int64), oracle.systematic_resample(w, u0))
    wt = np.nan_to_num(w, nan=0.0)
    ctx.set_particles(z, z, z, wt)
    assert np.array_equal(ctx.resample(_lib.RESAMPLE_TOURNAMENT, 0.0, seed=seed), oracle.tournament_philox(wt, seed))


@pytest.mark.parametrize("preset", ["NS", "REF"])
def test_raycast_very_long_rays_use_the_shifted_reciprocal(preset):
    """Rays of 20,000 cells (fine resolution, long range) push the minor-axis division past the
    shift-free reciprocal's exact range: the table switches to the shifted form; still bit-exact."""
    rng = np.random.Generator(np.random.PCG64(77))
    W, H, res = 300, 200, 0.005
    g = np.where(rng.random((H, W)) < 0.002, 100, 0).astype(np.int8)
    grid = maps.OccupancyGrid(maps.MapMetaData(np.float32(res), W, H), g)
    n = 800
    x = (rng.uniform(0, W, n) * res).astype(np.float32)
    y = (rng.uniform(0, H, n) * res).astype(np.float32)
    th = rng.uniform(-3.2, 3.2, n).astype(np.float32)
    scan = workloads.Scan(rng.uniform(0, 2, 90).astype(np.float32), float(np.float32(-3.1)), float(np.float32(0.07)),
                          float(np.float32(100.0)))
    c = _lib.Context(0)
    for df_mode in (1, 2, 4, 6, 0):
        kw = dict(df_mode=df_mode, max_range=-100.0) if preset == "REF" else dict(df_mode=df_mode)
        c.set_params(_lib.params(preset, **kw))
        c.set_map(grid)
        _ray_parity(c, grid, x, y, th, scan, preset, **kw)
    c.close()


def test_host_register_same_results(ctx, lg):
    """amclj_host_register: a plain (pageable) host buffer page-locked by the library gives the same update as
    an unregistered one; registering twice and unregistering an unknown buffer are not errors."""
    w = workloads.c1(4000)
    s = w.scan
    ctx.set_params(_lib.params("NS"))
    ctx.set_map(lg)
    outs = []
    for reg in (False, True):
        x, y, t = w.x.copy(), w.y.copy(), w.theta.copy()
        wt = np.zeros_like(x)
        if reg:
            for a in (x, y, t, wt):
                ctx.host_register(a)
            ctx.host_register(x)  # again: already registered
        ctx.set_particles(x, y, t)
        ctx.filter_update_host(x, y, t, wt, w.curr, w.prev, s.ranges, s.angle_min, s.angle_inc, s.range_max, w.u0, seed=11)
        if reg:
            for a in (x, y, t, wt):
                ctx.host_unregister(a)
            ctx.host_unregister(x)  # again: not registered any more
        outs.append((x, y, t, wt))
    for a, b in zip(*outs):
        assert np.array_equal(a, b)
    assert not np.array_equal(outs[0][0], w.x)  # the update did move the cloud
