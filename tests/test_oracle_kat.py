"""Hand-derived known answers for the oracle on tiny maps (cell lists written out by hand
from sensor.clj:112-139), Philox known-answer vectors (Random123 kat_vectors), and
structural properties of the integer resamplers."""
import math

import numpy as np
import pytest

import oracle
from amclj_b200.maps import MapMetaData, OccupancyGrid


def tiny(rows, res=1.0):
    """rows[0] is grid row y=0.  '.' free, '#' occupied, '?' unknown."""
    lut = {".": 0, "#": 100, "?": -1}
    data = np.array([[lut[c] for c in r] for r in rows], np.int8)
    return OccupancyGrid(MapMetaData(np.float32(res), data.shape[1], data.shape[0]), data)


def test_philox_known_answers():
    assert [hex(v) for v in oracle.philox([0] * 4, [0] * 2)] == [
        "0x6627e8d5", "0xe169c58d", "0xbc57ac4c", "0x9b00dbd8"]
    assert [hex(v) for v in oracle.philox([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2)] == [
        "0x408f276d", "0x41c83b0e", "0xa20bc7c6", "0x6d5451fd"]
    assert [hex(v) for v in oracle.philox([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344],
                                           [0xA4093822, 0x299F31D0])] == [
        "0xd16cfe09", "0x94fdcceb", "0x5001e420", "0x24126ea1"]


def test_blocked_contract():
    g = tiny(["..#", "?.."])
    cells = g.data
    assert oracle.lib().amclj_oracle_blocked(cells, 3, 2, 0, 0) == 0
    assert oracle.lib().amclj_oracle_blocked(cells, 3, 2, 2, 0) == 1   # occupied
    assert oracle.lib().amclj_oracle_blocked(cells, 3, 2, 0, 1) == 1   # unknown stops rays too
    for x, y in ((-1, 0), (3, 0), (0, -1), (0, 2)):                   # documented bounds contract
        assert oracle.lib().amclj_oracle_blocked(cells, 3, 2, x, y) == 1


def test_ns_horizontal_ray_hits_wall():
    g = tiny(["........#."] * 3)
    p = oracle.params("ns")
    r = oracle.map_range_f64(p, g, 1.0, 1.0, 0.0, 0.0, 20.0)  # +x from cell (1,1)
    assert (r.hit, r.hit_x, r.hit_y, r.steps, r.cells) == (1, 8, 1, 7, 8)
    assert r.range == pytest.approx(7.0)


def test_ns_miss_returns_range_max_after_endpoint_cell():
    g = tiny(["." * 12] * 3)
    p = oracle.params("ns")
    r = oracle.map_range_f64(p, g, 1.0, 1.0, 0.0, 0.0, 5.0)  # cells 1..6 tested, all free
    assert (r.hit, r.hit_x, r.hit_y, r.steps, r.cells) == (0, 6, 1, 5, 6)
    assert r.range == 5.0


def test_ns_heading_rotates_ray():
    g = tiny(["....."] * 4 + ["#####"])
    p = oracle.params("ns")
    r = oracle.map_range_f64(p, g, 2.0, 0.0, math.pi / 2, 0.0, 10.0)  # straight up (steep)
    assert (r.hit, r.hit_x, r.hit_y, r.steps, r.steep) == (1, 2, 4, 4, 1)
    assert r.range == pytest.approx(4.0)


def test_ns_bresenham_cell_list_by_hand():
    # from (0,0) to (5,2): error rule `2*err >= dx` with dx=5, dy=2
    # i: 0 1 2 3 4 5 ; y: 0 0 1 1 2 2  (err 2,4->y,1,3->... by hand: 2; 4>=2.5 -> y=1 err -1;
    # 1; 3 -> y=2 err -2; 0)
    want = [(0, 0), (1, 0), (2, 1), (3, 1), (4, 2), (5, 2)]
    for k, cell in enumerate(want):
        rows = [list("......") for _ in range(3)]
        rows[cell[1]][cell[0]] = "#"
        g = tiny(["".join(r) for r in rows])
        p = oracle.params("ns")
        ang = math.atan2(2, 5)
        r = oracle.map_range_f64(p, g, 0.0, 0.0, 0.0, ang, math.hypot(5, 2))
        assert (r.x1, r.y1) == (5, 2)
        assert (r.hit, r.hit_x, r.hit_y, r.steps) == (1, cell[0], cell[1], k), cell
    # a blocked cell just off the line is not hit
    g = tiny(["......", "...#..", "......"])
    g.data[1, 3] = 0
    g.data[0, 2] = 100  # (2,0) is not on the list: (2,1) is
    r = oracle.map_range_f64(oracle.params("ns"), g, 0.0, 0.0, 0.0, math.atan2(2, 5), math.hypot(5, 2))
    assert r.hit == 0


def test_start_cell_blocked_gives_zero_range():
    g = tiny(["#..", "..."])
    for preset in ("ref", "ns"):
        r = oracle.map_range_f64(oracle.params(preset), g, 0.0, 0.0, 0.3, 0.1, 5.0)
        assert (r.hit, r.steps, r.cells, r.range) == (1, 0, 1, 0.0)


def test_out_of_map_particle_is_blocked_at_start():
    g = tiny(["..."] * 3)
    r = oracle.map_range_f64(oracle.params("ns"), g, -7.0, 40.0, 0.0, 0.0, 5.0)
    assert (r.hit, r.steps, r.range) == (1, 0, 0.0)
    r = oracle.map_range_f64(oracle.params("ns"), g, float("nan"), 1.0, 0.0, 0.0, 5.0)
    assert r.hit == 1 and r.range == 0.0


def test_ref_nonsteep_is_diagonal_walk_to_origin():
    # S3: non-steep => end = start, steps (-1,-1), y steps every iteration (sensor.clj:137-138)
    g = tiny(["." * 8] * 8, res=0.5)
    p = oracle.params("ref")
    r = oracle.map_range_f64(p, g, 2.5, 1.5, 1.0, 0.05, 5.6)  # cell (5,3), bearing ~0 => non-steep
    # walks (5,3) (4,2) (3,1) (2,0) (1,-1): y=-1 is out of range => blocked
    assert (r.steep, r.hit, r.hit_x, r.hit_y, r.steps, r.cells) == (0, 1, 1, -1, 4, 5)
    assert r.range == pytest.approx(math.hypot(4, 4) * 0.5)


def test_ref_steep_backward_ray_misses_immediately():
    # S2/S4: the ray is cast -1 m along the bearing; steep with x0 > x1 (swapped) => -1.0
    g = tiny(["." * 40] * 40, res=0.05)
    p = oracle.params("ref")
    r = oracle.map_range_f64(p, g, 1.0, 1.0, 0.0, math.pi / 2, 5.6)  # bearing +90deg: end 20 cells BELOW
    assert r.steep == 1 and (r.x1, r.y1) == (20, 0)
    assert (r.hit, r.cells, r.range) == (0, 1, -1.0)


def test_ref_steep_forward_ray_tests_one_cell_past_endpoint():
    # bearing -90deg: end is 20 cells ABOVE (since R = -1), x-step +1 in the swapped frame;
    # cells y=20..41 are tested (dx+2 of them) before `(> x max-x)` fires
    g = tiny(["." * 60] * 60, res=0.05)
    p = oracle.params("ref")
    r = oracle.map_range_f64(p, g, 1.0, 1.0, 0.0, -math.pi / 2, 5.6)
    assert r.steep == 1 and (r.x1, r.y1) == (20, 40)
    assert (r.hit, r.cells, r.steps, r.range) == (0, 22, 21, -1.0)
    g.data[41, 20] = 100  # the extra cell past the endpoint is tested BEFORE the miss check
    r = oracle.map_range_f64(p, g, 1.0, 1.0, 0.0, -math.pi / 2, 5.6)
    assert (r.hit, r.hit_x, r.hit_y, r.steps) == (1, 20, 41, 21)
    assert r.range == pytest.approx(21 * float(np.float32(0.05)))


def test_ref_ignores_heading():
    g = tiny(["." * 60] * 60, res=0.05)
    p = oracle.params("ref")
    a = oracle.map_range_f64(p, g, 1.0, 1.0, 0.0, -1.2, 5.6)
    b = oracle.map_range_f64(p, g, 1.0, 1.0, 2.5, -1.2, 5.6)
    assert (a.hit, a.steps, a.range) == (b.hit, b.steps, b.range)


def test_beam_subsampling():
    # sensor.clj:144-145: n=180/360/500/1080 with max-beams 30 -> steps 6/12/17/37
    assert [oracle.lib().amclj_oracle_beam_step(n, 30) for n in (180, 360, 500, 1080)] == [6, 12, 17, 37]
    assert [oracle.lib().amclj_oracle_beams_used(n, 30) for n in (180, 360, 500, 1080)] == [30, 30, 30, 30]
    assert oracle.lib().amclj_oracle_beams_used(360, 0) == 360
    assert oracle.lib().amclj_oracle_beams_used(10, 30) == 10  # zero step clamps to 1
    assert oracle.lib().amclj_oracle_beams_used(0, 30) == 0


def test_weight_formulas_by_hand():
    # one beam, known range: NS log p and REF p^3, with the Clojure constants as the double literals they are
    # (sensor.clj:14-19: 0.95, 0.1, 0.05, 0.2; the oracle widens amclj_params' floats back to them)
    g = tiny(["........#."] * 3)
    ranges = np.array([6.5], np.float32)
    ns = oracle.params("ns")
    raw, _ = oracle.sensor_f64(ns, g, ranges, 0.0, 0.1, 20.0, [1.0], [1.0], [0.0])
    z = 6.5 - 7.0
    p = (0.95 * math.exp(-z * z / (2 * 0.2 ** 2))
         + 0.1 * 0.1 * math.exp(-0.1 * 6.5)
         + 0.05 / 20.0)
    assert raw[0] == pytest.approx(math.log(p), rel=1e-13)
    ref = oracle.params("ref", s1_use_heading=1, s2_use_scan_range_max=1, s3_proper_endpoint=1,
                        s4_endpoint_termination=1)
    raw, _ = oracle.sensor_f64(ref, g, ranges, 0.0, 0.1, 20.0, [1.0], [1.0], [0.0])
    p = (0.95 * math.exp(-z * z / (2 * 0.95 ** 2))
         + 0.1 * 0.1 * math.exp(-0.1 * 6.5)
         + 0.05 / 20.0)
    assert raw[0] == pytest.approx(p ** 3, rel=1e-13)
    # obs == rangeMax: pz3 replaces pz4
    raw, _ = oracle.sensor_f64(ns, g, np.array([20.0], np.float32), 0.0, 0.1, 20.0, [1.0], [1.0], [0.0])
    z = 20.0 - 7.0
    p = 0.95 * math.exp(-z * z / (2 * 0.2 ** 2)) + 0.05
    assert raw[0] == pytest.approx(math.log(p), rel=1e-13)


def test_sincosf_contract_accuracy():
    a = np.linspace(-40, 40, 20001).astype(np.float32)
    for v in a[::37]:
        s, c = oracle.sincosf(v)
        assert abs(float(s) - math.sin(float(v))) < 3e-7
        assert abs(float(c) - math.cos(float(v))) < 3e-7
    s, c = oracle.sincosf(0.0)
    assert (s, c) == (0.0, 1.0)


def test_fixed_weights_and_systematic_basic():
    w = np.array([0.0, 1.0, 0.0, 3.0], np.float32)
    q, total = oracle.fixed_weights(w, oracle.wmax(w))
    assert q.tolist() == [0, math.floor(2 ** 32 / 3 * (1 + 0)) if False else int(np.floor(np.ldexp(np.float32(1.0) / np.float32(3.0), 32))), 0, 2 ** 32]
    idx = oracle.systematic_resample(w, 0.5)
    assert idx.tolist() == [1, 3, 3, 3]
    assert (np.diff(idx) >= 0).all()
    # zero-weight particles are never selected, whatever u0
    for u0 in (0.0, 0.25, 0.999999):
        assert set(oracle.systematic_resample(w, u0).tolist()) <= {1, 3}


def test_systematic_counts_within_one_of_expectation():
    rng = np.random.default_rng(7)
    w = rng.random(5000).astype(np.float32) ** 4
    idx = oracle.systematic_resample(w, 0.37)
    q, total = oracle.fixed_weights(w, oracle.wmax(w))
    counts = np.bincount(idx, minlength=len(w))
    expect = q.astype(np.float64) * len(w) / total
    assert np.all(np.abs(counts - expect) < 1.0 + 1e-9)
    assert (np.diff(idx) >= 0).all() and counts.sum() == len(w)


@pytest.mark.parametrize("world", [2, 3, 8])
def test_sharded_systematic_equals_single(world):
    rng = np.random.default_rng(world)
    n = 4000
    w = (rng.random(n) ** 6).astype(np.float32)
    w[100:900] = 0  # a shard with (almost) no weight
    wm = oracle.wmax(w)
    q, total = oracle.fixed_weights(w, wm)
    u0 = 0.8125
    single = oracle.systematic_resample(w, u0)
    bounds = np.linspace(0, n, world + 1).astype(int)
    totals = np.array([int(q[a:b].sum()) for a, b in zip(bounds[:-1], bounds[1:])], np.uint64)
    r, off, m_lo, m_cnt = oracle.plan(totals, n, u0)
    assert m_cnt.sum() == n and m_lo[0] == 0
    assert (m_lo[1:] == (m_lo + m_cnt)[:-1]).all()
    out = np.full(n, -1, np.int64)
    for g in range(world):
        loc = oracle.systematic(q[bounds[g]:bounds[g + 1]], total, off[g], r, n, m_lo[g], m_cnt[g])
        assert (loc >= 0).all()
        out[m_lo[g]:m_lo[g] + m_cnt[g]] = loc + bounds[g]
    assert (out == single).all()


def test_roulette_philox_equals_materialised_stream():
    rng = np.random.default_rng(3)
    n = 200
    w = rng.random(n).astype(np.float32)
    q, total = oracle.fixed_weights(w, oracle.wmax(w))
    idx, used = oracle.roulette_philox(q, total, 6)
    sidx, su = oracle.roulette_attempts(6, 0, 400000, n)
    idx2, used2 = oracle.roulette_stream(q, total, sidx, su)
    assert (idx == idx2).all() and (idx >= 0).all()
    assert sidx.min() >= 0 and sidx.max() < n


def test_tournament_keeps_heavier_ties_second():
    w = np.array([1.0, 2.0, 2.0, 0.5], np.float32)
    sidx = np.array([0, 1, 1, 2, 2, 1, 3, 0], np.int32)
    assert oracle.tournament_stream(w, sidx).tolist() == [1, 2, 1, 0]


def test_uniform_init_lands_in_free_cells(lgfloor):
    x, y, th = oracle.uniform_init(lgfloor, 0, 2000, 1, 1)
    res = float(lgfloor.info.resolution)
    cx = (x / np.float32(res)).astype(int)
    cy = (y / np.float32(res)).astype(int)
    assert (lgfloor.data[cy, cx] == 0).mean() > 0.999  # float rounding of x*res/res at cell edges
    assert th.min() >= -math.pi - 1e-6 and th.max() <= math.pi + 1e-6
    x2, _, _ = oracle.uniform_init(lgfloor, 500, 10, 1, 1)
    assert (x2 == x[500:510]).all()  # keyed by global particle id => shard-independent


def test_f64_vs_f32_disagreement_rate_is_small(lgfloor):
    """Report (SURVEY 7 'hard parts'): how often float32 rounding moves a ray's outcome."""
    rng = np.random.Generator(np.random.PCG64(11))
    free = np.argwhere(lgfloor.data == 0)
    pick = free[rng.integers(len(free), size=300)]
    res = float(lgfloor.info.resolution)
    x = ((pick[:, 1] + rng.uniform(-0.5, 0.5, 300)) * res).astype(np.float32)
    y = ((pick[:, 0] + rng.uniform(-0.5, 0.5, 300)) * res).astype(np.float32)
    th = rng.uniform(-math.pi, math.pi, 300).astype(np.float32)
    ranges = np.full(360, 3.0, np.float32)
    p = oracle.params("ns")
    _, _, d32 = oracle.sensor_f32(p, lgfloor, ranges, -math.pi, 2 * math.pi / 360, 5.6, x, y, th, debug=True)
    amin, ainc = float(np.float32(-math.pi)), float(np.float32(2 * math.pi / 360))
    diff = 0
    for i in range(0, 300, 10):
        for k in range(0, 360, 7):
            r = oracle.map_range_f64(p, lgfloor, float(x[i]), float(y[i]), float(th[i]),
                                     amin + k * ainc, float(np.float32(5.6)))
            diff += (r.steps != d32["steps"][i, k]) or (r.hit != d32["hit"][i, k])
    assert diff / (30 * 52) < 0.02


def test_double_sincos_of_the_refinement_is_accurate():
    """contract.h sincos_det_d (restated in the oracle): fdlibm kernels after a Cody-Waite reduction."""
    for a in np.linspace(-7.0, 7.0, 4001):
        s, c = oracle.sincos_f64(float(a))
        assert abs(s - math.sin(a)) < 2.5e-16 and abs(c - math.cos(a)) < 2.5e-16


def test_fp32_contract_casts_the_rays_of_the_double_path():
    """with the start cell taken in double and end cells re-taken in double next to a .5
    boundary (csrc/contract.h CellRefine), the fp32 contract and the double restatement of the Clojure agree on
    (hit, cell, steps) of EVERY ray of these samples (round 1: 1.9e-4 of C2's rays and 9e-6 of C3's differed),
    and on every weight to 1e-5."""
    from amclj_b200 import workloads
    for w in (workloads.c2(400), workloads.c3(400)):
        s, p = w.scan, oracle.params("NS")
        raw32, _, d32 = oracle.sensor_f32(p, w.grid, s.ranges, s.angle_min, s.angle_inc, s.range_max, w.x, w.y, w.theta,
                                          debug=True)
        raw64, _ = oracle.sensor_f64(p, w.grid, s.ranges, s.angle_min, s.angle_inc, s.range_max, w.x, w.y, w.theta)
        for i in range(0, w.n, 8):
            for k in range(len(s.ranges)):
                r = oracle.map_range_f64(p, w.grid, float(w.x[i]), float(w.y[i]), float(w.theta[i]),
                                         s.angle_min + k * s.angle_inc, s.range_max)
                assert (r.hit, r.hit_x, r.hit_y, r.steps) == (d32["hit"][i, k], d32["hit_x"][i, k], d32["hit_y"][i, k],
                                                              d32["steps"][i, k])
        np.testing.assert_allclose(raw32, raw64, rtol=1e-5)


@pytest.mark.parametrize("preset", ["NS", "REF"])
def test_fp32_contract_casts_the_rays_of_the_double_path_in_bulk(preset):
    """The same statement over millions of rays (oracle.rays_f64: the double restatement ray by ray, OpenMP):
    21.6 M rays under NS, 1.8 M under REF (30 beams of -1 m, heading ignored) — no ray differs in (hit, steps),
    no hit range by more than float rounding.  The error budget in csrc/contract.h says none can; 144 M rays
    (the whole C2 cloud and 300k particles of C3) were checked once with the same code."""
    from amclj_b200 import workloads
    for w in (workloads.c2(30_000), workloads.c3(30_000)):
        s, p = w.scan, oracle.params(preset)
        _, _, d32 = oracle.sensor_f32(p, w.grid, s.ranges, s.angle_min, s.angle_inc, s.range_max, w.x, w.y, w.theta,
                                      debug=True, want_raw=False)
        steps, hit, rng = oracle.rays_f64(p, w.grid, len(s.ranges), s.angle_min, s.angle_inc, s.range_max,
                                          w.x, w.y, w.theta)
        bad = (steps != d32["steps"]) | (hit != d32["hit"])
        assert int(bad.sum()) == 0, f"{w.name}: {int(bad.sum())} of {bad.size} rays differ, first {np.argwhere(bad)[:4].tolist()}"
        assert float(np.abs(rng - d32["range_m"]).max()) < 2e-6


def test_divide_free_quotient_of_the_exact_cells_is_the_correctly_rounded_one():
    """csrc/sensor_common.cuh cell_exact_d replaces `v / res` (a subroutine call inside the look-up loop) by
    Markstein's final division step on the host's correctly rounded reciprocal: q0 = RN(v * rinv),
    r = fma(-q0, res, v) (exact), q = RN(q0 + r * rinv).  Restated here with exact rational arithmetic for the
    fused steps and checked against the IEEE quotient for the map resolutions in use and end points all over a
    map, including the ones engineered to land next to a cell boundary (where the device code takes this path)."""
    from fractions import Fraction

    def fma(a, b, c):
        return float(Fraction(a) * Fraction(b) + Fraction(c))  # one rounding, to nearest even

    rng = np.random.Generator(np.random.PCG64(77))
    for res32 in (np.float32(0.05), np.float32(0.025), np.float32(0.1), np.float32(0.0123)):
        res = float(res32)
        rinv = 1.0 / res
        vs = list(rng.uniform(-50.0, 500.0, 20000))
        cells = rng.integers(-1000, 9000, 20000)
        vs += [float(np.float32((c + 0.5) * res)) for c in cells]                      # float32 end points near a boundary
        vs += [(c + 0.5) * res * (1.0 + e) for c, e in zip(cells[:5000], rng.uniform(-3e-16, 3e-16, 5000))]
        for v in vs:
            q0 = v * rinv
            q = fma(fma(-q0, res, v), rinv, q0)
            assert q == v / res, (res, v)
