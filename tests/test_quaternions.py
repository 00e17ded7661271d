"""Quaternion boundary (reference src/amclj/quaternions.clj:82-97 to-heading, :57-62 from-heading,
:19-36 rotate): the host-side helper amclj_b200/quaternions.py against the golden vectors and the
oracle — pole guards at 0.4999 / -0.499 and non-unit quaternions included — and, on a GPU, the same
conversions behind the C ABI (amclj_set_poses_quat / amclj_get_poses_quat)."""
import math

import numpy as np
import pytest

import oracle
from amclj_b200 import quaternions


def _cases():
    rng = np.random.Generator(np.random.PCG64(41))
    q = rng.standard_normal((400, 4))                       # arbitrary, mostly non-unit
    unit = q / np.linalg.norm(q, axis=1, keepdims=True)
    h = rng.uniform(-7, 7, 200)
    planar = np.stack([np.zeros(200), np.zeros(200), np.sin(h / 2), np.cos(h / 2)], 1)
    # the pole guards: test = qx*qy + qz*qw just either side of 0.4999 and -0.499
    poles = []
    for t in (0.4999, -0.499):
        for d in (-1e-9, 0.0, 1e-9, 1e-4, -1e-4):
            a = math.sqrt(abs(t + d))
            poles.append([a, math.copysign(a, t), 0.0, 0.3])     # qx*qy carries the test value
            poles.append([0.1, 0.0, a, math.copysign(a, t)])     # qz*qw carries it
    fixtures = [[0.0, 0.0, 0.21, 0.97], [0.0, 0.0, 0.85, 0.519], [0.0, 0.0, 0.0, 1.0], [0.0, 0.0, 1.0, 0.0],
                [0.0, 0.0, 0.0, 0.0]]                              # sensor.clj:69-74, core.clj:56-63 (non-unit), degenerate
    return np.concatenate([q, unit, planar, np.asarray(poles), np.asarray(fixtures)])


def test_to_heading_matches_golden(golden):
    for c in golden["quaternion"]:
        q = c["q"]
        got = float(quaternions.to_heading(q["x"], q["y"], q["z"], q["w"]))
        assert got == pytest.approx(c["to_heading"], abs=1e-15)


def test_to_heading_matches_oracle_everywhere():
    q = _cases()
    got = quaternions.to_heading(q[:, 0], q[:, 1], q[:, 2], q[:, 3])
    ref = np.array([oracle.to_heading(*row) for row in q])
    np.testing.assert_allclose(got, ref, rtol=0, atol=1e-15)
    test = q[:, 0] * q[:, 1] + q[:, 2] * q[:, 3]
    assert (test > 0.4999).any() and (test < -0.499).any() and (np.abs(test) < 0.499).any()  # all three branches ran


def test_from_heading_is_rotate_of_the_identity():
    """quaternions.clj:57-62: from-heading = (rotate (quaternion) heading), (quaternion) = (0 0 0 1)."""
    for h in np.linspace(-9.0, 9.0, 181):
        x, y, z, w = (float(v) for v in quaternions.from_heading(h))
        np.testing.assert_allclose([x, y, z, w], oracle.quat_rotate([0.0, 0.0, 0.0, 1.0], float(h)), rtol=0, atol=1e-16)
        # and back: exact up to the 2 pi wrap for a unit planar quaternion — except where the reference's
        # pole guards misfire: for a planar q, test = qz*qw = sin(h)/2, so a heading within ~0.02 rad of
        # +pi/2 (or ~0.06 rad of -pi/2) takes a "singularity" branch and comes back as 0 or +-2 pi
        # (quaternions.clj:90-93).  The boundary mirrors that; DESIGN.md section 4 lists it.
        back = float(quaternions.to_heading(x, y, z, w))
        if -0.499 <= 0.5 * math.sin(h) <= 0.4999:
            assert math.isclose(math.remainder(back - h, 2 * math.pi), 0.0, abs_tol=1e-12)
        else:
            assert math.isclose(math.remainder(back, 2 * math.pi), 0.0, abs_tol=1e-12)


@pytest.mark.gpu
def test_quaternion_poses_through_the_c_abi():
    from amclj_b200 import _lib, pf
    q = _cases()
    n = len(q)
    rng = np.random.Generator(np.random.PCG64(42))
    rows = np.concatenate([rng.uniform(0, 30, (n, 3)), q], 1)
    want = np.array([oracle.to_heading(*row) for row in q]).astype(np.float32)
    with _lib.Context(0) as c:
        c.set_poses_quat(rows)
        x, y, t, w = c.get_particles()
        assert np.array_equal(x, rows[:, 0].astype(np.float32)) and np.array_equal(y, rows[:, 1].astype(np.float32))
        ok = ~np.isnan(want)
        assert np.array_equal(t[ok], want[ok]) and np.isnan(t[~ok]).all()
        np.testing.assert_allclose(w, 1.0 / n, rtol=1e-7)
        back, _ = c.get_poses_quat()
        assert np.array_equal(back[:, 0], x.astype(np.float64)) and np.array_equal(back[:, 2], np.zeros(n))
        for k in np.flatnonzero(ok)[::7]:
            np.testing.assert_allclose(back[k, 3:], oracle.quat_rotate([0.0, 0.0, 0.0, 1.0], float(t[k])), rtol=0, atol=1e-16)
    cloud = pf.ParticleCloud(0, "NS")
    try:
        cloud.set_poses(rows[:, 0], rows[:, 1], orientation=(q[:, 0], q[:, 1], q[:, 2], q[:, 3]))
        assert np.array_equal(cloud.poses()[2][ok], want[ok])
        qx, qy, qz, qw = cloud.orientations()
        assert np.array_equal(qx, np.zeros(n)) and np.array_equal(qz[ok], np.sin(want[ok].astype(np.float64) / 2))
    finally:
        cloud.close()
