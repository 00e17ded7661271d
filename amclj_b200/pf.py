"""Host-side mirror of the reference's filter API (src/amclj/pf.clj, motion.clj, sensor.clj,
map.clj) over libamclj_b200.so — same function names, argument meaning and error behaviour,
so a caller of the Clojure functions finds the same calls here.  The reference is Clojure on
the JVM; no JVM exists in this image, so this mirror is Python (the production shim is the
Clojure/JNA namespace in INTEGRATION.md).  Every function below drives the CUDA library;
there is no CPU path.

Data contract:
  PoseArray     -> ``ParticleCloud``: x, y, theta (+ weight), resident on the GPU; host numpy
                   views on request.  theta = quat/to-heading(orientation).
  OccupancyGrid -> ``amclj_b200.maps.OccupancyGrid`` (data[y, x] int8, info.resolution/width/height)
  LaserScan     -> ``amclj_b200.workloads.Scan`` (ranges, angle_min, angle_inc, range_max)
"""
from __future__ import annotations

import numpy as np

from . import _lib, quaternions

NUM_PARTICLES = 1500  # pf.clj:15


class ParticleCloud:
    """The ``:poses`` of a PoseArray, kept on the device behind one ``amclj_ctx``."""

    def __init__(self, device: int = 0, preset: str = "REF"):
        self.ctx = _lib.Context(device, preset)
        self._map = None
        self._rng = np.random.default_rng()

    # -- PoseArray <-> SoA -----------------------------------------------------------------
    def set_poses(self, x, y, theta=None, orientation=None, weight=None):
        """``orientation`` = (qx, qy, qz, qw) arrays as in geometry_msgs/Pose: converted by the library
        (amclj_set_poses_quat = quat/to-heading, quaternions.clj:82-97)."""
        if theta is None:
            qx, qy, qz, qw = (np.asarray(v, np.float64) for v in orientation)
            rows = np.stack([np.asarray(x, np.float64), np.asarray(y, np.float64), np.zeros(len(qx)), qx, qy, qz, qw], 1)
            self.ctx.set_poses_quat(rows, weight)
        else:
            self.ctx.set_particles(x, y, theta, weight)
        return self

    def poses(self):
        """x, y, theta, weight as host arrays (geometry_msgs.clj:54-72 Pose fields)."""
        return self.ctx.get_particles()

    def orientations(self):
        """(qx, qy, qz, qw) of every particle: quat/from-heading inside the library (amclj_get_poses_quat)."""
        rows, _ = self.ctx.get_poses_quat()
        return rows[:, 3], rows[:, 4], rows[:, 5], rows[:, 6]

    def __len__(self):
        return self.ctx.n

    def close(self):
        self.ctx.close()


def initialize_pf(cloud: ParticleCloud, grid, num_particles=NUM_PARTICLES, pose=None, seed=1):
    """pf.clj:19-31: uniform over the map's free space without a pose (map.clj:71-84), a
    Gaussian about the pose otherwise (map.clj:108-111: sd 0.1 m, heading sd 0)."""
    cloud.ctx.set_map(grid)
    cloud._map = grid
    if pose is None:
        ref_mode = cloud.ctx.get_params().s1_use_heading == 0
        cloud.ctx.init_uniform(num_particles, 0 if ref_mode else 1, seed)
    else:
        x, y, theta = pose
        cloud.ctx.init_gaussian(num_particles, x, y, theta, 0.1, 0.0, seed)
    return cloud


def apply_motion_model(control, cloud: ParticleCloud, normals=None, seed=None):
    """pf.clj:139-145.  ``control`` = [curr prev], each (x, y, yaw) of the odom->base
    transform (motion.clj:68-79 unpacks it the same way)."""
    curr, prev = control
    cloud.ctx.motion_update(curr, prev, normals, _seed(cloud, seed))
    return cloud


def apply_sensor_model(scan, grid, cloud: ParticleCloud):
    """pf.clj:148-154: every particle gets its ``:weight`` from the measurement model."""
    if grid is not cloud._map:
        cloud.ctx.set_map(grid)
        cloud._map = grid
    cloud.ctx.sensor_update(scan.ranges, scan.angle_min, scan.angle_inc, scan.range_max)
    return cloud


def weights(cloud: ParticleCloud):
    """The ``:weight`` written by apply-sensor-model (sum p^3, or sum log p under S7)."""
    return cloud.ctx.raw_weights()


def roulette_selection(cloud: ParticleCloud, seed=None):
    """pf.clj:156-167."""
    return cloud.ctx.resample(_lib.RESAMPLE_ROULETTE, 0.0, seed=_seed(cloud, seed))


def tournament_selection(cloud: ParticleCloud, seed=None):
    """pf.clj:182-195."""
    return cloud.ctx.resample(_lib.RESAMPLE_TOURNAMENT, 0.0, seed=_seed(cloud, seed))


def systematic_selection(cloud: ParticleCloud, u0=None):
    """Low-variance resampling (north star; Probabilistic Robotics table 4.4)."""
    u0 = float(cloud._rng.random()) if u0 is None else u0
    return cloud.ctx.resample(_lib.RESAMPLE_SYSTEMATIC, u0)


def pose_estimate(cloud: ParticleCloud):
    """pf.clj:67-84: unweighted mean position and mean quaternion (z, w)."""
    mx, my, qz, qw = cloud.ctx.pose_estimate()
    return {"position": {"x": mx, "y": my}, "orientation": {"x": 0.0, "y": 0.0, "z": qz, "w": qw}}


def mcl_step(control, scan, grid, cloud: ParticleCloud, u0=None, seed=None):
    """One iteration of mcl* (pf.clj:229-234): motion -> sensor -> selection -> estimate, as
    one fused device update."""
    if grid is not cloud._map:
        cloud.ctx.set_map(grid)
        cloud._map = grid
    curr, prev = control
    u0 = float(cloud._rng.random()) if u0 is None else u0
    cloud.ctx.filter_update(curr, prev, scan.ranges, scan.angle_min, scan.angle_inc, scan.range_max, u0,
                            seed=_seed(cloud, seed))
    return pose_estimate(cloud)


def amcl_step(control, scan, grid, cloud: ParticleCloud, alpha_slow=0.01, alpha_fast=0.1, seed=None):
    """One iteration of amcl* (pf.clj:243-279): motion -> sensor -> augmented selection (median
    weight, w-slow / w-fast, random injection over free space, roulette-step otherwise) ->
    estimate.  The w-slow / w-fast atoms live in the ctx (``ctx.augment_state``).  As in the
    reference the selection works on linear weights (S7 = 0, i.e. the REF preset's sum of cubes)."""
    if grid is not cloud._map:
        cloud.ctx.set_map(grid)
        cloud._map = grid
    curr, prev = control
    cloud.ctx.motion_update(curr, prev, None, _seed(cloud, seed))
    cloud.ctx.sensor_update(scan.ranges, scan.angle_min, scan.angle_inc, scan.range_max)
    cloud.ctx.augmented_resample(alpha_slow, alpha_fast, _seed(cloud, seed))
    return pose_estimate(cloud)


def likelihood_field_range_finder_model(scan, pose, grid, cloud: ParticleCloud):
    """sensor.clj:165-173 for ONE pose (x, y, theta): returns its weight."""
    if grid is not cloud._map:
        cloud.ctx.set_map(grid)
        cloud._map = grid
    x, y, t = (np.asarray([v], np.float32) for v in pose)
    cloud.ctx.set_particles(x, y, t)
    cloud.ctx.sensor_update(scan.ranges, scan.angle_min, scan.angle_inc, scan.range_max)
    return float(cloud.ctx.raw_weights()[0])


def sample_motion_model_odometry(pose, control, cloud: ParticleCloud, normals=None, seed=None):
    """motion.clj:68-79 for ONE pose: returns the sampled (x, y, theta)."""
    x, y, t = (np.asarray([v], np.float32) for v in pose)
    cloud.ctx.set_particles(x, y, t)
    curr, prev = control
    cloud.ctx.motion_update(curr, prev, None if normals is None else np.asarray(normals, np.float32).reshape(1, 3),
                            _seed(cloud, seed))
    gx, gy, gt, _ = cloud.ctx.get_particles()
    return float(gx[0]), float(gy[0]), float(gt[0])


def _seed(cloud, seed):
    return int(cloud._rng.integers(0, 2**63)) if seed is None else int(seed)
