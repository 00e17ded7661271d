"""Boundary conversions between the reference's quaternion poses and the theta the kernels
store (reference src/amclj/quaternions.clj).  Host-side numpy, not on the timed path."""
from __future__ import annotations

import numpy as np


def to_heading(qx, qy, qz, qw):
    """quaternions.clj:82-97 ``to-heading`` (vectorised), pole guards included."""
    qx, qy, qz, qw = (np.asarray(v, np.float64) for v in (qx, qy, qz, qw))
    test = qx * qy + qz * qw
    out = np.arctan2(2 * qz * qw - 2 * qx * qy, 1 - 2 * qz * qz - 2 * qy * qy)
    out = np.where(test > 0.4999, 2 * np.arctan2(qx, qw), out)
    out = np.where(test < -0.499, -2 * np.arctan2(qx, qw), out)
    return out


def from_heading(theta):
    """quaternions.clj:57-62 ``from-heading``: (x, y, z, w) = (0, 0, sin(t/2), cos(t/2))."""
    theta = np.asarray(theta, np.float64)
    z = np.zeros_like(theta)
    return z, z, np.sin(theta / 2), np.cos(theta / 2)
