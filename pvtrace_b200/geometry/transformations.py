"""Homogeneous 4x4 transforms used by the scene graph.

The reference vendors a 1900-line transformations module but only uses four functions of it
(pvtrace/geometry/transformable.py:2-6, pvtrace/scene/node.py:9); these are written from the textbook
formulas (Rodrigues rotation, axis from the skew part) and return plain float64 arrays.
"""
import math

import numpy as np


def identity_matrix():
    return np.identity(4)


def translation_matrix(direction):
    m = np.identity(4)
    m[:3, 3] = np.asarray(direction, dtype=float)[:3]
    return m


def translation_from_matrix(matrix):
    return np.array(matrix, dtype=float)[:3, 3].copy()


def rotation_matrix(angle, direction, point=None):
    """Rotation by `angle` (radians, right handed) about the axis `direction` through `point`."""
    axis = np.asarray(direction, dtype=float)[:3]
    length = math.sqrt(float(axis @ axis))
    if length == 0.0:
        raise ValueError("rotation axis must be non-zero")
    k = axis / length
    s, c = math.sin(angle), math.cos(angle)
    cross = np.array([[0.0, -k[2], k[1]], [k[2], 0.0, -k[0]], [-k[1], k[0], 0.0]])
    rot = c * np.identity(3) + s * cross + (1.0 - c) * np.outer(k, k)
    m = np.identity(4)
    m[:3, :3] = rot
    if point is not None:
        p = np.asarray(point, dtype=float)[:3]
        m[:3, 3] = p - rot @ p
    return m


def rotation_from_matrix(matrix):
    """Inverse of `rotation_matrix`: returns (angle, axis, point-on-axis)."""
    m = np.array(matrix, dtype=float)
    rot = m[:3, :3]
    cos_a = min(1.0, max(-1.0, (np.trace(rot) - 1.0) / 2.0))
    skew = np.array([rot[2, 1] - rot[1, 2], rot[0, 2] - rot[2, 0], rot[1, 0] - rot[0, 1]])
    sin_a = 0.5 * math.sqrt(float(skew @ skew))
    angle = math.atan2(sin_a, cos_a)
    if sin_a > 1e-12:
        axis = skew / (2.0 * sin_a)
    else:
        # angle is 0 or pi: axis is the eigenvector of rot with eigenvalue 1
        w, v = np.linalg.eig(rot)
        axis = np.real(v[:, int(np.argmin(np.abs(w - 1.0)))])
        axis = axis / np.linalg.norm(axis)
    # a fixed point of the affine map: solve (I - R) p = t in the least-squares sense
    point, *_ = np.linalg.lstsq(np.identity(3) - rot, m[:3, 3], rcond=None)
    return angle, axis, np.append(point, 1.0)
