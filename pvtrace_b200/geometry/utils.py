"""Geometry helpers for the host-side scene classes.

Host restatements (numpy/float64) of the analytic forms the device code uses, following the reference's
compiled kernel rather than its trimesh/np.roots paths:
  ray_aabb          <- pvtrace/engine/_kernel.pyx:245-276   (slab test)
  ray_sphere        <- pvtrace/geometry/sphere.py:35-66
  ray_z_cylinder    <- pvtrace/geometry/utils.py:131-350 (results) via _kernel.pyx:301-345 (closed form)
Tolerances: EPS_ZERO = 1000 * machine epsilon (pvtrace/geometry/utils.py:12).
"""
import math

import numpy as np

EPS_ZERO = float(np.finfo(float).eps * 1000)


def close_to_zero(value) -> bool:
    return bool(np.all(np.absolute(value) < EPS_ZERO))


def points_equal(point1, point2) -> bool:
    return close_to_zero(distance_between(point1, point2))


def floats_close(a, b) -> bool:
    return close_to_zero(a - b)


def allinrange(x, x_range) -> bool:
    """True when every element of x lies in the closed interval x_range."""
    arr = np.asarray(x, dtype=float)
    return bool(np.all((arr >= x_range[0]) & (arr <= x_range[1])))


def flip(vector):
    return -np.asarray(vector, dtype=float)


def magnitude(vector) -> float:
    v = np.asarray(vector, dtype=float)
    return math.sqrt(float(v @ v))


def norm(vector):
    v = np.asarray(vector, dtype=float)
    return v / magnitude(v)


def angle_between(normal, vector) -> float:
    """Angle in [0, pi] between two unit vectors."""
    c = float(np.dot(normal, vector))
    return math.acos(max(-1.0, min(1.0, c)))


def smallest_angle_between(normal, vector) -> float:
    a = angle_between(normal, vector)
    return min(a, math.pi - a)


def distance_between(point1, point2) -> float:
    return magnitude(np.asarray(point1, dtype=float) - np.asarray(point2, dtype=float))


def is_ahead(position, direction, point) -> bool:
    offset = np.asarray(point, dtype=float) - np.asarray(position, dtype=float)
    return float(np.dot(offset, direction)) > 0.0


def intersection_point_is_ahead(ray_position, ray_direction, intersection_point) -> bool:
    """Forward test with the EPS_ZERO dead band used by Scene.intersections (geometry/utils.py:431-440)."""
    offset = np.asarray(intersection_point, dtype=float) - np.asarray(ray_position, dtype=float)
    return float(np.dot(offset, ray_direction)) > EPS_ZERO


# ---------------------------------------------------------------------------------------------
# Ray / primitive roots.  Each returns (points, distances) sorted by distance, distances >= 0.


def _points(origin, direction, ts):
    o = np.asarray(origin, dtype=float)
    d = np.asarray(direction, dtype=float)
    ts = sorted(t for t in ts if t >= 0.0)
    return tuple(tuple((o + t * d).tolist()) for t in ts), tuple(ts)


def ray_aabb(size, origin, direction):
    """Axis-aligned box centred on the origin with edge lengths `size` (slab method)."""
    o = np.asarray(origin, dtype=float)
    d = np.asarray(direction, dtype=float)
    t_near, t_far = -math.inf, math.inf
    for axis in range(3):
        lo, hi = -0.5 * size[axis], 0.5 * size[axis]
        if abs(d[axis]) < 1e-300:
            if o[axis] < lo or o[axis] > hi:
                return (), ()
            continue
        ta, tb = (lo - o[axis]) / d[axis], (hi - o[axis]) / d[axis]
        if ta > tb:
            ta, tb = tb, ta
        t_near, t_far = max(t_near, ta), min(t_far, tb)
    if t_far < t_near:
        return (), ()
    return _points(o, d, (t_near, t_far))


def ray_sphere(radius, origin, direction):
    o = np.asarray(origin, dtype=float)
    d = np.asarray(direction, dtype=float)
    a = float(d @ d)
    b = 2.0 * float(d @ o)
    c = float(o @ o) - radius * radius
    disc = b * b - 4.0 * a * c
    if disc < 0.0:
        return (), ()
    if np.isclose(disc, 0.0):
        return _points(o, d, (-b / (2.0 * a),))
    sq = math.sqrt(disc)
    return _points(o, d, ((-b - sq) / (2.0 * a), (-b + sq) / (2.0 * a)))


def ray_z_cylinder(length, radius, ray_origin, ray_direction):
    """Capped cylinder along z, centred on the origin.  Returns (points, distances)."""
    o = np.asarray(ray_origin, dtype=float)
    d = np.asarray(ray_direction, dtype=float)
    half = 0.5 * length
    roots = []
    a = d[0] * d[0] + d[1] * d[1]
    if a > 1e-300:
        b = 2.0 * (o[0] * d[0] + o[1] * d[1])
        c = o[0] * o[0] + o[1] * o[1] - radius * radius
        disc = b * b - 4.0 * a * c
        if disc >= 0.0:
            sq = math.sqrt(disc)
            for t in ((-b - sq) / (2.0 * a), (-b + sq) / (2.0 * a)):
                if -half < o[2] + t * d[2] < half:
                    roots.append(t)
    if abs(d[2]) > 1e-300:
        for cap in (-half, half):
            t = (cap - o[2]) / d[2]
            x, y = o[0] + t * d[0], o[1] + t * d[1]
            if x * x + y * y <= radius * radius:
                roots.append(t)
    # a tangent/edge ray can produce the same point twice (side root == cap root)
    unique = []
    for t in sorted(roots):
        if not unique or abs(t - unique[-1]) > EPS_ZERO:
            unique.append(t)
    return _points(o, d, unique)


def on_aabb_surface(size, point, centre=(0.0, 0.0, 0.0), atol=EPS_ZERO):
    """(is_on_surface, [face indices]) with faces ordered (xmin, xmax, ymin, ymax, zmin, zmax).

    A face is touched when the point is within atol/2 of its plane (pvtrace/geometry/utils.py:15-62).
    """
    p = np.asarray(point, dtype=float)
    c = np.asarray(centre, dtype=float)
    half = 0.5 * np.asarray(size, dtype=float)
    faces = []
    for axis in range(3):
        for side, plane in enumerate((c[axis] - half[axis], c[axis] + half[axis])):
            if abs(p[axis] - plane) < 0.5 * atol:
                faces.append(2 * axis + side)
    return len(faces) > 0, faces
