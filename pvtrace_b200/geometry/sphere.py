"""Sphere centred on the local origin (API of pvtrace/geometry/sphere.py:9-86)."""
import math

import numpy as np

from pvtrace_b200.geometry.geometry import Geometry
from pvtrace_b200.geometry.utils import EPS_ZERO, magnitude, ray_sphere


class Sphere(Geometry):
    def __init__(self, radius, material=None):
        super(Sphere, self).__init__(material=material)
        self.radius = radius

    def is_on_surface(self, point) -> bool:
        return abs(magnitude(point) - self.radius) < EPS_ZERO

    def contains(self, point) -> bool:
        return self.radius - (magnitude(point) + EPS_ZERO) > 0.0

    def intersections(self, origin, direction):
        points, _ = ray_sphere(self.radius, origin, direction)
        return points

    def normal(self, surface_point):
        p = np.asarray(surface_point, dtype=float)
        return tuple((p / math.sqrt(float(p @ p))).tolist())
