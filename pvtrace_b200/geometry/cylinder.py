"""Capped cylinder along local z (API of pvtrace/geometry/cylinder.py:9-77)."""
import math

import numpy as np

from pvtrace_b200.common.errors import GeometryError
from pvtrace_b200.geometry.geometry import Geometry
from pvtrace_b200.geometry.utils import close_to_zero, norm, ray_z_cylinder


class Cylinder(Geometry):
    def __init__(self, length, radius, material=None):
        super(Cylinder, self).__init__(material=material)
        self.length = length
        self.radius = radius

    def is_on_surface(self, point) -> bool:
        # probe with an arbitrary ray: a surface point is its own first intersection
        _, dist = ray_z_cylinder(self.length, self.radius, point, norm((1, 1, 1)))
        return len(dist) > 0 and close_to_zero(dist[0])

    def contains(self, point) -> bool:
        r = math.hypot(point[0], point[1])
        return -0.5 * self.length < point[2] < 0.5 * self.length and r < self.radius

    def intersections(self, origin, direction):
        points, _ = ray_z_cylinder(self.length, self.radius, origin, direction)
        return points

    def normal(self, surface_point):
        z = surface_point[2]
        if np.isclose(z, -0.5 * self.length):
            return (0.0, 0.0, -1.0)
        if np.isclose(z, 0.5 * self.length):
            return (0.0, 0.0, 1.0)
        r = math.hypot(surface_point[0], surface_point[1])
        if np.isclose(self.radius, r):
            return (surface_point[0] / r, surface_point[1] / r, 0.0)
        raise GeometryError("Not a surface point.", {"point": surface_point, "geometry": self})
