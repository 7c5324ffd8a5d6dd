"""Axis-aligned box centred on the local origin (API of pvtrace/geometry/box.py:22-66).

The reference Box is a trimesh `Mesh`; here it is analytic (slab intersection), which is what the reference's
compiled engine uses for boxes (pvtrace/engine/_kernel.pyx:245-276) and what the device code implements.
"""
import numpy as np

from pvtrace_b200.common.errors import GeometryError
from pvtrace_b200.geometry.geometry import Geometry
from pvtrace_b200.geometry.utils import EPS_ZERO, on_aabb_surface, ray_aabb

# outward normals in face order (xmin, xmax, ymin, ymax, zmin, zmax)
NORMALS = ((-1, 0, 0), (1, 0, 0), (0, -1, 0), (0, 1, 0), (0, 0, -1), (0, 0, 1))


class Box(Geometry):
    def __init__(self, size, material=None):
        super(Box, self).__init__(material=material)
        self._size = np.array(size, dtype=float)

    @property
    def size(self):
        return tuple(self._size.tolist())

    def is_on_surface(self, point) -> bool:
        touching, _ = on_aabb_surface(self._size, point, atol=2 * EPS_ZERO)
        inside = np.all(np.abs(np.asarray(point, dtype=float)) <= 0.5 * self._size + EPS_ZERO)
        return bool(touching and inside)

    def contains(self, point) -> bool:
        if self.is_on_surface(point):
            return False
        return bool(np.all(np.abs(np.asarray(point, dtype=float)) < 0.5 * self._size))

    def intersections(self, origin, direction):
        points, _ = ray_aabb(self._size, origin, direction)
        return points

    def normal(self, surface_point):
        touching, faces = on_aabb_surface(self._size, surface_point, atol=2 * EPS_ZERO)
        if not touching:
            raise GeometryError(
                "Point is not on surface. Is the point in the local frame?",
                {"point": surface_point, "geometry": self},
            )
        if len(faces) != 1:
            raise GeometryError("Point is on multiple surfaces.", {"point": surface_point, "geometry": self})
        return NORMALS[faces[0]]
