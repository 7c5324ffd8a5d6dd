"""Abstract primitive interface (pvtrace/geometry/geometry.py:8-58)."""
import abc


class Geometry(abc.ABC):
    """A solid in its own local frame; attach it to a Node to place it in a scene."""

    def __init__(self, material=None):
        self._material = material

    @property
    def material(self):
        return self._material

    @material.setter
    def material(self, new_value):
        self._material = new_value

    @abc.abstractmethod
    def is_on_surface(self, point) -> bool:
        ...

    @abc.abstractmethod
    def contains(self, point) -> bool:
        ...

    @abc.abstractmethod
    def intersections(self, origin, direction):
        """Forward intersection points, nearest first."""

    @abc.abstractmethod
    def normal(self, surface_point):
        """Outward unit normal."""

    def is_entering(self, surface_point, direction) -> bool:
        if not self.is_on_surface(surface_point):
            from pvtrace_b200.common.errors import GeometryError

            raise GeometryError("Not a surface point.", {"point": surface_point, "geometry": self})
        import numpy as np

        return float(np.dot(self.normal(surface_point), direction)) < 0.0
