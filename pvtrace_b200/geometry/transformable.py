"""Pose mixin: a coordinate system that can be translated/rotated incrementally.

API of pvtrace/geometry/transformable.py:11-96 (location, pose, translate, rotate, from_pose).
"""
import numpy as np

from pvtrace_b200.geometry.transformations import (
    rotation_matrix,
    translation_from_matrix,
    translation_matrix,
)


class Transformable(object):
    def __init__(self, location=None):
        super(Transformable, self).__init__()
        origin = np.zeros(3) if location is None else np.array(location, dtype=float)
        self._pose = translation_matrix(origin)

    @classmethod
    def from_pose(cls, new_value):
        new_value = np.asarray(new_value, dtype=float)
        if new_value.shape != (4, 4):
            raise ValueError("Must be a 4x4 transform matrix")
        obj = cls()
        obj.pose = new_value
        return obj

    @property
    def pose(self):
        """4x4 matrix taking local coordinates to the parent's coordinates."""
        return self._pose

    @pose.setter
    def pose(self, new_value):
        self._pose = np.array(new_value, dtype=float)

    @property
    def location(self):
        return translation_from_matrix(self._pose)

    @location.setter
    def location(self, new_value):
        self._pose[:3, 3] = np.asarray(new_value, dtype=float)

    def translate(self, vector):
        """Relative translation, `vector` expressed in the parent's frame."""
        self._pose = translation_matrix(vector) @ self._pose
        return self

    def rotate(self, angle, axis):
        """Rotation about `axis` through the current location (location is preserved)."""
        self._pose = rotation_matrix(angle, axis, point=self.location) @ self._pose
        return self
