"""Immutable ray record (fields and methods of pvtrace/light/ray.py:15-95)."""
from __future__ import annotations

from dataclasses import dataclass, replace
from typing import Optional

speed_of_light_cm_per_s = 2.99792458e10  # lengths are centimetres


@dataclass(frozen=True)
class Ray:
    position: tuple
    direction: tuple
    wavelength: Optional[float]
    travelled: float = 0.0
    duration: float = 0.0
    source: Optional[str] = None

    def __repr__(self):
        fmt = lambda v: "(" + ", ".join(f"{x:.2f}" for x in v) + ")"  # noqa: E731
        return f"Ray(pos={fmt(self.position)}, dir={fmt(self.direction)}, nm={self.wavelength:.2f})"

    def propagate(self, distance: float, refractive_index: float) -> "Ray":
        """Ray moved `distance` along its direction; the clock advances by distance * n / c."""
        moved = tuple(p + d * distance for p, d in zip(self.position, self.direction))
        return replace(
            self,
            position=moved,
            travelled=self.travelled + distance,
            duration=self.duration + distance * refractive_index / speed_of_light_cm_per_s,
        )

    def representation(self, from_node, to_node) -> "Ray":
        """The same ray expressed in `to_node`'s coordinate system."""
        return replace(
            self,
            position=from_node.point_to_node(self.position, to_node),
            direction=from_node.vector_to_node(self.direction, to_node),
        )
