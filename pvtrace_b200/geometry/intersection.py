"""Intersection record (fields of pvtrace/geometry/intersection.py:9-43)."""
from dataclasses import dataclass, replace
from typing import Any


@dataclass(frozen=True)
class Intersection:
    coordsys: Any  # node whose frame `point` is expressed in
    point: tuple
    hit: Any       # node that owns the surface
    distance: float

    def to(self, node):
        """The same intersection expressed in `node`'s frame."""
        return replace(self, coordsys=node, point=self.coordsys.point_to_node(self.point, node))
