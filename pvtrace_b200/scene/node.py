"""Scene-graph node: a named coordinate system with optional geometry, light and recorders.

API of pvtrace/scene/node.py:15-205.  The tree (parent / children / traversal) is implemented here directly;
the reference delegates it to the third-party `anytree` package.
"""
from __future__ import annotations

from typing import Iterator, Sequence

import numpy as np

from pvtrace_b200.common.errors import AppError
from pvtrace_b200.geometry.intersection import Intersection
from pvtrace_b200.geometry.transformable import Transformable
from pvtrace_b200.geometry.transformations import rotation_from_matrix
from pvtrace_b200.geometry.utils import distance_between


class Node(Transformable):
    def __init__(self, name=None, parent=None, location=None, geometry=None, light=None, recorders=None):
        super(Node, self).__init__(location=location)
        self.name = name
        self._parent = None
        self._children = []
        self.parent = parent
        self.geometry = geometry
        self.light = light
        self.recorders = [] if recorders is None else list(recorders)

    def __repr__(self):
        return "Node({})".format(self.name)

    # -- tree ------------------------------------------------------------------------------

    @property
    def parent(self):
        return self._parent

    @parent.setter
    def parent(self, node):
        if node is self or (node is not None and self in node.path):
            raise AppError("A node cannot be its own ancestor.")
        if self._parent is not None:
            self._parent._children.remove(self)
        self._parent = node
        if node is not None:
            node._children.append(self)

    @property
    def children(self):
        return tuple(self._children)

    @property
    def path(self):
        """Nodes from the root down to (and including) this node."""
        chain, node = [], self
        while node is not None:
            chain.append(node)
            node = node._parent
        return tuple(reversed(chain))

    @property
    def root(self):
        return self.path[0]

    @property
    def leaves(self):
        return tuple(n for n in self.iter_preorder() if not n._children)

    def iter_preorder(self) -> Iterator["Node"]:
        yield self
        for child in self._children:
            yield from child.iter_preorder()

    def iter_postorder(self) -> Iterator["Node"]:
        for child in self._children:
            yield from child.iter_postorder()
        yield self

    def iter_levelorder(self) -> Iterator["Node"]:
        level = [self]
        while level:
            yield from level
            level = [c for n in level for c in n._children]

    def path_to(self, node) -> Sequence["Node"]:
        up, common, down = self._walk(node)
        return up + (common,) + down

    def _walk(self, node):
        a, b = self.path, node.path
        if a[0] is not b[0]:
            raise AppError("Nodes are not in the same tree.")
        k = 0
        while k < min(len(a), len(b)) and a[k] is b[k]:
            k += 1
        return tuple(reversed(a[k:])), a[k - 1], tuple(b[k:])

    # -- frames ----------------------------------------------------------------------------

    def look_at(self, vector) -> None:
        """Rotate so that the node's +z axis points along `vector` (location preserved)."""
        a = np.array([0.0, 0.0, 1.0])
        b = np.asarray(vector, dtype=float)
        c = float(a @ b)
        if np.isclose(c, -1.0):
            self.rotate(np.pi, [0, 1, 0])
            return
        v = np.cross(a, b)
        vx = np.array([[0.0, -v[2], v[1]], [v[2], 0.0, -v[0]], [-v[1], v[0], 0.0]])
        rot = np.identity(4)
        rot[:3, :3] = np.identity(3) + vx + vx @ vx / (1.0 + c)
        angle, axis, _ = rotation_from_matrix(rot)
        self.rotate(angle, axis)

    def transformation_to(self, node) -> np.ndarray:
        """4x4 matrix taking coordinates in this node's frame to `node`'s frame."""
        if self is node:
            return np.identity(4)
        up, _, down = self._walk(node)
        m = np.identity(4)
        for n in up:  # climb: local -> parent
            m = n.pose @ m
        for n in down:  # descend: parent -> local
            m = np.linalg.inv(n.pose) @ m
        return m

    def point_to_node(self, point, node) -> tuple:
        m = self.transformation_to(node)
        p = m[:3, :3] @ np.asarray(point, dtype=float) + m[:3, 3]
        return tuple(p)

    def vector_to_node(self, vector, node) -> tuple:
        m = self.transformation_to(node)
        return tuple(m[:3, :3] @ np.asarray(tuple(vector), dtype=float))

    # -- queries ---------------------------------------------------------------------------

    def intersections(self, ray_origin, ray_direction) -> Sequence[Intersection]:
        """Intersections of a ray (given in this node's frame) with this node and its subtree."""
        found = []
        if self.geometry is not None:
            for point in self.geometry.intersections(ray_origin, ray_direction):
                found.append(Intersection(coordsys=self, point=point, hit=self,
                                          distance=distance_between(ray_origin, point)))
        for child in self._children:
            found.extend(child.intersections(self.point_to_node(ray_origin, child),
                                             self.vector_to_node(ray_direction, child)))
        return tuple(found)

    def emit(self, num_rays=None):
        if self.light is None:
            raise AppError("Not a lighting node.")
        yield from self.light.emit(num_rays=num_rays)
