"""Scene: a tree of nodes plus ray emission and tracing entry points.

Public surface of pvtrace/scene/scene.py:94-313 (root, light_nodes, component_nodes, emit, intersections,
simulate).  `simulate` runs on the GPU engine; the reference's multiprocessing pool over the Python tracer
(scene.py:266-313) is superseded and only its signature is kept.
"""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np

from pvtrace_b200.geometry.utils import intersection_point_is_ahead
from pvtrace_b200.light.light import Light
from pvtrace_b200.scene.node import Node


class Scene(object):
    def __init__(self, root: Optional[Node] = None):
        self.root = root

    def finalise_nodes(self):
        """Kept for API compatibility; the flattener recomputes transforms on every compile."""

    @property
    def light_nodes(self) -> Sequence[Node]:
        return [n for n in self.root.iter_levelorder() if isinstance(n.light, Light)]

    @property
    def component_nodes(self):
        found = []
        for n in self.root.iter_levelorder():
            if n.geometry is not None and n.geometry.material is not None:
                found.extend(n.geometry.material.components)
        return found

    def emit(self, num_rays):
        """Rays in the root frame; lights take turns (ray i comes from light i % n_lights)."""
        lights = self.light_nodes
        for idx in range(num_rays):
            node = lights[idx % len(lights)]
            for ray in node.emit(1):
                yield ray.representation(node, self.root)

    def intersections(self, ray_origin, ray_direction):
        """All forward intersections of a root-frame ray with the scene, nearest first."""
        if self.root is None:
            return tuple()
        origin = np.asarray(ray_origin, dtype=float)
        hits = (i.to(self.root) for i in self.root.intersections(ray_origin, ray_direction))
        ahead = [i for i in hits if intersection_point_is_ahead(ray_origin, ray_direction, i.point)]
        ahead.sort(key=lambda i: float(np.linalg.norm(np.asarray(i.point) - origin)))
        return tuple(ahead)

    def simulate(self, num_rays: int, workers: Optional[int] = None, seed: Optional[int] = None,
                 queue=None, end_rays: bool = False, **engine_kwargs):
        """Trace `num_rays` on the GPU and return one history per ray: [(Ray, Event), ...].

        `workers`, `queue` and `end_rays` belong to the reference's CPU pool and are accepted but unused.
        """
        from pvtrace_b200 import engine

        engine_kwargs.setdefault("record_every", 1)
        result = engine.simulate(self, num_rays, seed=seed, **engine_kwargs)
        return [[(ray, event) for ray, event, _ in history] for history in result.histories()]
