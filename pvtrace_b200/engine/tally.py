"""Host-side recorder tallies over (ray, event, metadata) histories.

Defines what a recorder counts independently of the device code (role of pvtrace/engine/tally.py:26-47,86-156):
tests recompute every tally from the engine's own event log with this function and demand exact agreement.
"""
import math

import numpy as np

from pvtrace_b200.engine.recorder import Heatmap
from pvtrace_b200.light.event import Event

_MOMENTS = ("wavelength", "angle", "duration", "pathlength")


def _selector_matches(kind, name, event, meta):
    if event == Event.TRANSMIT:
        if meta.get("hit") != name:
            return False
        if kind == "entering":
            return meta.get("adjacent") == name
        if kind == "escaping":
            return meta.get("container") == name
        return False
    if event == Event.REFLECT:
        return kind == "reflected" and meta.get("hit") == name and meta.get("adjacent") == name
    if event == Event.NONRADIATIVE:
        return kind == "lost" and meta.get("container") == name
    if event == Event.REACT:
        return kind == "reacted" and meta.get("container") == name
    if event == Event.KILL:
        return kind == "killed" and meta.get("container") == name
    if event == Event.EXIT:
        return kind == "exit" and meta.get("hit") == name
    return False


def _bin(value, h):
    index = int((value - h.start) / (h.stop - h.start) * h.bins)
    return index if 0 <= index < h.bins else -1


class _Running:
    def __init__(self, recorder):
        self.recorder = recorder
        self.rays = 0
        self.crossings = 0
        self.moments = np.zeros((4, 2))
        self.bins = [np.zeros(h.a.bins * h.b.bins if isinstance(h, Heatmap) else h.bins, dtype=np.int64)
                     for h in recorder.histograms]

    def add(self, values):
        self.rays += 1
        for k, prop in enumerate(_MOMENTS):
            self.moments[k, 0] += values[prop]
            self.moments[k, 1] += values[prop] * values[prop]
        for h, bins in zip(self.recorder.histograms, self.bins):
            if isinstance(h, Heatmap):
                ia, ib = _bin(values[h.a.prop], h.a), _bin(values[h.b.prop], h.b)
                if ia >= 0 and ib >= 0:
                    bins[ia * h.b.bins + ib] += 1
            else:
                i = _bin(values[h.prop], h)
                if i >= 0:
                    bins[i] += 1


def tally_histories(scene, histories):
    """dict recorder name -> RecorderResult, from one history per ray."""
    from pvtrace_b200.engine.api import RecorderResult

    root = scene.root
    slots = [(node, rec, _Running(rec)) for node in root.iter_preorder() for rec in getattr(node, "recorders", [])]

    for history in histories:
        seen = set()
        before = None
        for ray, event, meta in history:
            meta = meta or {}
            for node, rec, state in slots:
                if not _selector_matches(rec.event, node.name, event, meta):
                    continue
                normal = meta.get("normal")
                local = tuple(ray.position) if node is root else root.point_to_node(ray.position, node)
                if event == Event.EXIT and normal is None:
                    normal = node.vector_to_node(node.geometry.normal(local), root)
                if rec.facet is not None:
                    if normal is None or any(abs(f - n) > rec.atol for f, n in zip(rec.facet, normal)):
                        continue
                state.crossings += 1
                if rec.name in seen:
                    continue
                seen.add(rec.name)
                incident = ray.direction if event == Event.EXIT else (before or ray).direction
                angle = 0.0
                if normal is not None:
                    angle = math.acos(min(abs(float(np.dot(incident, normal))), 1.0))
                state.add({"wavelength": ray.wavelength, "angle": angle, "duration": ray.duration,
                           "pathlength": ray.travelled, "x": local[0], "y": local[1], "z": local[2]})
            before = ray

    return {rec.name: RecorderResult(rec, st.rays, st.crossings, st.moments, st.bins) for _, rec, st in slots}
