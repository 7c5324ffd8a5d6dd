"""`photon_tracer.follow` entry (pvtrace/algorithm/photon_tracer.py:276-328) on top of the GPU engine.

The reference walks one ray at a time through Python objects; here a ray is a bundle of one.  Event semantics are
those of the reference's compiled engine (which replicates `step_forward`); SURVEY.md appendix B lists where the
two reference tracers themselves differ.
"""
import numpy as np

from pvtrace_b200.engine import _cuda
from pvtrace_b200.engine.api import EngineResult
from pvtrace_b200.engine.compiler import EMIT_METHODS, compile_scene


def follow_many(scene, rays, maxsteps=1000, emit_method="kT", seed=0, max_events=None):
    """Histories [(Ray, Event), ...] of several root-frame rays traced in one device call."""
    if emit_method not in EMIT_METHODS:
        raise ValueError(f"emit_method must be one of {sorted(EMIT_METHODS)}")
    compiled = compile_scene(scene)
    positions = np.array([r.position for r in rays], dtype=np.float64).reshape(-1, 3)
    directions = np.array([r.direction for r in rays], dtype=np.float64).reshape(-1, 3)
    wavelengths = np.array([r.wavelength for r in rays], dtype=np.float64)
    budget = int(max_events or (2 * maxsteps + 4))
    data = _cuda.trace_bundle(compiled, positions, directions, wavelengths, int(seed), int(maxsteps), budget,
                              EMIT_METHODS[emit_method], 0, 1)
    result = EngineResult(compiled, data, [r.source for r in rays], budget, 1, 0.0)
    return [[(ray, event) for ray, event, _ in history] for history in result.histories()]


def follow(scene, ray, maxsteps=1000, maxpathlength=np.inf, emit_method="kT", seed=0):
    """The full history of `ray` (root frame) as a list of (Ray, Event)."""
    if np.isfinite(maxpathlength):
        raise NotImplementedError("maxpathlength is not supported by the engine (see SURVEY.md appendix B)")
    return follow_many(scene, [ray], maxsteps=maxsteps, emit_method=emit_method, seed=seed)[0]
