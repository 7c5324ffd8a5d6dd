"""Initial rays for the engine (role of pvtrace/engine/emit.py:92-134).

Lights built from recognised delegates are sampled ON THE DEVICE from a seeded Philox stream
(`compile_emitter` -> pvt_emit_t -> emit_ray in csrc/pvt_photon.cuh); `emit_bundle` returns those rays as
host arrays through `pvt_emit_bundle`.  Scenes with a custom Python delegate fall back to calling the delegates
once per ray on the host (`emit_bundle_host`), which keeps every light working at Python speed.
"""
import ctypes as C

import numpy as np

from pvtrace_b200.engine import _cuda
from pvtrace_b200.engine.compiler import compile_emitter


class LightNames:
    """Read-only sequence: name of the light that emitted ray i (lights take turns, scene/scene.py:141-151)."""

    def __init__(self, names, count, first_index=0):
        self._names, self._count, self._first = list(names), int(count), int(first_index)

    def __len__(self):
        return self._count

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[k] for k in range(*i.indices(self._count))]
        if i < 0:
            i += self._count
        if not 0 <= i < self._count:
            raise IndexError(i)
        return self._names[(self._first + i) % len(self._names)]

    def __iter__(self):
        return (self[i] for i in range(self._count))

    def tolist(self):
        return list(self)


def emit_bundle_host(scene, num_rays):
    """Per-ray host emission through the light delegates (global numpy RNG), world frame."""
    positions = np.zeros((num_rays, 3))
    directions = np.zeros((num_rays, 3))
    wavelengths = np.zeros(num_rays)
    sources = []
    for row, ray in enumerate(scene.emit(num_rays)):
        positions[row] = ray.position
        directions[row] = ray.direction
        wavelengths[row] = ray.wavelength
        sources.append(ray.source)
    return positions, directions, wavelengths, sources


def emit_bundle(scene, num_rays, seed=0, first_index=0, device=0):
    """(positions[N,3], directions[N,3], wavelengths[N], sources) in the world frame."""
    emitter = compile_emitter(scene)
    if emitter is None:
        return emit_bundle_host(scene, num_rays)
    lib = _cuda.load_library()
    positions = np.zeros((num_rays, 3))
    directions = np.zeros((num_rays, 3))
    wavelengths = np.zeros(num_rays)
    struct, keep = _cuda.marshal_emitter(emitter)
    status = lib.pvt_emit_bundle(C.byref(struct), _cuda._vp(positions), _cuda._vp(directions), _cuda._vp(wavelengths),
                                 int(num_rays), int(first_index), int(seed) & 0xFFFFFFFFFFFFFFFF, int(device))
    _cuda.check(status, "pvt_emit_bundle")
    del keep
    return positions, directions, wavelengths, LightNames(emitter.light_names, num_rays, first_index)
