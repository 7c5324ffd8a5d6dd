"""Python-facing API of the GPU engine: `simulate`, `simulate_stream`, `EngineResult`, `RecorderResult`.

Signatures, result shapes and error behaviour follow pvtrace/engine/api.py:17-264.  Differences, all additive:
keyword-only `rng`, `device`, `devices`, `first_index`; `workers` -- OpenMP threads in the reference -- is the number of
GPUs of this process to spread the bundle over (None: one); initial rays are sampled on the device from the run's seed when every
light uses built-in delegates (the reference samples them with the unseeded global numpy RNG, emit.py:31-88);
`elapsed` is the device time of the trace call measured with CUDA events; under an initialised
`torch.distributed` process group `simulate` shards the photons over the ranks and all-reduces the tallies
(engine/distributed.py).
"""
import collections

import numpy as np

from pvtrace_b200.engine import _cuda
from pvtrace_b200.engine.compiler import EMIT_METHODS, compile_emitter, compile_scene
from pvtrace_b200.engine.emit import LightNames, emit_bundle_host
from pvtrace_b200.engine.recorder import Heatmap
from pvtrace_b200.light.event import Event
from pvtrace_b200.light.ray import Ray

# properties with always-on (sum, sum of squares) accumulators, in device order
MOMENT_PROPERTIES = ("wavelength", "angle", "duration", "pathlength")


def is_available() -> bool:
    """True when the CUDA library is built and at least one CUDA device is usable."""
    try:
        return _cuda.device_count() > 0
    except (_cuda.LibraryError, OSError, AttributeError):
        return False


class RecorderResult:
    """Tallied statistics of one recorder: `rays` (distinct rays), `crossings` (every matching interaction),
    moments and histograms per distinct ray."""

    def __init__(self, spec, rays, crossings, moments, bins):
        self.spec = spec
        self.rays = int(rays)
        self.crossings = int(crossings)
        self._moments = moments  # (4, 2)
        self._bins = bins

    def mean(self, prop):
        k = MOMENT_PROPERTIES.index(prop)
        return float("nan") if self.rays == 0 else self._moments[k, 0] / self.rays

    def std(self, prop):
        k = MOMENT_PROPERTIES.index(prop)
        if self.rays == 0:
            return float("nan")
        mean = self._moments[k, 0] / self.rays
        return float(np.sqrt(max(self._moments[k, 1] / self.rays - mean * mean, 0.0)))

    def error(self, prop):
        return float("nan") if self.rays == 0 else self.std(prop) / np.sqrt(self.rays)

    def histogram(self, index=0):
        """(edges, counts) for a Histogram, (edges_a, edges_b, counts[na, nb]) for a Heatmap."""
        spec, values = self.spec.histograms[index], self._bins[index]
        if isinstance(spec, Heatmap):
            return (np.linspace(spec.a.start, spec.a.stop, spec.a.bins + 1),
                    np.linspace(spec.b.start, spec.b.stop, spec.b.bins + 1),
                    values.reshape(spec.a.bins, spec.b.bins))
        return np.linspace(spec.start, spec.stop, spec.bins + 1), values

    def __repr__(self):
        return f"RecorderResult({self.spec.name!r}, rays={self.rays}, crossings={self.crossings})"


class EngineResult:
    """Results of one traced bundle.  `data` is the dict the kernel returns (pvtrace/engine/_kernel.pyx:1097-1115):
    tallies over every ray, and an event log for every `record_every`-th ray where event k of recorded ray j is
    row `j * max_events + k`."""

    def __init__(self, compiled, data, sources, max_events, record_every, elapsed, num_rays=None, first_index=0):
        self.compiled = compiled
        self.data = data
        self.sources = sources  # light name per ray of THIS process's slice (all rays when not sharded)
        self.max_events = max_events
        self.record_every = record_every
        self.elapsed = elapsed
        self.first_index = int(first_index)  # global index of the first ray of the slice (numbering of sinks)
        self._num_rays = num_rays

    @property
    def num_rays(self):
        """Rays the tallies refer to: under a process group the GLOBAL count (the tallies are all-reduced)."""
        return len(self.sources) if self._num_rays is None else int(self._num_rays)

    @property
    def num_recorded(self):
        return len(self.data["counts"])

    @property
    def recorded_indices(self):
        if self.record_every <= 0:
            return np.zeros(0, dtype=np.int64)
        return np.arange(0, len(self.sources), self.record_every, dtype=np.int64)

    @property
    def stats(self):
        """Device counters of the run: photon steps, rays retired, kernel launches, events generated."""
        s = self.data.get("stats")
        if s is None:
            return {}
        return {"steps": int(s[_cuda.STAT_STEPS]), "rays": int(s[_cuda.STAT_RAYS]),
                "launches": int(s[_cuda.STAT_LAUNCHES]), "events": int(s[_cuda.STAT_EVENTS])}

    @property
    def recorders(self):
        c, d = self.compiled, self.data
        found = {}
        for r, spec in enumerate(c.recorder_specs):
            start = int(c.rec_hist_start[r])
            bins = []
            for h in range(len(spec.histograms)):
                offset = int(c.hist_offset[start + h])
                bins.append(d["rec_bins"][offset:offset + int(c.hist_na[start + h]) * int(c.hist_nb[start + h])])
            found[spec.name] = RecorderResult(spec, d["rec_distinct"][r], d["rec_crossings"][r], d["rec_sums"][r], bins)
        return found

    def event_counts(self):
        """Counter of LOGGED events by Event member (recorded rays only)."""
        counts = self.data["counts"]
        if len(counts) == 0:
            return collections.Counter()
        kinds = self.data["kind"].reshape(self.num_recorded, self.max_events)
        logged = kinds[np.arange(self.max_events)[None, :] < counts[:, None]]
        values, tallies = np.unique(logged, return_counts=True)
        return collections.Counter({Event(int(v)): int(t) for v, t in zip(values, tallies)})

    def _node_name(self, index):
        return self.compiled.node_names[index] if index >= 0 else None

    def _component_name(self, index):
        return self.compiled.component_names[index] if index >= 0 else None

    def histories(self):
        """One history per recorded ray: [(Ray, Event, metadata), ...]."""
        d, indices = self.data, self.recorded_indices
        for j in range(self.num_recorded):
            history = []
            for row in range(j * self.max_events, j * self.max_events + int(d["counts"][j])):
                src = int(d["source"][row])
                ray = Ray(position=tuple(d["position"][row].tolist()), direction=tuple(d["direction"][row].tolist()),
                          wavelength=float(d["wavelength"][row]), travelled=float(d["travelled"][row]),
                          duration=float(d["duration"][row]),
                          source=self.sources[int(indices[j])] if src < 0 else self._component_name(src))
                event = Event(int(d["kind"][row]))
                meta = {"hit": self._node_name(int(d["hit"][row])),
                        "container": self._node_name(int(d["container"][row])),
                        "adjacent": self._node_name(int(d["adjacent"][row])),
                        "component": self._component_name(int(d["component"][row]))}
                if event in (Event.REFLECT, Event.TRANSMIT):
                    meta["normal"] = tuple(d["normal"][row].tolist())
                history.append((ray, event, meta))
            yield history


def _to_sqlite(self, dbfilepath, first_throw_id=None, end_rays=False):
    """Write the logged histories into the reference CLI's database layout (engine/sinks.py)."""
    from pvtrace_b200.engine import sinks

    return sinks.write_sqlite(self, dbfilepath, first_throw_id=first_throw_id, end_rays=end_rays)


EngineResult.to_sqlite = _to_sqlite


def simulate(scene, num_rays, seed=None, workers=None, maxsteps=1000, max_events=128, emit_method="kT",
             record_every=1, *, rng="philox", device=None, devices=None, first_index=0):
    """Trace `num_rays` through `scene` on the GPU.

    Recorders attached to scene nodes tally every ray; full event histories are kept for every
    `record_every`-th ray (all when 1, none when 0).  Raises `UnsupportedSceneError` if the scene cannot be
    flattened and `ValueError` for a bad `emit_method`.  `workers` (threads of the reference's CPU engine) is the
    number of GPUs to use: the first `workers` devices, or pass `devices=[...]` explicitly; the result does not depend
    on it (ray i is a function of (seed, i, scene) alone).
    """
    if emit_method not in EMIT_METHODS:
        raise ValueError(f"emit_method must be one of {sorted(EMIT_METHODS)}")
    if rng not in _cuda.RNG_MODES:
        raise ValueError(f"rng must be one of {sorted(_cuda.RNG_MODES)}")
    compiled = compile_scene(scene)
    if seed is None:
        seed = np.random.randint(0, 2 ** 31 - 1)

    from pvtrace_b200.engine import distributed

    if distributed.is_active():
        return distributed.simulate_sharded(scene, compiled, num_rays, int(seed), maxsteps, max_events, emit_method,
                                            record_every, rng=rng, first_index=first_index)

    if devices is None and workers is not None and int(workers) > 1 and device is None:
        devices = list(range(min(int(workers), max(_cuda.device_count(), 1))))
    emitter = compile_emitter(scene)
    if emitter is not None:
        positions = directions = wavelengths = None
        sources = LightNames(emitter.light_names, num_rays, first_index)
    else:
        positions, directions, wavelengths, sources = emit_bundle_host(scene, num_rays)
    data, elapsed = _cuda.trace_bundle(
        compiled, positions, directions, wavelengths, int(seed), int(maxsteps), int(max_events),
        EMIT_METHODS[emit_method], 0, int(record_every), emitter=emitter, n=int(num_rays),
        first_index=int(first_index), rng_mode=_cuda.RNG_MODES[rng], device=int(device or 0), return_elapsed=True,
        devices=devices)
    return EngineResult(compiled, data, sources, max_events, record_every, elapsed, first_index=first_index)


def simulate_stream(scene, num_rays, bundle=50000, seed=None, **kwargs):
    """Trace in bundles, yielding (EngineResult, rays_traced_so_far).  Ray i of the whole run always uses stream
    `seed + i`, so the union of the bundles equals one `simulate` call with the same seed (api.py:249-264);
    accumulate tallies by summing the `rec_*` arrays."""
    if seed is None:
        seed = np.random.randint(0, 2 ** 31 - 1)
    traced = 0
    while traced < num_rays:
        n = min(bundle, num_rays - traced)
        result = simulate(scene, n, seed=int(seed), first_index=traced, **kwargs)
        traced += n
        yield result, traced
