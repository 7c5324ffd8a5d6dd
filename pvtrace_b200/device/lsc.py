"""High-level luminescent-solar-concentrator factory (constructor and `add_*` methods of
pvtrace/device/lsc.py:89-336).

The reference attaches Python subclasses of `FresnelSurfaceDelegate` to the LSC node (`OptionalMirrorAndSolarCell`,
`AirGapMirror`, lsc.py:22-86), which its compiled engine rejects (compiler.py:239-247).  Here the same options are
lowered to DATA -- a `FacetSurfaceDelegate` facet table -- so the device traces them:

  back-surface mirror   facet (0,0,-1): reflectivity 1                               (lsc.py:36-37)
  edge solar cells      facet (+-1,0,0)/(0,+-1,0): reflectivity 0, straight transmit (lsc.py:38-62)
  air-gap mirror        separate thin n0 box under the LSC, every facet reflectivity 1, specular or Lambertian
                        (lsc.py:65-86; the reference's "specular" branch returns the *refracted* direction,
                        a bug we do not reproduce)

`simulate` runs the GPU engine with per-face recorders instead of per-ray pandas rows (lsc.py:338-454).
"""
import functools

import numpy as np

from pvtrace_b200.data import lumogen_f_red_305
from pvtrace_b200.engine.recorder import Histogram, Recorder
from pvtrace_b200.geometry.box import Box
from pvtrace_b200.light.light import Light
from pvtrace_b200.material.component import Absorber, Luminophore, Scatterer
from pvtrace_b200.material.material import Material
from pvtrace_b200.material.surface import Facet, FacetSurfaceDelegate, Surface
from pvtrace_b200.material.utils import cone
from pvtrace_b200.scene.node import Node
from pvtrace_b200.scene.scene import Scene

FACES = {"left": (-1, 0, 0), "right": (1, 0, 0), "near": (0, -1, 0), "far": (0, 1, 0), "top": (0, 0, 1),
         "bottom": (0, 0, -1)}


class LSC(object):
    def __init__(self, size, wavelength_range=None, n0=1.0, n1=1.5):
        self.wavelength_range = np.arange(400, 800) if wavelength_range is None else wavelength_range
        self.size = size  # centimetres
        self.n0, self.n1 = n0, n1
        self._solar_cell_surfaces = set()
        self._back_surface_mirror_info = {"want_back_surface_mirror": False}
        self._air_gap_mirror_info = {"want_air_gap_mirror": False, "lambertian": False}
        self._scene = None
        self._result = None
        self._store = None  # end-ray rows of every simulate() call so far (the reference appends too)
        self._user_lights = []
        self._user_components = []

    # -- defaults (lsc.py:115-146) -----------------------------------------------------------

    def _make_default_components(self):
        x = self.wavelength_range
        return [
            {"cls": Luminophore, "name": "Lumogen F Red 305",
             "coefficient": np.column_stack((x, lumogen_f_red_305.absorption(x) * 10.0)),  # 10 cm-1 at peak
             "emission": np.column_stack((x, lumogen_f_red_305.emission(x))),
             "quantum_yield": 1.0, "phase_function": None},
            {"cls": Absorber, "coefficient": 0.1, "name": "Background"},
        ]

    def _make_default_lights(self):
        return [{"name": "Light", "location": (0.0, 0.0, self.size[-1] * 5), "rotation": (np.radians(180), (1, 0, 0)),
                 "direction": functools.partial(cone, np.radians(20)), "wavelength": None, "position": None}]

    # -- configuration -------------------------------------------------------------------------

    def add_luminophore(self, name, coefficient, emission, quantum_yield, phase_function=None):
        self._user_components.append({"cls": Luminophore, "name": name, "coefficient": coefficient,
                                      "emission": emission, "quantum_yield": quantum_yield,
                                      "phase_function": phase_function})

    def add_absorber(self, name, coefficient):
        self._user_components.append({"cls": Absorber, "name": name, "coefficient": coefficient})

    def add_scatterer(self, name, coefficient, phase_function=None):
        self._user_components.append({"cls": Scatterer, "name": name, "coefficient": coefficient,
                                      "phase_function": phase_function})

    def add_light(self, name, location, rotation=None, direction=None, wavelength=None, position=None):
        self._user_lights.append({"name": name, "location": location, "rotation": rotation, "direction": direction,
                                  "wavelength": wavelength, "position": position})

    def add_solar_cell(self, facets):
        if not isinstance(facets, (list, tuple, set)):
            raise ValueError("Facets should be a set. e.g. `{'left', 'right'}`")
        facets = set(facets)
        allowed = {"left", "near", "far", "right"}
        if not facets.issubset(allowed):
            raise ValueError("Solar cell have allowed surfaces", allowed)
        self._solar_cell_surfaces = facets.union(self._solar_cell_surfaces)

    def add_back_surface_mirror(self):
        self._back_surface_mirror_info = {"want_back_surface_mirror": True}

    def add_air_gap_mirror(self, lambertian=False):
        self._air_gap_mirror_info = {"want_air_gap_mirror": True, "lambertian": lambertian}

    def component_names(self):
        if self._scene is None:
            raise ValueError("Run a simulation before calling this method.")
        return {c["name"] for c in self._user_components}

    def light_names(self):
        if self._scene is None:
            raise ValueError("Run a simulation before calling this method.")
        return {light["name"] for light in self._user_lights}

    # -- scene (lsc.py:148-219) --------------------------------------------------------------------

    def _lsc_facets(self):
        facets = []
        if self._back_surface_mirror_info["want_back_surface_mirror"]:
            facets.append(Facet(FACES["bottom"], reflectivity=1.0, atol=1e-8))
        for name in ("left", "right", "near", "far"):
            if name in self._solar_cell_surfaces:
                facets.append(Facet(FACES[name], reflectivity=0.0, transmit="straight", atol=1e-8))
        return facets

    def _make_scene(self, record=True):
        (l, w, d) = self.size
        world = Node(name="World", geometry=Box((l * 100, w * 100, d * 100), material=Material(refractive_index=self.n0)))
        if len(self._user_components) == 0:
            self._user_components = self._make_default_components()
        components = []
        for spec in self._user_components:
            spec = dict(spec)
            cls, coefficient = spec.pop("cls"), spec.pop("coefficient")
            components.append(cls(coefficient, **spec))
        lsc = Node(name="LSC", parent=world, geometry=Box((l, w, d), material=Material(
            refractive_index=self.n1, components=components,
            surface=Surface(delegate=FacetSurfaceDelegate(self._lsc_facets())))))
        if self._air_gap_mirror_info["want_air_gap_mirror"]:
            thickness = 0.25 * d
            mode = "lambertian" if self._air_gap_mirror_info["lambertian"] else "specular"
            mirror = Node(name="Air Gap Mirror", parent=world, geometry=Box((l, w, thickness), material=Material(
                refractive_index=self.n0, components=[],
                surface=Surface(delegate=FacetSurfaceDelegate(
                    [Facet(n, reflectivity=1.0, reflect=mode, atol=1e-8) for n in FACES.values()])))))
            mirror.translate((0.0, 0.0, -(0.5 * d + thickness)))
        if len(self._user_lights) == 0:
            self._user_lights = self._make_default_lights()
        for spec in self._user_lights:
            node = Node(name=spec["name"], parent=world, light=Light(
                name=spec["name"], direction=spec["direction"], wavelength=spec["wavelength"],
                position=spec["position"]))
            node.location = spec["location"]
            if spec["rotation"]:
                node.rotate(*spec["rotation"])
        if record:
            wl = (300.0, 1000.0, 100)
            for face, normal in FACES.items():
                lsc.recorders.append(Recorder(f"escaping-{face}", event="escaping", facet=normal,
                                              histograms=[Histogram("wavelength", *wl)]))
                lsc.recorders.append(Recorder(f"reflected-{face}", event="reflected", facet=normal))
                lsc.recorders.append(Recorder(f"entering-{face}", event="entering", facet=normal))
            lsc.recorders.append(Recorder("lost", event="lost", histograms=[Histogram("wavelength", *wl)]))
            world.recorders.append(Recorder("exit", event="exit"))
            world.recorders.append(Recorder("killed", event="killed"))
        self._scene = Scene(world)
        return self._scene

    # -- run + summary -----------------------------------------------------------------------------

    def simulate(self, n, progress=None, emit_method="kT", seed=None, record_every=None, **engine_kwargs):
        """Trace `n` rays on the GPU; returns the EngineResult.

        Like the reference's (lsc.py:338-377) the call keeps what `report()` / `spectrum()` / `counts()` need -- the
        entrance and exit rows of the logged rays -- and a second call APPENDS to that store.  The reference logs every
        ray; here `record_every` defaults to every ray up to 10^5 rays and to a sample of about 10^5 histories beyond
        (a full log of 10^7 rays would be 150 GB); the recorder tallies (`recorder_counts()`) always cover every ray."""
        from pvtrace_b200 import engine

        if self._scene is None:
            self._make_scene()
        if record_every is None:
            record_every = max(1, int(n) // 100_000)
        self._result = engine.simulate(self._scene, n, seed=seed, emit_method=emit_method,
                                       record_every=record_every, **engine_kwargs)
        if self._result.num_recorded:
            rows = self._rows_of(self._result)
            self._store = rows if self._store is None else {k: np.concatenate([self._store[k], rows[k]]) for k in rows}
        if progress:
            progress(int(n))
        return self._result

    def recorder_counts(self):
        """Distinct-ray counts of the LAST run's recorders, over every ray (not only the logged ones):
        {'escaping': {face: n}, 'reflected': {...}, 'entering': {...}, 'lost': n, 'exit': n, 'killed': n, 'thrown': n}."""
        if self._result is None:
            raise ValueError("Run a simulation before calling this method.")
        rec = self._result.recorders
        out = {kind: {face: rec[f"{kind}-{face}"].rays for face in FACES} for kind in ("escaping", "reflected", "entering")}
        out.update(lost=rec["lost"].rays, exit=rec["exit"].rays, killed=rec["killed"].rays,
                   thrown=self._result.num_rays)
        return out

    def counts(self):
        """The reference's surface-count table (lsc.py:455-506): {'Solar In' | 'Solar Out' | 'Luminescent Out' |
        'Luminescent In': {facet: count}} over the stored end rays of every `simulate` call so far."""
        return self.counts_table()

    # -- report (lsc.py:379-607 rebuilt on the engine's event log instead of per-ray pandas rows) ----------------

    def _end_rays(self):
        """Entrance and exit rows of every logged ray, as the reference stores them (lsc.py:338-364): the entrance is
        the first event after GENERATE; the exit is the last event of a ray that was lost or killed, the event before
        EXIT otherwise (where the ray left the LSC region).  Returns a dict of equally long arrays:
        kind ('entrance' | 'exit'), event (lower-case name), facet (name or None), luminescent (bool: the ray's
        source is a component, not a light), wavelength.

        Two deliberate differences from the reference's code, which is older than its own tracer: rays whose last
        event is NONRADIATIVE or REACT count as lost (the reference only tests for ABSORB, which is never last any
        more), and facets are named by their outward normal as the solar-cell delegate names them (lsc.py:38-47:
        near = -y, far = +y; `label_facets`, lsc.py:447-452, has the two swapped)."""
        if self._result is None:
            raise ValueError("Run a simulation before calling this method.")
        if self._store is None:
            raise ValueError("The report is built from logged histories: simulate(..., record_every=k) with k >= 1.")
        return self._store

    def _rows_of(self, result):
        """The end-ray rows of one run (see `_end_rays`); `source` holds the NAME of the light or component."""
        from pvtrace_b200.light.event import Event

        d, m = result.data, result.max_events
        counts = np.asarray(d["counts"], dtype=np.int64)
        base = np.arange(len(counts), dtype=np.int64) * m
        has_entrance = counts > 1
        last = base + counts - 1
        last_kind = d["kind"][last]
        lost = np.isin(last_kind, [Event.ABSORB.value, Event.NONRADIATIVE.value, Event.REACT.value, Event.KILL.value])
        left = (last_kind == Event.EXIT.value) & (counts > 1)
        rows = np.concatenate([(base + 1)[has_entrance], last[lost], (last - 1)[left]])
        kinds = np.array(["entrance"] * int(has_entrance.sum()) + ["exit"] * int(lost.sum() + left.sum()))
        position = d["position"][rows]
        half = 0.5 * np.asarray(self.size, dtype=float)
        facet = np.full(len(rows), None, dtype=object)
        for name, normal in FACES.items():  # later names win, like the reference's successive .loc assignments
            axis = int(np.argmax(np.abs(normal)))
            on = np.isclose(position[:, axis], normal[axis] * half[axis], atol=2.220446049250313e-13)
            facet[on] = name
        names = {e.value: e.name.lower() for e in Event}
        ray_of_row = rows // m  # ordinal of the logged ray each row belongs to
        recorded = result.recorded_indices
        components = list(result.compiled.component_names)
        src = d["source"][rows]
        source = np.array([components[int(c)] if c >= 0 else result.sources[int(recorded[int(j)])]
                           for c, j in zip(src, ray_of_row)], dtype=object)
        return {"kind": kinds, "event": np.array([names[int(k)] for k in d["kind"][rows]]), "facet": facet,
                "luminescent": src >= 0, "source": source, "wavelength": d["wavelength"][rows]}

    def spectrum(self, facets=(), kind="last", source="all", events=None):
        """Wavelengths of the stored end rays (lsc.py:508-575): kind 'first' (entrance) | 'last' (exit) | None;
        source 'all', the name of a light or component, or a collection of names -- as in the reference -- and, as
        shorthands, 'light' (any light) | 'luminescent' (any component); facets a collection of face names (empty:
        any); events a collection of lower-case event names (None: any)."""
        if kind not in (None, "first", "last"):
            raise ValueError("Direction must be either `'first'` or `'last'.`")
        rows = self._end_rays()
        keep = np.ones(len(rows["kind"]), dtype=bool)
        if kind is not None:
            keep &= rows["kind"] == ("entrance" if kind == "first" else "exit")
        if isinstance(source, str) and source in ("light", "luminescent") and source not in self.light_names() | self.component_names():
            keep &= rows["luminescent"] == (source == "luminescent")
        elif not (isinstance(source, str) and source == "all"):
            wanted = {source} if isinstance(source, str) else set(source)
            unknown = wanted - (self.component_names() | self.light_names())
            if unknown:
                raise ValueError("Unknown source requested.", unknown)
            keep &= np.isin(rows["source"].astype(str), sorted(wanted))
        if len(facets) > 0:
            keep &= np.isin(rows["facet"].astype(str), list(facets))
        if events is not None:
            keep &= np.isin(rows["event"], list(events))
        return rows["wavelength"][keep]

    def counts_table(self):
        """{column: {facet: count}} with the reference's four columns (lsc.py:455-506)."""
        table = {}
        for column, source, kind in (("Solar In", "light", "first"), ("Solar Out", "light", "last"),
                                     ("Luminescent Out", "luminescent", "last"), ("Luminescent In", "luminescent", "first")):
            table[column] = {face: int(len(self.spectrum(facets={face}, source=source, kind=kind)))
                             for face in ("left", "right", "near", "far", "top", "bottom")}
        return table

    def summary(self):
        """Efficiencies of the run (lsc.py:577-607), over the logged rays."""
        table = self.counts_table()
        faces = set(FACES)
        collected = sum(table["Luminescent Out"][f] for f in self._solar_cell_surfaces)
        escaped = sum(table["Luminescent Out"][f] for f in faces - self._solar_cell_surfaces)
        incident = sum(table["Solar In"][f] for f in faces)
        lost = int(len(self.spectrum(kind="last", events={"absorb", "nonradiative", "react"})))
        (l, w, d) = self.size
        geometric = (w * l) / (2 * l * d + 2 * w * d)
        n = self.n1
        return {"Optical Efficiency": collected / incident if incident else float("nan"),
                "Waveguide Efficiency": collected / (collected + escaped) if collected + escaped else float("nan"),
                "Waveguide Efficiency (Thermodynamic Prediction)": n ** 2 / (geometric + n ** 2),
                "Non-radiative Loss (fraction):": lost / incident if incident else float("nan"),
                "Incident": incident, "Geometric Concentration": geometric, "Refractive Index": n,
                "Cell Surfaces": set(self._solar_cell_surfaces), "Components": self.component_names(),
                "Lights": self.light_names()}

    def report(self):
        print("\nSimulation Report\n-----------------\n\nSurface Counts:")
        table = self.counts_table()
        columns = list(table)
        print("        " + "  ".join(f"{c:>15s}" for c in columns))
        for face in ("left", "right", "near", "far", "top", "bottom"):
            print(f"{face:8s}" + "  ".join(f"{table[c][face]:15d}" for c in columns))
        print("\nSummary:")
        for key, value in self.summary().items():
            print(f"{key:50s} {value}")
