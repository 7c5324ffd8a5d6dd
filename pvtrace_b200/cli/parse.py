"""YAML scene specification (version "1.0") -> Scene, including the `recorders:` section and `record: true`.

The file format is the reference's (pvtrace/cli/parse.py:72-551 with its JSON schema): `nodes` (one of box /
sphere / cylinder / light per node, optional parent / location / direction / record), `components` (absorber /
scatterer / luminophore with constant coefficients, named or CSV spectra, phase functions) and `recorders`.  It is the
widening step SURVEY 8(f)3: the caller one step upstream of the flattener.  The scene it returns feeds
`engine.simulate` directly -- every light mask it can express lowers to the on-device emitter.

Not supported: `mesh` nodes (the engine has no triangle meshes; the reference's own engine rejects them too,
pvtrace/engine/compiler.py:220-223) -> UnsupportedSceneError.
"""
import os

import numpy as np
import yaml

from pvtrace_b200.data import fluro_red, lumogen_f_red_305
from pvtrace_b200.engine.compiler import UnsupportedSceneError
from pvtrace_b200.engine.recorder import Heatmap, Histogram, Recorder, auto_recorders
from pvtrace_b200.geometry.box import Box
from pvtrace_b200.geometry.cylinder import Cylinder
from pvtrace_b200.geometry.sphere import Sphere
from pvtrace_b200.light.light import (CircularMask, ConstantWavelengthMask, CubeMask, Light, RectangularMask,
                                      SpectrumWavelengthMask)
from pvtrace_b200.material.component import Absorber, Luminophore, Scatterer
from pvtrace_b200.material.distribution import Distribution
from pvtrace_b200.material.material import Material
from pvtrace_b200.material.utils import Cone, HenyeyGreenstein, isotropic, lambertian
from pvtrace_b200.scene.node import Node
from pvtrace_b200.scene.scene import Scene

NAMED_SPECTRA = {"lumogen-f-red-305": lumogen_f_red_305, "fluro-red": fluro_red}
SUPPORTED_VERSIONS = ("1.0",)


class SpecError(ValueError):
    """The scene file does not follow the specification."""


def _need(mapping, key, where):
    if not isinstance(mapping, dict) or key not in mapping:
        raise SpecError(f"{where}: missing `{key}`")
    return mapping[key]


class _Builder:
    def __init__(self, spec, directory):
        self.spec, self.directory = spec, directory
        self.components = {}

    # -- spectra -------------------------------------------------------------------------------

    def spectrum(self, spec, kind, where):
        """(N, 2) array of (nanometres, value) from a CSV file or one of the built-in dyes."""
        if "file" in spec:
            path = spec["file"]
            if not os.path.isabs(path):
                path = os.path.abspath(os.path.join(self.directory, path))
            table = np.genfromtxt(path, delimiter=",", skip_header=1)
            if table.ndim != 2 or table.shape[1] < 3:
                raise SpecError(f"{where}: {path} must have an index column followed by x and y columns")
            return np.ascontiguousarray(table[:, 1:3], dtype=float)  # column 0 is the row index
        name = _need(spec, "name", where)
        if name not in NAMED_SPECTRA:
            raise SpecError(f"{where}: unknown spectrum {name!r}; use one of {sorted(NAMED_SPECTRA)}")
        rng = _need(spec, "range", where)
        x = np.arange(rng["min"], rng["max"] + rng["spacing"], rng["spacing"])
        module = NAMED_SPECTRA[name]
        y = module.absorption(x) if kind == "absorption" else module.emission(x)
        return np.column_stack((x, y))

    @staticmethod
    def scaled(spectrum, coefficient):
        """Spectrum rescaled so that its peak equals `coefficient` (cm^-1)."""
        out = spectrum.copy()
        out[:, 1] = out[:, 1] / np.max(out[:, 1]) * coefficient
        return out

    # -- components ----------------------------------------------------------------------------

    def phase_function(self, spec, where):
        if spec == "isotropic" or (isinstance(spec, dict) and "isotropic" in spec):
            return isotropic
        if spec == "lambertian" or (isinstance(spec, dict) and "lambertian" in spec):
            return lambertian
        if isinstance(spec, dict) and "cone" in spec:
            return Cone(float(np.radians(float(_need(spec["cone"], "half-angle", where)))))
        if isinstance(spec, dict) and "henyey-greenstein" in spec:
            return HenyeyGreenstein(float(_need(spec["henyey-greenstein"], "g", where)))
        raise SpecError(f"{where}: unknown phase function {spec!r}")

    def attenuation(self, spec, kind, where):
        """Constant coefficient, spectrum, or spectrum scaled to a peak coefficient."""
        coefficient = spec.get("coefficient")
        table = self.spectrum(spec["spectrum"], kind, where) if "spectrum" in spec else None
        if coefficient and table is not None:
            return self.scaled(table, coefficient)
        if table is not None:
            return table
        if coefficient:
            return float(coefficient)
        raise SpecError(f"{where}: needs a `coefficient`, a `spectrum`, or both")

    def component(self, name, spec):
        where = f"component {name!r}"
        if "absorber" in spec:
            body = spec["absorber"]
            return Absorber(self.attenuation(body, "absorption", where), name=name, hist=bool(body.get("hist", False)))
        if "scatterer" in spec:
            body = spec["scatterer"]
            phase = self.phase_function(body["phase-function"], where) if "phase-function" in body else None
            return Scatterer(self.attenuation(body, "absorption", where), quantum_yield=body.get("quantum-yield", 1.0),
                             phase_function=phase, name=name, hist=bool(body.get("hist", False)))
        if "luminophore" in spec:
            body = spec["luminophore"]
            emission = _need(body, "emission", where)
            phase = self.phase_function(emission["phase-function"], where) if "phase-function" in emission else isotropic
            return Luminophore(self.attenuation(_need(body, "absorption", where), "absorption", where),
                               emission=self.spectrum(_need(emission, "spectrum", where), "emission", where),
                               quantum_yield=emission.get("quantum-yield", 1.0), phase_function=phase, name=name,
                               hist=bool(body.get("hist", False)))
        raise SpecError(f"{where}: expected one of absorber / scatterer / luminophore")

    # -- nodes ---------------------------------------------------------------------------------

    def material(self, spec, where):
        keys = spec.get("components", []) or []
        missing = [k for k in keys if k not in self.components]
        if missing:
            raise SpecError(f"{where}: missing {missing[0]} component")
        return Material(refractive_index=_need(spec, "refractive-index", where),
                        components=[self.components[k] for k in keys])

    def light(self, name, spec):
        where = f"light {name!r}"
        wavelength = ConstantWavelengthMask(spec["wavelength"]) if spec.get("wavelength") else None
        position = direction = None
        mask = spec.get("mask") or {}
        if mask.get("wavelength"):
            w = mask["wavelength"]
            if "nanometers" in w:
                wavelength = ConstantWavelengthMask(float(w["nanometers"]))
            elif "spectrum" in w:
                table = self.spectrum(w["spectrum"], "absorption", where)
                wavelength = SpectrumWavelengthMask(Distribution(table[:, 0], table[:, 1]))
            else:
                raise SpecError(f"{where}: wavelength mask needs `nanometers` or `spectrum`")
        if mask.get("position"):
            p = mask["position"]
            if "rect" in p:
                position = RectangularMask(*p["rect"])
            elif "cube" in p:
                position = CubeMask(*p["cube"])
            elif "circle" in p:
                position = CircularMask(p["circle"])
            else:
                raise SpecError(f"{where}: position mask needs rect / cube / circle")
        if mask.get("direction"):
            direction = self.phase_function(mask["direction"], where)
        return Light(position=position, direction=direction, wavelength=wavelength, name=name)

    def node(self, name, spec):
        where = f"node {name!r}"
        if "mesh" in spec:
            raise UnsupportedSceneError(f"{where}: mesh geometry is not supported by the engine")
        if "box" in spec:
            geometry = Box(size=_need(spec["box"], "size", where), material=self.material(_need(spec["box"], "material", where), where))
        elif "sphere" in spec:
            geometry = Sphere(radius=_need(spec["sphere"], "radius", where),
                              material=self.material(_need(spec["sphere"], "material", where), where))
        elif "cylinder" in spec:
            body = spec["cylinder"]
            geometry = Cylinder(length=_need(body, "length", where), radius=_need(body, "radius", where),
                                material=self.material(_need(body, "material", where), where))
        elif "light" in spec:
            return Node(name=name, light=self.light(name, spec["light"] or {}))
        else:
            raise SpecError(f"{where}: expected one of box / sphere / cylinder / light")
        return Node(name=name, geometry=geometry)

    # -- recorders -----------------------------------------------------------------------------

    @staticmethod
    def recorder(name, spec):
        histograms = []
        for prop, values in (spec.get("histograms") or {}).items():
            if prop == "position":
                prop_a, prop_b, range_a, range_b = values
                histograms.append(Heatmap(prop_a, prop_b, range_a, range_b))
            else:
                histograms.append(Histogram(prop, *values))
        return Recorder(name, event=_need(spec, "event", f"recorder {name!r}"), facet=spec.get("facet"),
                        atol=spec.get("atol", 1e-6), histograms=histograms)

    def build(self):
        for name, body in (self.spec.get("components") or {}).items():
            self.components[name] = self.component(name, body)
        node_specs = _need(self.spec, "nodes", "scene")
        if "world" not in node_specs:
            raise SpecError("scene: a node named `world` is required")
        nodes = {name: self.node(name, body) for name, body in node_specs.items()}
        for name, body in node_specs.items():  # parents and poses once every node exists
            node = nodes[name]
            if name != "world":
                parent = body.get("parent") or "world"
                if parent not in nodes:
                    raise SpecError(f"node {name!r}: unknown parent {parent!r}")
                node.parent = nodes[parent]
            if body.get("location"):
                node.location = body["location"]
            if body.get("direction"):
                node.look_at(body["direction"])
        # explicit recorders first (in file order), then the `record: true` expansions that they do not override
        explicit = dict(self.spec.get("recorders") or {})
        for name, body in explicit.items():
            target = _need(body, "node", f"recorder {name!r}")
            if target not in nodes:
                raise SpecError(f"Recorder {name!r}: unknown node {target!r}")
            nodes[target].recorders.append(self.recorder(name, body))
        for name, body in node_specs.items():
            if body.get("record") and nodes[name].geometry is not None:
                for rec in auto_recorders(name, nodes[name].geometry):
                    if rec.name not in explicit:
                        nodes[name].recorders.append(rec)
        return Scene(nodes["world"])


def parse_spec(spec, directory="."):
    """Scene from an already-loaded specification dictionary."""
    version = str(_need(spec, "version", "scene"))
    if version not in SUPPORTED_VERSIONS:
        raise ValueError(f"Version {version} not supported")
    return _Builder(spec, directory).build()


def parse(filename):
    """Scene from a YAML scene file (role of pvtrace.cli.parse.parse)."""
    with open(filename, "r") as handle:
        spec = yaml.safe_load(handle)
    return parse_spec(spec, os.path.dirname(os.path.abspath(filename)))
