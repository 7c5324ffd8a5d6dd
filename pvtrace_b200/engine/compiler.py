"""Flattener: Scene graph -> SoA tables for the device tracer.

Table contract (names, dtypes, shapes, tag values) is that of pvtrace.engine.compiler.CompiledScene
(pvtrace/engine/compiler.py:57-204, read by the kernel at pvtrace/engine/_kernel.pyx:933-1017), so a table set
built here can be fed unchanged to the reference's compiled kernel (that is how the CPU baseline and the
per-ray parity tests run).  Two extensions:

* facet tables (`facet_*`): per-node, per-facet surface overrides lowered from `FacetSurfaceDelegate`
  (and therefore from the LSC device's solar-cell / mirror options, pvtrace/device/lsc.py:22-86, which the
  reference engine rejects at compiler.py:239-247);
* `CompiledEmitter`: descriptors of the built-in light delegates, including functools.partial wrappers
  (which the reference's vectorised emitter does not recognise, pvtrace/engine/emit.py:67,116-124), so initial
  rays are sampled on the device.
"""
import functools

import numpy as np

from pvtrace_b200.engine.recorder import EVENTS, PROPERTIES, Heatmap, Recorder
from pvtrace_b200.geometry.box import Box
from pvtrace_b200.geometry.cylinder import Cylinder
from pvtrace_b200.geometry.sphere import Sphere
from pvtrace_b200.light import light as light_module
from pvtrace_b200.material import utils as material_utils
from pvtrace_b200.material.component import Absorber, Luminophore, Reactor, Scatterer
from pvtrace_b200.material.surface import FacetSurfaceDelegate, FresnelSurfaceDelegate, NullSurfaceDelegate

VOLUME_EVENTS = {"lost", "reacted", "killed"}  # cannot be filtered by facet
MAX_RECORDERS = 256
MAX_NODES = 128

GEOM_BOX, GEOM_SPHERE, GEOM_CYLINDER = 0, 1, 2
SURF_FRESNEL, SURF_NULL = 0, 1
COMP_ABSORBER, COMP_SCATTERER, COMP_LUMINOPHORE, COMP_REACTOR = 0, 1, 2, 3
PHASE_ISOTROPIC, PHASE_HENYEY_GREENSTEIN, PHASE_CONE = 0, 1, 2
EMIT_KT, EMIT_REDSHIFT, EMIT_FULL = 0, 1, 2
EMIT_METHODS = {"kT": EMIT_KT, "redshift": EMIT_REDSHIFT, "full": EMIT_FULL}
FACET_TRANSMIT_STRAIGHT, FACET_REFLECT_LAMBERTIAN = 1, 2

LPOS_POINT, LPOS_RECT, LPOS_CIRCLE, LPOS_CUBE = 0, 1, 2, 3
LDIR_Z, LDIR_CONE, LDIR_ISOTROPIC, LDIR_LAMBERTIAN, LDIR_HG = 0, 1, 2, 3, 4
LWL_CONSTANT, LWL_SPECTRUM = 0, 1


class UnsupportedSceneError(Exception):
    """The scene uses a feature the compiled engine does not support."""


def _i32(values):
    return np.ascontiguousarray(values, dtype=np.int32)


def _f64(values):
    return np.ascontiguousarray(values, dtype=np.float64)


class _SpectrumPool:
    """Concatenated (x, y) knot arrays with (start, n) handles."""

    def __init__(self):
        self.x, self.y = [], []

    def add(self, x, y):
        start = len(self.x)
        self.x.extend(np.asarray(x, dtype=np.float64).tolist())
        self.y.extend(np.asarray(y, dtype=np.float64).tolist())
        return start, len(self.x) - start


class CompiledScene:
    """Flat-table representation of a scene (see module docstring for the contract)."""

    def __init__(self, scene):
        root = scene.root
        nodes = [n for n in root.iter_preorder() if n.geometry is not None]
        if not nodes:
            raise UnsupportedSceneError("Scene has no geometry nodes.")
        if root.geometry is None:
            raise UnsupportedSceneError("Root node must have a geometry.")
        if len(nodes) > MAX_NODES:
            raise ValueError(f"Engine supports at most {MAX_NODES} geometry nodes.")

        self.scene = scene
        self.nodes = nodes
        self.node_names = [n.name for n in nodes]
        self.root_id = nodes.index(root)

        self._lower_geometry(nodes, root)
        self._lower_materials(nodes)
        self._lower_recorders(nodes)

    # -- geometry + pose ------------------------------------------------------------------

    def _lower_geometry(self, nodes, root):
        count = len(nodes)
        self.geom_type = np.zeros(count, dtype=np.int32)
        self.geom_params = np.zeros((count, 4), dtype=np.float64)
        self.local_to_world = np.zeros((count, 4, 4), dtype=np.float64)
        self.world_to_local = np.zeros((count, 4, 4), dtype=np.float64)
        for i, node in enumerate(nodes):
            g = node.geometry
            if isinstance(g, Box):
                self.geom_type[i] = GEOM_BOX
                self.geom_params[i, :3] = np.asarray(g.size, dtype=np.float64)
            elif isinstance(g, Sphere):
                self.geom_type[i] = GEOM_SPHERE
                self.geom_params[i, 0] = float(g.radius)
            elif isinstance(g, Cylinder):
                self.geom_type[i] = GEOM_CYLINDER
                self.geom_params[i, :2] = (float(g.length), float(g.radius))
            else:
                raise UnsupportedSceneError(f"Geometry type {type(g).__name__} is not supported.")
            pose = np.asarray(node.transformation_to(root), dtype=np.float64)
            rot = pose[:3, :3]
            if not np.allclose(rot @ rot.T, np.eye(3), atol=1e-9):
                raise UnsupportedSceneError(f"Node {node.name!r} transform is not rigid (has scale or shear).")
            self.local_to_world[i] = pose
            self.world_to_local[i] = np.linalg.inv(pose)

    # -- materials: refractive index, surface, components, spectra --------------------------

    def _lower_materials(self, nodes):
        count = len(nodes)
        self.refractive_index = np.zeros(count, dtype=np.float64)
        self.surface_type = np.zeros(count, dtype=np.int32)
        self.comp_start = np.zeros(count, dtype=np.int32)
        self.comp_count = np.zeros(count, dtype=np.int32)
        self.facet_start = np.zeros(count, dtype=np.int32)
        self.facet_count = np.zeros(count, dtype=np.int32)
        self.component_names = []
        absorb, emit = _SpectrumPool(), _SpectrumPool()
        columns = {k: [] for k in ("type", "qy", "tau_rad", "tau_nr", "phase_type", "phase_param",
                                   "abs_start", "abs_n", "ems_start", "ems_n")}
        facets = {k: [] for k in ("normal", "atol", "reflectivity", "flags", "region", "refl_start", "refl_n")}
        facets["spectra"] = _SpectrumPool()

        for i, node in enumerate(nodes):
            material = node.geometry.material
            if material is None:
                raise UnsupportedSceneError(f"Node {node.name!r} has geometry without a material.")
            self.refractive_index[i] = float(material.refractive_index)
            self.surface_type[i] = self._lower_surface(node, material, i, facets)
            self.comp_start[i] = len(self.component_names)
            self.comp_count[i] = len(material.components)
            for component in material.components:
                self._lower_component(node, component, columns, absorb, emit)
                self.component_names.append(component.name)

        self.comp_type = _i32(columns["type"])
        self.comp_qy = _f64(columns["qy"])
        self.comp_tau_rad = _f64(columns["tau_rad"])
        self.comp_tau_nr = _f64(columns["tau_nr"])
        self.comp_phase_type = _i32(columns["phase_type"])
        self.comp_phase_param = _f64(columns["phase_param"])
        self.comp_abs_start = _i32(columns["abs_start"])
        self.comp_abs_n = _i32(columns["abs_n"])
        self.comp_ems_start = _i32(columns["ems_start"])
        self.comp_ems_n = _i32(columns["ems_n"])
        self.abs_x, self.abs_y = _f64(absorb.x), _f64(absorb.y)
        self.ems_x, self.ems_cdf = _f64(emit.x), _f64(emit.y)

        self.n_facets = len(facets["atol"])
        self.facet_normal = _f64(facets["normal"]).reshape(-1, 3)
        self.facet_atol = _f64(facets["atol"])
        self.facet_reflectivity = _f64(facets["reflectivity"])
        self.facet_flags = _i32(facets["flags"])
        self.facet_region = _f64(facets["region"]).reshape(-1, 6)
        self.facet_refl_start = _i32(facets["refl_start"])
        self.facet_refl_n = _i32(facets["refl_n"])
        self.refl_x, self.refl_y = _f64(facets["spectra"].x), _f64(facets["spectra"].y)

    def _lower_surface(self, node, material, index, facets):
        delegate = material.surface.delegate
        if isinstance(delegate, FacetSurfaceDelegate):
            self.facet_start[index] = len(facets["atol"])
            for facet in delegate.facets:
                facets["normal"].append(facet.normal)
                facets["atol"].append(facet.atol)
                facets["reflectivity"].append(-1.0 if facet.reflectivity is None else facet.reflectivity)
                facets["flags"].append((FACET_TRANSMIT_STRAIGHT if facet.transmit == "straight" else 0)
                                       | (FACET_REFLECT_LAMBERTIAN if facet.reflect == "lambertian" else 0))
                facets["region"].append(tuple(facet.region[0]) + tuple(facet.region[1]))
                if facet.spectrum is not None:
                    start, n = facets["spectra"].add(facet.spectrum[0], facet.spectrum[1])
                else:
                    start, n = 0, 0
                facets["refl_start"].append(start)
                facets["refl_n"].append(n)
            self.facet_count[index] = len(facets["atol"]) - self.facet_start[index]
            return SURF_FRESNEL
        if type(delegate) is FresnelSurfaceDelegate:
            return SURF_FRESNEL
        if type(delegate) is NullSurfaceDelegate:
            return SURF_NULL
        raise UnsupportedSceneError(
            f"Node {node.name!r} uses surface delegate {type(delegate).__name__}; only FresnelSurfaceDelegate, "
            "NullSurfaceDelegate and FacetSurfaceDelegate are supported.")

    @staticmethod
    def _phase(node, component):
        phase = component.phase_function
        if phase is material_utils.isotropic:
            return PHASE_ISOTROPIC, 0.0
        if isinstance(phase, material_utils.HenyeyGreenstein):
            return PHASE_HENYEY_GREENSTEIN, float(phase.g)
        if isinstance(phase, material_utils.Cone):
            return PHASE_CONE, float(phase.theta_max)
        kind = _partial_of(phase, material_utils.henyey_greenstein, ("g",))
        if kind is not None:
            return PHASE_HENYEY_GREENSTEIN, float(kind[0])
        kind = _partial_of(phase, material_utils.cone, ("theta_max",))
        if kind is not None:
            return PHASE_CONE, float(kind[0])
        raise UnsupportedSceneError(f"Node {node.name!r}: custom phase functions are not supported.")

    def _lower_component(self, node, component, columns, absorb, emit):
        # subclass order matters: Reactor < Absorber < Scatterer and Luminophore < Scatterer
        if isinstance(component, Reactor):
            ctype = COMP_REACTOR
        elif isinstance(component, Absorber):
            ctype = COMP_ABSORBER
        elif isinstance(component, Luminophore):
            ctype = COMP_LUMINOPHORE
        elif isinstance(component, Scatterer):
            ctype = COMP_SCATTERER
        else:
            raise UnsupportedSceneError(f"Component type {type(component).__name__} is not supported.")
        phase_type, phase_param = self._phase(node, component)

        dist = component._abs_dist
        if dist.hist:
            raise UnsupportedSceneError(f"Node {node.name!r}: histogram-sampled spectra are not supported.")
        if dist._x is None:
            a_start, a_n = absorb.add([0.0], [float(dist._y)])  # wavelength-independent coefficient
        else:
            a_start, a_n = absorb.add(dist._x, dist._y)

        e_start, e_n = 0, 0
        if ctype == COMP_LUMINOPHORE:
            ems = component._ems_dist
            if ems.hist:
                raise UnsupportedSceneError(
                    f"Node {node.name!r}: histogram-sampled emission spectra are not supported.")
            e_start, e_n = emit.add(ems._x, ems._cdf)

        row = {"type": ctype, "qy": float(component.quantum_yield),
               "tau_rad": float(component.tau_rad or 0.0), "tau_nr": float(component.tau_nr or 0.0),
               "phase_type": phase_type, "phase_param": phase_param,
               "abs_start": a_start, "abs_n": a_n, "ems_start": e_start, "ems_n": e_n}
        for key, value in row.items():
            columns[key].append(value)

    # -- recorders ------------------------------------------------------------------------

    def _lower_recorders(self, nodes):
        attached = []
        for i, node in enumerate(nodes):
            for rec in getattr(node, "recorders", []):
                if not isinstance(rec, Recorder):
                    raise UnsupportedSceneError(f"Node {node.name!r} recorders must be Recorder objects.")
                if rec.event in VOLUME_EVENTS and rec.facet is not None:
                    raise UnsupportedSceneError(
                        f"Recorder {rec.name!r}: facet filters only apply to surface events.")
                attached.append((i, rec))
        if len(attached) > MAX_RECORDERS:
            raise UnsupportedSceneError(f"At most {MAX_RECORDERS} recorders are supported.")
        names = [rec.name for _, rec in attached]
        if len(set(names)) != len(names):
            raise UnsupportedSceneError("Recorder names must be unique.")

        count = len(attached)
        self.recorder_names = names
        self.recorder_specs = [rec for _, rec in attached]
        self.rec_node = _i32([i for i, _ in attached])
        self.rec_event = _i32([EVENTS[rec.event] for _, rec in attached])
        self.rec_has_facet = _i32([rec.facet is not None for _, rec in attached])
        self.rec_facet = np.zeros((max(count, 1), 3), dtype=np.float64)
        self.rec_atol = _f64([rec.atol for _, rec in attached])
        self.rec_hist_start = np.zeros(count, dtype=np.int32)
        self.rec_hist_n = np.zeros(count, dtype=np.int32)

        hists = {k: [] for k in ("prop_a", "prop_b", "na", "nb", "lo_a", "hi_a", "lo_b", "hi_b", "offset")}
        offset = 0
        for r, (_, rec) in enumerate(attached):
            if rec.facet is not None:
                self.rec_facet[r] = rec.facet
            self.rec_hist_start[r] = len(hists["offset"])
            self.rec_hist_n[r] = len(rec.histograms)
            for h in rec.histograms:
                a, b = (h.a, h.b) if isinstance(h, Heatmap) else (h, None)
                hists["prop_a"].append(PROPERTIES[a.prop])
                hists["na"].append(a.bins)
                hists["lo_a"].append(a.start)
                hists["hi_a"].append(a.stop)
                hists["prop_b"].append(-1 if b is None else PROPERTIES[b.prop])
                hists["nb"].append(1 if b is None else b.bins)
                hists["lo_b"].append(0.0 if b is None else b.start)
                hists["hi_b"].append(1.0 if b is None else b.stop)
                hists["offset"].append(offset)
                offset += a.bins * (1 if b is None else b.bins)
        self.hist_prop_a, self.hist_prop_b = _i32(hists["prop_a"]), _i32(hists["prop_b"])
        self.hist_na, self.hist_nb = _i32(hists["na"]), _i32(hists["nb"])
        self.hist_lo_a, self.hist_hi_a = _f64(hists["lo_a"]), _f64(hists["hi_a"])
        self.hist_lo_b, self.hist_hi_b = _f64(hists["lo_b"]), _f64(hists["hi_b"])
        self.hist_offset = _i32(hists["offset"])
        self.total_bins = int(offset)


def compile_scene(scene) -> CompiledScene:
    """Compile `scene` into flat tables, or raise `UnsupportedSceneError`."""
    return CompiledScene(scene)


# ---------------------------------------------------------------------------------------------
# Emission descriptors


def _partial_of(delegate, func, arg_names):
    """If `delegate` is functools.partial(func, ...) binding exactly `arg_names`, return the bound values."""
    if not isinstance(delegate, functools.partial) or delegate.func is not func:
        return None
    bound = dict(zip(arg_names, delegate.args))
    bound.update(delegate.keywords or {})
    if set(bound) != set(arg_names) or len(delegate.args) > len(arg_names):
        return None
    return [bound[name] for name in arg_names]


def _describe_wavelength(delegate):
    if delegate is light_module.default_wavelength or isinstance(delegate, light_module.DefaultWavelength):
        return LWL_CONSTANT, 555.0, None
    if isinstance(delegate, light_module.ConstantWavelengthMask):
        return LWL_CONSTANT, float(delegate.nanometers), None
    if isinstance(delegate, light_module.SpectrumWavelengthMask):
        dist = delegate.distribution
        if dist.hist or dist._x is None:
            return None
        return LWL_SPECTRUM, 0.0, (dist._x, dist._cdf)
    return None


def _describe_position(delegate):
    L = light_module
    if delegate is L.default_position or isinstance(delegate, L.DefaultPosition):
        return LPOS_POINT, (0.0, 0.0, 0.0)
    if isinstance(delegate, L.RectangularMask):
        return LPOS_RECT, (delegate.x, delegate.y, 0.0)
    if isinstance(delegate, L.CircularMask):
        return LPOS_CIRCLE, (delegate.radius, 0.0, 0.0)
    if isinstance(delegate, L.CubeMask):
        return LPOS_CUBE, (delegate.x, delegate.y, delegate.z)
    for func, names, kind in ((L.rectangular_mask, ("X", "Y"), LPOS_RECT),
                              (L.circular_mask, ("radius",), LPOS_CIRCLE),
                              (L.cube_mask, ("X", "Y", "Z"), LPOS_CUBE)):
        bound = _partial_of(delegate, func, names)
        if bound is not None:
            return kind, tuple(float(v) for v in bound) + (0.0,) * (3 - len(bound))
    return None


def _describe_direction(delegate):
    M = material_utils
    if delegate is light_module.default_direction or isinstance(delegate, light_module.DefaultDirection):
        return LDIR_Z, 0.0
    if isinstance(delegate, M.Cone):
        return LDIR_CONE, delegate.theta_max
    if delegate is M.isotropic:
        return LDIR_ISOTROPIC, 0.0
    if delegate is M.lambertian:
        return LDIR_LAMBERTIAN, 0.0
    if isinstance(delegate, M.HenyeyGreenstein):
        return LDIR_HG, delegate.g
    bound = _partial_of(delegate, M.cone, ("theta_max",))
    if bound is not None:
        theta = float(bound[0])
        if np.isclose(theta, 0.0) or theta > np.pi / 2:
            raise ValueError("Expected 0 < theta_max <= pi/2")
        return LDIR_CONE, theta
    bound = _partial_of(delegate, M.henyey_greenstein, ("g",))
    if bound is not None:
        return LDIR_HG, float(bound[0])
    return None


class CompiledEmitter:
    """Device-sampleable description of every light in the scene (None-able: see `compile_emitter`)."""

    def __init__(self, scene, described):
        lights = scene.light_nodes
        count = len(lights)
        self.light_names = [node.light.name for node in lights]
        self.light_to_world = np.zeros((count, 4, 4), dtype=np.float64)
        self.pos_kind = np.zeros(count, dtype=np.int32)
        self.pos_param = np.zeros((count, 3), dtype=np.float64)
        self.dir_kind = np.zeros(count, dtype=np.int32)
        self.dir_param = np.zeros(count, dtype=np.float64)
        self.wl_kind = np.zeros(count, dtype=np.int32)
        self.wl_param = np.zeros(count, dtype=np.float64)
        self.wl_start = np.zeros(count, dtype=np.int32)
        self.wl_n = np.zeros(count, dtype=np.int32)
        pool = _SpectrumPool()
        for i, (node, (wl, pos, direction)) in enumerate(zip(lights, described)):
            self.light_to_world[i] = np.asarray(node.transformation_to(scene.root), dtype=np.float64)
            self.wl_kind[i], self.wl_param[i] = wl[0], wl[1]
            if wl[2] is not None:
                self.wl_start[i], self.wl_n[i] = pool.add(*wl[2])
            self.pos_kind[i], self.pos_param[i] = pos[0], pos[1]
            self.dir_kind[i], self.dir_param[i] = direction[0], float(direction[1])
        self.wl_x, self.wl_cdf = _f64(pool.x), _f64(pool.y)
        self.n_lights = count


def compile_emitter(scene):
    """Emission descriptors when EVERY light uses recognised built-in delegates, else None."""
    lights = scene.light_nodes
    if not lights:
        raise UnsupportedSceneError("Scene has no light nodes.")
    described = []
    for node in lights:
        light = node.light
        parts = (_describe_wavelength(light.wavelength), _describe_position(light.position),
                 _describe_direction(light.direction))
        if any(p is None for p in parts):
            return None
        described.append(parts)
    return CompiledEmitter(scene, described)
