"""Pins for the ORACLE'S EXTENSIONS beyond the reference's compiled kernel -- seeded emission and the facet / coating
table -- generated from the UNMODIFIED reference (imported from /root/reference through oracle/ref_loader.py with a
`meshcat` stand-in; run in the build container only):

    python tests/golden/make_reference_pins.py

emission.npz        quantiles and second moments of pvtrace.engine.emit.emit_bundle (emit.py:92-134) for every built-in
                    light delegate (light.py:48-157, material/utils.py cone / isotropic / lambertian / Henyey-Greenstein),
                    10^5 rays each.
lsc_delegates.npz   reflectivity / reflected / transmitted direction of OptionalMirrorAndSolarCell and AirGapMirror
                    (pvtrace/device/lsc.py:22-86) called directly on rays at every face of the LSC, from both sides.
coatings.npz        the same for two user delegates in the style of examples/006 Coatings.ipynb cell 3: the quarter
                    mirror on the top face (reflectivity depends on WHERE the face is hit) and a coating whose
                    reflectivity depends on the wavelength; plus per-ray event statistics of the reference's PYTHON
                    tracer (photon_tracer.follow) on a disc carrying both coatings.
tests/test_reference_pins.py compares the oracle / host restatements with these files.
"""
import functools
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import ref_loader  # noqa: E402

ref_loader.load_reference_package()
_meshcat = types.ModuleType("meshcat")
for _sub in ("geometry", "transformations"):
    _m = types.ModuleType("meshcat." + _sub)
    setattr(_meshcat, _sub, _m)
    sys.modules["meshcat." + _sub] = _m
_meshcat.Visualizer = object
sys.modules.setdefault("meshcat", _meshcat)

from pvtrace.algorithm import photon_tracer  # noqa: E402
from pvtrace.device import lsc as ref_lsc  # noqa: E402
from pvtrace.engine.emit import emit_bundle  # noqa: E402
from pvtrace.geometry.box import Box  # noqa: E402
from pvtrace.geometry.cylinder import Cylinder  # noqa: E402
from pvtrace.geometry.sphere import Sphere  # noqa: E402
from pvtrace.light import light as ref_light  # noqa: E402
from pvtrace.light.event import Event  # noqa: E402
from pvtrace.light.ray import Ray  # noqa: E402
from pvtrace.material.component import Absorber  # noqa: E402
from pvtrace.material.distribution import Distribution  # noqa: E402
from pvtrace.material.material import Material  # noqa: E402
from pvtrace.material.surface import FresnelSurfaceDelegate, Surface  # noqa: E402
from pvtrace.material import utils as ref_utils  # noqa: E402
from pvtrace.scene.node import Node  # noqa: E402
from pvtrace.scene.scene import Scene  # noqa: E402

QUANTILES = np.linspace(0.0, 1.0, 201)


def lamp(x):
    return np.exp(-(((x - 520.0) / 60.0) ** 2)) + 0.6 * np.exp(-(((x - 600.0) / 25.0) ** 2))


# ---- emission ------------------------------------------------------------------------------------------------
# (name, position delegate, direction delegate, wavelength delegate); tests/test_reference_pins.py builds the same
# lights with pvtrace_b200's classes
def emission_cases():
    x = np.linspace(400.0, 700.0, 121)
    return {
        "rect_cone": (ref_light.RectangularMask(1.5, 0.7), ref_utils.Cone(0.4), ref_light.ConstantWavelengthMask(600.0)),
        "circle_isotropic": (ref_light.CircularMask(2.0), ref_utils.isotropic,
                             ref_light.SpectrumWavelengthMask(Distribution(x, lamp(x)))),
        "cube_lambertian": (ref_light.CubeMask(1.0, 2.0, 0.5), ref_utils.lambertian, None),
        "point_hg": (None, ref_utils.HenyeyGreenstein(0.7), ref_light.ConstantWavelengthMask(450.0)),
        "point_hg_zero": (None, ref_utils.HenyeyGreenstein(0.0), None),
        "partial_forms": (functools.partial(ref_light.rectangular_mask, 0.5, 2.5), functools.partial(ref_utils.cone, 0.9),
                          None),
    }


LIGHT_LOCATION, LIGHT_ROTATION = (0.3, -0.2, 1.0), (0.6, (1.0, 1.0, 0.0))


def make_emission():
    out = {}
    for k, (name, (position, direction, wavelength)) in enumerate(emission_cases().items()):
        world = Node(name="world", geometry=Sphere(radius=50.0, material=Material(refractive_index=1.0)))
        kwargs = {}
        if position is not None:
            kwargs["position"] = position
        if direction is not None:
            kwargs["direction"] = direction
        if wavelength is not None:
            kwargs["wavelength"] = wavelength
        node = Node(name="light", parent=world, light=ref_light.Light(**kwargs))
        node.location = LIGHT_LOCATION
        node.rotate(LIGHT_ROTATION[0], np.asarray(LIGHT_ROTATION[1]) / np.linalg.norm(LIGHT_ROTATION[1]))
        n = 100_000 if name != "partial_forms" else 30_000  # (unrecognised partials take the reference's per-ray path)
        np.random.seed(100 + k)
        pos, dirs, wl, _ = emit_bundle(Scene(world), n)
        rows = np.column_stack((pos, dirs, wl))
        out[f"{name}_n"] = np.int64(n)
        out[f"{name}_quantiles"] = np.quantile(rows, QUANTILES, axis=0)  # [201, 7]
        out[f"{name}_second_moments"] = rows.T @ rows / n                # [7, 7]
    np.savez_compressed(os.path.join(HERE, "emission.npz"), probabilities=QUANTILES, **out)


# ---- surface delegates called directly -----------------------------------------------------------------------
FACES = {"left": (-1, 0, 0), "right": (1, 0, 0), "near": (0, -1, 0), "far": (0, 1, 0), "bottom": (0, 0, -1), "top": (0, 0, 1)}


def probe_rays(size, rng, per_face=40):
    """Rays ON every face of a box of `size`, leaving (from inside) and arriving (from outside), mixed wavelengths."""
    half = 0.5 * np.asarray(size, dtype=float)
    rows = []
    for normal in FACES.values():
        normal = np.asarray(normal, dtype=float)
        axis = int(np.argmax(np.abs(normal)))
        for _ in range(per_face):
            p = rng.uniform(-0.95, 0.95, 3) * half
            p[axis] = normal[axis] * half[axis]
            d = rng.normal(size=3)
            d /= np.linalg.norm(d)
            if abs(d @ normal) < 0.05:
                d = (d + normal) / np.linalg.norm(d + normal)
            rows.append((p, d, rng.uniform(420.0, 680.0), 1.0 if d @ normal > 0 else 0.0))
    pos = np.array([r[0] for r in rows]); dirs = np.array([r[1] for r in rows])
    return pos, dirs, np.array([r[2] for r in rows]), np.array([r[3] for r in rows])


def call_delegate(delegate, geometry, inner, outer, pos, dirs, wl, leaving, lambertian_seed=None):
    """(reflectivity, reflected direction, transmitted direction) of a reference delegate, ray by ray.  NaN rows where
    the reference gives no direction (total internal reflection: fresnel_refraction has no real root)."""
    n = len(wl)
    refl, r_dir, t_dir = np.zeros(n), np.full((n, 3), np.nan), np.full((n, 3), np.nan)
    for i in range(n):
        ray = Ray(position=tuple(pos[i]), direction=tuple(dirs[i]), wavelength=float(wl[i]))
        container, adjacent = (inner, outer) if leaving[i] else (outer, inner)
        refl[i] = delegate.reflectivity(None, ray, geometry, container, adjacent)
        if lambertian_seed is not None:
            np.random.seed(lambertian_seed + i)
        r_dir[i] = delegate.reflected_direction(None, ray, geometry, container, adjacent)
        if refl[i] < 1.0:
            with np.errstate(invalid="ignore"):
                t_dir[i] = delegate.transmitted_direction(None, ray, geometry, container, adjacent)
    return refl, r_dir, t_dir


def make_lsc_delegates():
    rng = np.random.default_rng(11)
    size = (5.0, 5.0, 1.0)
    out = {}
    for tag, cells, mirror in (("cells_and_mirror", {"left", "right", "near", "far"}, True), ("two_cells", {"left", "far"}, False),
                               ("bare", set(), False)):
        lsc = ref_lsc.LSC(size)
        if cells:
            lsc.add_solar_cell(cells)
        if mirror:
            lsc.add_back_surface_mirror()
        lsc._make_scene()
        world = lsc._scene.root
        node = next(n for n in world.children if n.name == "LSC")
        delegate = node.geometry.material.surface.delegate
        assert isinstance(delegate, ref_lsc.OptionalMirrorAndSolarCell)
        pos, dirs, wl, leaving = probe_rays(size, rng)
        refl, r_dir, t_dir = call_delegate(delegate, node.geometry, node, world, pos, dirs, wl, leaving)
        out.update({f"{tag}_pos": pos, f"{tag}_dir": dirs, f"{tag}_wl": wl, f"{tag}_leaving": leaving,
                    f"{tag}_R": refl, f"{tag}_reflected": r_dir, f"{tag}_transmitted": t_dir})
    # the air-gap mirror below the LSC: reflectivity 1 everywhere; its "specular" branch returns the REFRACTED direction
    # (lsc.py:74-78, a known reference bug), so only the reflectivity is pinned
    lsc = ref_lsc.LSC(size)
    lsc.add_air_gap_mirror(lambertian=False)
    lsc._make_scene()
    world = lsc._scene.root
    mirror_node = next(n for n in world.children if n.name != "LSC" and n.geometry is not None)
    delegate = mirror_node.geometry.material.surface.delegate
    assert isinstance(delegate, ref_lsc.AirGapMirror)
    msize = mirror_node.geometry.size if hasattr(mirror_node.geometry, "size") else mirror_node.geometry._size
    pos, dirs, wl, leaving = probe_rays(msize, rng, per_face=10)
    refl = np.array([delegate.reflectivity(None, Ray(position=tuple(p), direction=tuple(d), wavelength=float(w)),
                                           mirror_node.geometry, world, mirror_node) for p, d, w in zip(pos, dirs, wl)])
    out.update({"air_gap_R": refl, "air_gap_size": np.asarray(msize, dtype=float),
                "air_gap_location": np.asarray(mirror_node.location, dtype=float)})
    np.savez_compressed(os.path.join(HERE, "lsc_delegates.npz"), **out)


# ---- coatings (examples/006 Coatings.ipynb cell 3 and a spectral variant) ------------------------------------------
COATING_X = np.array([450.0, 500.0, 560.0, 600.0, 650.0])
COATING_R = np.array([0.10, 0.25, 0.55, 0.80, 0.90])


class PartialTopSurfaceMirror(FresnelSurfaceDelegate):
    """The notebook's delegate: a perfect mirror on the quarter x > 0, y > 0 of the top surface."""

    def reflectivity(self, surface, ray, geometry, container, adjacent):
        normal = geometry.normal(ray.position)
        if np.allclose(normal, (0, 0, 1)):
            x, y = ray.position[0], ray.position[1]
            if x > 0 and y > 0:
                return 1.0
        return super(PartialTopSurfaceMirror, self).reflectivity(surface, ray, geometry, container, adjacent)


class SpectralBottomCoating(PartialTopSurfaceMirror):
    """... plus a coating on the bottom surface whose reflectivity is a spectrum."""

    def reflectivity(self, surface, ray, geometry, container, adjacent):
        normal = geometry.normal(ray.position)
        if np.allclose(normal, (0, 0, -1)):
            return float(np.interp(ray.wavelength, COATING_X, COATING_R))
        return super(SpectralBottomCoating, self).reflectivity(surface, ray, geometry, container, adjacent)


def coated_disc_scene():
    x = np.linspace(440.0, 660.0, 111)
    world = Node(name="world", geometry=Sphere(radius=10.0, material=Material(refractive_index=1.0)))
    Node(name="disc", parent=world, geometry=Cylinder(length=1.0, radius=3.0, material=Material(
        refractive_index=1.5, surface=Surface(delegate=SpectralBottomCoating()),
        components=[Absorber(coefficient=0.3, name="grey")])))
    light = Node(name="lamp", parent=world, light=ref_light.Light(
        position=ref_light.RectangularMask(2.0, 2.0), direction=ref_utils.Cone(0.3),
        wavelength=ref_light.SpectrumWavelengthMask(Distribution(x, lamp(x)))))
    light.location = (0.0, 0.0, 3.0)
    light.rotate(np.radians(180), (1, 0, 0))
    return Scene(world)


EVENTS = [Event.REFLECT, Event.TRANSMIT, Event.ABSORB, Event.NONRADIATIVE, Event.EXIT, Event.KILL]


def make_coatings():
    rng = np.random.default_rng(23)
    size = (10.0, 10.0, 1.0)
    world = Node(name="world", geometry=Box((15.0, 15.0, 15.0), material=Material(refractive_index=1.0)))
    slab = Node(name="slab", parent=world, geometry=Box(size, material=Material(
        refractive_index=1.5, surface=Surface(delegate=SpectralBottomCoating()))))
    pos, dirs, wl, leaving = probe_rays(size, rng, per_face=60)
    refl, r_dir, t_dir = call_delegate(slab.geometry.material.surface.delegate, slab.geometry, slab, world, pos, dirs, wl,
                                       leaving)
    out = {"box_pos": pos, "box_dir": dirs, "box_wl": wl, "box_leaving": leaving, "box_R": refl, "box_reflected": r_dir,
           "box_transmitted": t_dir, "coating_x": COATING_X, "coating_R": COATING_R}
    # the reference's Python tracer through a disc with both coatings (cylinders need no trimesh)
    scene = coated_disc_scene()
    n = 8000
    np.random.seed(77)
    counts = {e: np.zeros(n) for e in EVENTS}
    down = np.zeros(n)
    for i, ray in enumerate(scene.emit(n)):
        last = None
        for step_ray, event in photon_tracer.follow(scene, ray):
            if event in counts:
                counts[event][i] += 1
            last = (step_ray, event)
        down[i] = 1.0 if last[1] == Event.EXIT and last[0].direction[2] < 0 else 0.0
    out["disc_n"] = np.int64(n)
    for e in EVENTS:
        out[f"disc_mean_{e.name}"] = counts[e].mean()
        out[f"disc_var_{e.name}"] = counts[e].var(ddof=1)
    out["disc_exits_downwards"] = down.mean()
    np.savez_compressed(os.path.join(HERE, "coatings.npz"), **out)


if __name__ == "__main__":
    make_emission()
    make_lsc_delegates()
    make_coatings()
    print("wrote emission.npz, lsc_delegates.npz, coatings.npz")
