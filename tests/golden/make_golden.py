"""Generate the golden fixtures in tests/golden/ by running the UNMODIFIED reference from /root/reference.

Run in the build container only (the GPU box has no reference tree):   python tests/golden/make_golden.py

Writes
  optics.npz       fresnel_reflectivity / specular_reflection / fresnel_refraction of pvtrace/material/utils.py:8-45
                   on seeded grids; phase functions' deterministic cores are covered through trace logs below
  geometry.npz     Sphere.intersections / ray_z_cylinder / Cylinder.normal / Sphere.normal of pvtrace/geometry on
                   seeded random rays (the Box goes through trimesh in the reference and is pinned via the engine logs)
  distribution.npz Distribution.__call__/lookup/sample (pvtrace/material/distribution.py) on a Gaussian spectrum
  tables_<scene>.npz   every array of pvtrace.engine.compiler.CompiledScene for scenes built with the REFERENCE classes
  engine_<scene>.npz   pvtrace.engine._kernel.trace_bundle (compiled from the reference .pyx into oracle/_ref) full
                       event logs + tallies for seeded input rays: the vectors the oracle must reproduce bit for bit
  python_tracer_<scene>.npz  per-ray event counts of pvtrace.algorithm.photon_tracer.follow (sphere / cylinder scenes)
"""
import functools
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_loader  # noqa: E402

ref_loader.load_reference_package()
kernel = ref_loader.load_ref_kernel()
assert kernel is not None, "build the reference kernel first: make -C oracle ref"

from pvtrace.algorithm import photon_tracer  # noqa: E402
from pvtrace.engine.compiler import compile_scene  # noqa: E402
from pvtrace.engine.recorder import Heatmap, Histogram, Recorder  # noqa: E402
from pvtrace.geometry.box import Box  # noqa: E402
from pvtrace.geometry.cylinder import Cylinder  # noqa: E402
from pvtrace.geometry.sphere import Sphere  # noqa: E402
from pvtrace.geometry.utils import ray_z_cylinder  # noqa: E402
from pvtrace.light.event import Event  # noqa: E402
from pvtrace.light.light import Light  # noqa: E402
from pvtrace.material.component import Absorber, Luminophore, Reactor, Scatterer  # noqa: E402
from pvtrace.material.distribution import Distribution  # noqa: E402
from pvtrace.material.material import Material  # noqa: E402
from pvtrace.material.surface import NullSurfaceDelegate, Surface  # noqa: E402
from pvtrace.material.utils import (  # noqa: E402
    Cone, HenyeyGreenstein, cone, fresnel_reflectivity, fresnel_refraction, gaussian, specular_reflection)
from pvtrace.scene.node import Node  # noqa: E402
from pvtrace.scene.scene import Scene  # noqa: E402
from pvtrace.data import lumogen_f_red_305  # noqa: E402


def unit(v):
    v = np.asarray(v, dtype=float)
    return v / np.linalg.norm(v, axis=-1, keepdims=True)


def make_optics():
    rng = np.random.default_rng(1)
    n = 4000
    n1 = rng.choice([1.0, 1.33, 1.5, 2.0], n)
    n2 = rng.choice([1.0, 1.33, 1.5, 1.7], n)
    angle = rng.uniform(0.0, np.pi / 2, n)
    angle[:8] = [0.0, 1e-9, np.pi / 4, np.pi / 2 - 1e-9, 0.7, 0.72, 0.73, 0.8]
    R = np.array([fresnel_reflectivity(a, x, y) for a, x, y in zip(angle, n1, n2)])
    d = unit(rng.normal(size=(n, 3)))
    nrm = unit(rng.normal(size=(n, 3)))
    refl = np.array([specular_reflection(a, b) for a, b in zip(d, nrm)])
    # refraction is only evaluated where it is defined (no total internal reflection); normal flipped along the ray
    # first, as FresnelSurfaceDelegate.transmitted_direction does (surface.py:165-177)
    nf = np.where((np.sum(d * nrm, axis=1) < 0)[:, None], -nrm, nrm)
    cosi = np.sum(d * nf, axis=1)
    ok = 1.0 - (n1 / n2) ** 2 * (1.0 - cosi ** 2) > 0.0
    refr = np.zeros((n, 3))
    refr[ok] = np.array([fresnel_refraction(a, b, x, y) for a, b, x, y in zip(d[ok], nf[ok], n1[ok], n2[ok])])
    np.savez_compressed(os.path.join(HERE, "optics.npz"), angle=angle, n1=n1, n2=n2, R=R, d=d, nrm=nrm, refl=refl,
                        refr=refr, refr_ok=ok)


def make_geometry():
    rng = np.random.default_rng(2)
    n = 2000
    origins = rng.uniform(-3, 3, size=(n, 3))
    dirs = unit(rng.normal(size=(n, 3)))
    sphere = Sphere(radius=1.3)
    cyl_len, cyl_rad = 2.0, 0.7
    sph_pts = np.full((n, 2, 3), np.nan)
    sph_cnt = np.zeros(n, dtype=np.int32)
    cyl_pts = np.full((n, 4, 3), np.nan)
    cyl_cnt = np.zeros(n, dtype=np.int32)
    for i in range(n):
        pts = sphere.intersections(tuple(origins[i]), tuple(dirs[i]))
        sph_cnt[i] = len(pts)
        for k, p in enumerate(pts):
            sph_pts[i, k] = p
        pts, _ = ray_z_cylinder(cyl_len, cyl_rad, tuple(origins[i]), tuple(dirs[i]))
        cyl_cnt[i] = len(pts)
        for k, p in enumerate(pts):
            cyl_pts[i, k] = p
    cyl = Cylinder(length=cyl_len, radius=cyl_rad)
    surf = []
    for i in range(n):
        for k in range(cyl_cnt[i]):
            surf.append(cyl_pts[i, k])
    surf = np.array(surf[:1500])
    cyl_nrm = np.array([cyl.normal(tuple(p)) for p in surf])
    sph_surf = unit(rng.normal(size=(500, 3))) * 1.3
    sph_nrm = np.array([sphere.normal(tuple(p)) for p in sph_surf])
    np.savez_compressed(os.path.join(HERE, "geometry.npz"), origins=origins, dirs=dirs, sphere_radius=1.3,
                        sph_pts=sph_pts, sph_cnt=sph_cnt, cyl_len=cyl_len, cyl_rad=cyl_rad, cyl_pts=cyl_pts,
                        cyl_cnt=cyl_cnt, cyl_surf=surf, cyl_nrm=cyl_nrm, sph_surf=sph_surf, sph_nrm=sph_nrm)


def make_distribution():
    x = np.linspace(300.0, 1000.0, 200)
    y = gaussian(x, 1.0, 600.0, 40.0)
    dist = Distribution(x, y)
    rng = np.random.default_rng(3)
    xq = rng.uniform(300.0, 1000.0, 500)
    pq = rng.uniform(0.0, 1.0, 500)
    np.savez_compressed(os.path.join(HERE, "distribution.npz"), x=x, y=y, cdf=dist._cdf, xq=xq, pq=pq,
                        value=np.array([dist(v) for v in xq]), lookup=np.array([dist.lookup(v) for v in xq]),
                        sample=np.array([dist.sample(v) for v in pq]))


# ---- scenes built with the REFERENCE classes (mirrors of tests/scenes.py, which uses pvtrace_b200 classes) ----

def scene_fresnel():
    """tests/test_engine.py:36-52 (make_fresnel_scene) + recorders"""
    world = Node(name="world", geometry=Sphere(radius=10.0, material=Material(refractive_index=1.0)))
    box = Node(name="box", geometry=Box((1.0, 1.0, 1.0), material=Material(refractive_index=1.5)), parent=world)
    box.location = (0.0, 0.0, 2.0)
    Node(name="light", light=Light(direction=functools.partial(cone, np.pi / 16)), parent=world)
    world.recorders = [Recorder("exit", event="exit", histograms=[Histogram("angle", 0.0, 1.6, 16)])]
    box.recorders = [Recorder("in", event="entering", histograms=[Heatmap("x", "y", (-0.5, 0.5, 8), (-0.5, 0.5, 8))]),
                     Recorder("out-top", event="escaping", facet=(0, 0, 1)),
                     Recorder("refl", event="reflected")]
    return Scene(world)


def scene_lsc():
    """tests/test_engine.py:55-96 (make_lsc_scene) + recorders"""
    x = np.linspace(300.0, 1000.0, 200)
    absorption = np.column_stack((x, 5.0 * gaussian(x, 1.0, 480.0, 40.0)))
    emission = np.column_stack((x, gaussian(x, 1.0, 600.0, 40.0)))
    world = Node(name="world", geometry=Sphere(radius=10.0, material=Material(refractive_index=1.0)))
    slab = Node(name="slab", parent=world, geometry=Box((5.0, 5.0, 1.0), material=Material(
        refractive_index=1.5, components=[
            Luminophore(coefficient=absorption, emission=emission, quantum_yield=0.9, name="dye"),
            Absorber(coefficient=0.3, name="background")])))
    light = Node(name="light", light=Light(), parent=world)
    light.location = (0.0, 0.0, -3.0)
    world.recorders = [Recorder("exit", event="exit", histograms=[Histogram("wavelength", 300.0, 1000.0, 70)])]
    slab.recorders = [Recorder("lost", event="lost", histograms=[Histogram("pathlength", 0.0, 20.0, 40)]),
                      Recorder("edge+x", event="escaping", facet=(1, 0, 0),
                               histograms=[Heatmap("y", "z", (-2.5, 2.5, 10), (-0.5, 0.5, 4))]),
                      Recorder("top", event="escaping", facet=(0, 0, 1), histograms=[Histogram("angle", 0.0, 1.6, 16)]),
                      Recorder("entering", event="entering")]
    return Scene(world)


def scene_mixed():
    """every primitive, component type, phase function and surface tag the engine supports, posed off-axis"""
    x = np.linspace(350.0, 900.0, 111)
    world = Node(name="world", geometry=Box((30.0, 30.0, 30.0), material=Material(refractive_index=1.0)))
    cyl = Node(name="cyl", parent=world, geometry=Cylinder(length=3.0, radius=1.0, material=Material(
        refractive_index=1.4, components=[
            Scatterer(coefficient=0.4, phase_function=HenyeyGreenstein(0.6), name="hg"),
            Luminophore(coefficient=np.column_stack((x, 2.0 * gaussian(x, 1.0, 500.0, 50.0))),
                        emission=np.column_stack((x, gaussian(x, 1.0, 620.0, 35.0))), quantum_yield=0.8,
                        phase_function=Cone(0.5), name="lum")])))
    cyl.translate((0.5, -0.3, 4.0))
    cyl.rotate(0.7, (1.0, 0.3, 0.0))
    ball = Node(name="ball", parent=world, geometry=Sphere(radius=1.2, material=Material(
        refractive_index=1.6, components=[Reactor(coefficient=0.5, name="react"),
                                          Absorber(coefficient=0.2, name="abs")])))
    ball.translate((-2.0, 1.0, 6.5))
    ghost = Node(name="ghost", parent=world, geometry=Box((2.0, 2.0, 0.5), material=Material(
        refractive_index=1.0, surface=Surface(delegate=NullSurfaceDelegate()),
        components=[Scatterer(coefficient=1.0, quantum_yield=0.7, name="fog")])))
    ghost.translate((1.0, 1.5, 2.0))
    ghost.rotate(0.4, (0.0, 1.0, 0.2))
    Node(name="light", parent=world, light=Light(direction=functools.partial(cone, 0.35)))
    world.recorders = [Recorder("exit", event="exit"), Recorder("killed", event="killed")]
    cyl.recorders = [Recorder("cyl-in", event="entering"), Recorder("cyl-out", event="escaping"),
                     Recorder("cyl-lost", event="lost")]
    ball.recorders = [Recorder("ball-react", event="reacted"), Recorder("ball-refl", event="reflected")]
    ghost.recorders = [Recorder("ghost-in", event="entering"), Recorder("ghost-lost", event="lost")]
    return Scene(world)


def scene_hello_world():
    world = Node(name="world", geometry=Sphere(radius=10.0, material=Material(refractive_index=1.0)))
    ball = Node(name="ball-lens", parent=world, geometry=Sphere(radius=1.0, material=Material(refractive_index=1.5)))
    ball.location = (0, 0, 2)
    Node(name="green-laser", parent=world, light=Light(direction=functools.partial(cone, np.pi / 8), name="green-laser"))
    return Scene(world)


def scene_nested_cylinders():
    world = Node(name="World", geometry=Sphere(radius=10.0, material=Material(refractive_index=1.0)))
    a = Node(name="A", parent=world, geometry=Cylinder(length=2, radius=0.5, material=Material(refractive_index=1.5)))
    a.translate((0, 0, 2))
    a.rotate(np.pi * 0.2, (0, 1, 0))
    b = Node(name="B", parent=a, geometry=Cylinder(length=2.0, radius=0.4, material=Material(refractive_index=1.5)))
    b.rotate(np.pi / 2, (1, 0, 0))
    light = Node(name="Light (555nm)", parent=world, light=Light(direction=functools.partial(cone, np.radians(30))))
    light.translate((0, 0, -1))
    return Scene(world)


def scene_lsc_device():
    """LSC((5,5,1)) default tables: pvtrace/device/lsc.py:115-219 with a plain Fresnel surface (the reference
    engine rejects the LSC's delegate subclass, compiler.py:239-247)"""
    x = np.arange(400, 800)
    world = Node(name="World", geometry=Box((500.0, 500.0, 100.0), material=Material(refractive_index=1.0)))
    Node(name="LSC", parent=world, geometry=Box((5.0, 5.0, 1.0), material=Material(refractive_index=1.5, components=[
        Luminophore(np.column_stack((x, lumogen_f_red_305.absorption(x) * 10.0)),
                    emission=np.column_stack((x, lumogen_f_red_305.emission(x))), quantum_yield=1.0,
                    phase_function=None, name="Lumogen F Red 305"),
        Absorber(0.1, name="Background")])))
    light = Node(name="Light", parent=world, light=Light(name="Light", direction=functools.partial(cone, np.radians(20))))
    light.location = (0.0, 0.0, 5.0)
    light.rotate(np.radians(180), (1, 0, 0))
    return Scene(world)


SCENES = {"fresnel": scene_fresnel, "lsc": scene_lsc, "mixed": scene_mixed, "hello_world": scene_hello_world,
          "nested_cylinders": scene_nested_cylinders, "lsc_device": scene_lsc_device}
TABLES = ("geom_type geom_params local_to_world world_to_local refractive_index surface_type comp_start comp_count "
          "comp_type comp_qy comp_tau_rad comp_tau_nr comp_phase_type comp_phase_param comp_abs_start comp_abs_n "
          "comp_ems_start comp_ems_n abs_x abs_y ems_x ems_cdf rec_node rec_event rec_has_facet rec_facet rec_atol "
          "rec_hist_start rec_hist_n hist_prop_a hist_prop_b hist_na hist_nb hist_lo_a hist_hi_a hist_lo_b hist_hi_b "
          "hist_offset").split()


def emit_rays(scene, n, seed):
    np.random.seed(seed)
    pos, dirs, wl = np.zeros((n, 3)), np.zeros((n, 3)), np.zeros(n)
    for i, ray in enumerate(scene.emit(n)):
        pos[i], dirs[i], wl[i] = ray.position, ray.direction, ray.wavelength
    return pos, dirs, wl


def make_engine_goldens():
    for name, build in SCENES.items():
        scene = build()
        compiled = compile_scene(scene)
        tables = {t: np.asarray(getattr(compiled, t)) for t in TABLES}
        tables["root_id"] = np.int64(compiled.root_id)
        tables["total_bins"] = np.int64(compiled.total_bins)
        tables["node_names"] = np.array(compiled.node_names)
        tables["component_names"] = np.array(compiled.component_names)
        tables["recorder_names"] = np.array(compiled.recorder_names)
        np.savez_compressed(os.path.join(HERE, f"tables_{name}.npz"), **tables)
        n, max_events = 96, 48
        pos, dirs, wl = emit_rays(scene, n, seed=5)
        for method, tag in ((0, "kT"), (1, "redshift"), (2, "full")):
            if tag != "kT" and name not in ("lsc", "mixed"):
                continue
            out = kernel.trace_bundle(compiled, pos, dirs, wl, 1234, 1000, max_events, method, 2, 1)
            big = kernel.trace_bundle(compiled, *emit_rays(scene, 20000, seed=6), 99, 1000, max_events, method, 4, 0)
            np.savez_compressed(os.path.join(HERE, f"engine_{name}_{tag}.npz"), positions=pos, directions=dirs,
                                wavelengths=wl, seed=np.int64(1234), max_events=np.int64(max_events),
                                emit_method=np.int64(method),
                                big_seed=np.int64(99), big_n=np.int64(20000), big_emit_seed=np.int64(6),
                                **{f"out_{k}": v for k, v in out.items()},
                                **{f"big_{k}": big[k] for k in ("rec_distinct", "rec_crossings", "rec_sums", "rec_bins")})


def make_python_tracer_goldens():
    """Per-ray event counts of the reference PYTHON tracer (sphere / cylinder scenes only: Box needs trimesh)."""
    kinds = [e for e in Event]
    for name, n in (("hello_world", 1500), ("nested_cylinders", 1500)):
        scene = SCENES[name]()
        np.random.seed(0)
        counts = np.zeros((n, len(kinds)), dtype=np.int32)
        for i, ray in enumerate(scene.emit(n)):
            for _, event in photon_tracer.follow(scene, ray):
                counts[i, event.value] += 1
        np.savez_compressed(os.path.join(HERE, f"python_tracer_{name}.npz"), counts=counts,
                            event_names=np.array([e.name for e in kinds]))


def make_yaml_goldens():
    """Tables (and light poses / emitted rays) of the scenes the REFERENCE's YAML front end (pvtrace/cli/parse.py)
    builds from the fixtures in tests/data/: what pvtrace_b200.cli.parse must reproduce."""
    import pvtrace as pkg  # the stub top-level package: give it the names pvtrace/cli/parse.py imports from it
    import pvtrace.geometry.mesh  # noqa: F401
    for name, mod in (("Scene", "pvtrace.scene.scene"), ("Node", "pvtrace.scene.node"), ("Box", "pvtrace.geometry.box"),
                      ("Mesh", "pvtrace.geometry.mesh"), ("Cylinder", "pvtrace.geometry.cylinder"),
                      ("Sphere", "pvtrace.geometry.sphere"), ("Material", "pvtrace.material.material"),
                      ("Absorber", "pvtrace.material.component"), ("Scatterer", "pvtrace.material.component"),
                      ("Luminophore", "pvtrace.material.component"), ("Light", "pvtrace.light.light")):
        __import__(mod)
        setattr(pkg, name, getattr(sys.modules[mod], name))
    pkg.MeshcatRenderer = object
    import pvtrace.cli.parse as refparse

    for stem in ("lsc_recorded", "primitives"):
        scene = refparse.parse(os.path.join(ROOT, "tests", "data", stem + ".yml"))
        compiled = compile_scene(scene)
        tables = {t: np.asarray(getattr(compiled, t)) for t in TABLES}
        tables["root_id"] = np.int64(compiled.root_id)
        tables["total_bins"] = np.int64(compiled.total_bins)
        tables["node_names"] = np.array(compiled.node_names)
        tables["component_names"] = np.array(compiled.component_names)
        tables["recorder_names"] = np.array(compiled.recorder_names)
        lights = scene.light_nodes
        tables["light_names"] = np.array([n.name for n in lights])
        tables["light_to_world"] = np.array([np.asarray(n.transformation_to(scene.root)) for n in lights])
        tables["light_delegates"] = np.array(["|".join(type(d).__name__ if not callable(getattr(d, "__name__", None)) else d.__name__
                                                        for d in (n.light.wavelength, n.light.position, n.light.direction))
                                              for n in lights])
        np.savez_compressed(os.path.join(HERE, f"yaml_{stem}.npz"), **tables)


if __name__ == "__main__":
    make_yaml_goldens()
    make_optics()
    make_geometry()
    make_distribution()
    make_engine_goldens()
    make_python_tracer_goldens()
    print("golden fixtures written to", HERE)
