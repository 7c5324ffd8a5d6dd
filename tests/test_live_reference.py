"""Randomised, LIVE comparison of the flattener with the reference's: the same seeded random scene is built twice --
with pvtrace_b200's host classes and with the unmodified reference's (imported from /root/reference through
oracle/ref_loader.py) -- and every CompiledScene table must agree (pvtrace/engine/compiler.py:57-204).  Complements
tests/test_golden_tables.py (six fixed scenes, committed fixtures): nested nodes, rotated boxes, every component type
and phase function, tabulated and constant spectra, recorders with facets / histograms / heat maps.  The oracle then
traces random rays through the same scenes beside the compiled reference kernel (oracle/_ref): bit-identical logs.

Runs only where the reference tree is present (the build container); skipped on the GPU box."""
import functools
import types

import numpy as np
import pytest

from oracle import ref_loader
from tests import scenes

pytestmark = pytest.mark.skipif(not ref_loader.reference_available(), reason="reference tree not present")


def ours():
    import pvtrace_b200 as pv
    from pvtrace_b200.engine import Heatmap, Histogram, Recorder
    from pvtrace_b200.material.utils import Cone, HenyeyGreenstein, gaussian

    return types.SimpleNamespace(
        Node=pv.Node, Scene=pv.Scene, Box=pv.Box, Sphere=pv.Sphere, Cylinder=pv.Cylinder, Material=pv.Material,
        Absorber=pv.Absorber, Scatterer=pv.Scatterer, Reactor=pv.Reactor, Luminophore=pv.Luminophore, Light=pv.Light,
        Surface=pv.Surface, NullSurfaceDelegate=pv.NullSurfaceDelegate, cone=pv.cone, Cone=Cone,
        HenyeyGreenstein=HenyeyGreenstein, gaussian=gaussian, Recorder=Recorder, Histogram=Histogram, Heatmap=Heatmap,
        compile_scene=pv.engine.compile_scene)


def reference():
    ref_loader.load_reference_package()
    from pvtrace.engine.compiler import compile_scene
    from pvtrace.engine.recorder import Heatmap, Histogram, Recorder
    from pvtrace.geometry.box import Box
    from pvtrace.geometry.cylinder import Cylinder
    from pvtrace.geometry.sphere import Sphere
    from pvtrace.light.light import Light
    from pvtrace.material.component import Absorber, Luminophore, Reactor, Scatterer
    from pvtrace.material.material import Material
    from pvtrace.material.surface import NullSurfaceDelegate, Surface
    from pvtrace.material.utils import Cone, HenyeyGreenstein, cone, gaussian
    from pvtrace.scene.node import Node
    from pvtrace.scene.scene import Scene

    return types.SimpleNamespace(**{k: v for k, v in locals().items()})


SURFACE_EVENTS = ("entering", "escaping", "reflected")
VOLUME_EVENTS = ("lost", "reacted", "killed")
PROPS = ("wavelength", "angle", "pathlength", "x", "y", "z")


def random_scene(ns, seed):
    """One scene from `seed`, built with the classes of namespace `ns` (identical calls for both libraries)."""
    rng = np.random.default_rng(seed)
    if rng.random() < 0.5:
        world = ns.Node(name="world", geometry=ns.Sphere(radius=25.0, material=ns.Material(refractive_index=1.0)))
    else:
        world = ns.Node(name="world", geometry=ns.Box((60.0, 50.0, 40.0), material=ns.Material(refractive_index=1.0)))
    world.recorders = [ns.Recorder("exit", event="exit", histograms=[ns.Histogram("wavelength", 300.0, 900.0, 30)])]
    x = np.linspace(320.0, 880.0, int(rng.integers(20, 150)))

    def spectrum(scale, centre):
        return np.column_stack((x, scale * ns.gaussian(x, 1.0, centre, float(rng.uniform(20.0, 60.0)))))

    def phase():
        pick = rng.integers(0, 3)
        if pick == 0:
            return None
        return ns.HenyeyGreenstein(float(rng.uniform(-0.8, 0.8))) if pick == 1 else ns.Cone(float(rng.uniform(0.1, 1.2)))

    def components(tag):
        out = []
        for c in range(int(rng.integers(0, 4))):
            kind = rng.integers(0, 4)
            name = f"{tag}-c{c}"
            coeff = float(rng.uniform(0.05, 3.0)) if rng.random() < 0.5 else spectrum(float(rng.uniform(0.5, 8.0)), float(rng.uniform(400, 600)))
            if kind == 0:
                out.append(ns.Absorber(coefficient=coeff, name=name))
            elif kind == 1:
                out.append(ns.Scatterer(coefficient=coeff, quantum_yield=float(rng.uniform(0.3, 1.0)), phase_function=phase(), name=name))
            elif kind == 2:
                out.append(ns.Reactor(coefficient=coeff, name=name))
            else:
                out.append(ns.Luminophore(coefficient=spectrum(float(rng.uniform(1.0, 9.0)), float(rng.uniform(420, 560))),
                                          emission=spectrum(1.0, float(rng.uniform(560, 760))),
                                          quantum_yield=float(rng.uniform(0.5, 1.0)), phase_function=phase(), name=name))
        return out

    def geometry(tag):
        kind = rng.integers(0, 3)
        surface = ns.Surface(delegate=ns.NullSurfaceDelegate()) if rng.random() < 0.2 else None
        kw = dict(refractive_index=float(rng.uniform(1.1, 1.9)), components=components(tag))
        if surface is not None:
            kw["surface"] = surface
        material = ns.Material(**kw)
        if kind == 0:
            return "box", ns.Box(tuple(float(v) for v in rng.uniform(0.5, 3.0, 3)), material=material)
        if kind == 1:
            return "sphere", ns.Sphere(radius=float(rng.uniform(0.3, 1.5)), material=material)
        return "cylinder", ns.Cylinder(length=float(rng.uniform(0.5, 3.0)), radius=float(rng.uniform(0.2, 1.0)), material=material)

    def recorders(tag, kind):
        out = []
        for r in range(int(rng.integers(0, 4))):
            hists = []
            for _ in range(int(rng.integers(0, 3))):
                if rng.random() < 0.6:
                    hists.append(ns.Histogram(str(rng.choice(PROPS)), float(rng.uniform(-1.0, 0.0)), float(rng.uniform(1.0, 900.0)), int(rng.integers(2, 40))))
                else:
                    hists.append(ns.Heatmap("x", str(rng.choice(("y", "z"))), (-2.0, 2.0, int(rng.integers(2, 12))), (-1.0, 1.5, int(rng.integers(2, 12)))))
            if rng.random() < 0.6:
                event, facet = str(rng.choice(SURFACE_EVENTS)), None
                if kind == "box" and rng.random() < 0.5:
                    axis, sign = int(rng.integers(0, 3)), float(rng.choice((-1.0, 1.0)))
                    facet = tuple(sign if k == axis else 0.0 for k in range(3))
                out.append(ns.Recorder(f"{tag}-r{r}", event=event, facet=facet, histograms=hists))
            else:
                out.append(ns.Recorder(f"{tag}-r{r}", event=str(rng.choice(VOLUME_EVENTS)), histograms=hists))
        return out

    parents = [world]
    for k in range(int(rng.integers(1, 6))):
        tag = f"n{k}"
        kind, geom = geometry(tag)
        parent = parents[int(rng.integers(0, len(parents)))] if rng.random() < 0.3 else world
        node = ns.Node(name=tag, parent=parent, geometry=geom)
        node.translate(tuple(float(v) for v in rng.uniform(-4.0, 4.0, 3)))
        if rng.random() < 0.6:
            axis = rng.normal(size=3)
            node.rotate(float(rng.uniform(-3.0, 3.0)), tuple(float(v) for v in axis / np.linalg.norm(axis)))
        node.recorders = recorders(tag, kind)
        parents.append(node)
    light = ns.Node(name="light", parent=world,
                    light=ns.Light(direction=functools.partial(ns.cone, float(rng.uniform(0.05, 1.0)))) if rng.random() < 0.7 else ns.Light())
    light.translate(tuple(float(v) for v in rng.uniform(-6.0, 6.0, 3)))
    return ns.Scene(world)


@pytest.mark.parametrize("seed", range(60))
def test_random_scene_tables_equal_the_reference(seed):
    got = ours().compile_scene(random_scene(ours(), seed))
    want = reference().compile_scene(random_scene(reference(), seed))
    for table in scenes.TABLES:
        g, w = np.asarray(getattr(got, table)), np.asarray(getattr(want, table))
        assert g.shape == w.shape, (table, g.shape, w.shape)
        assert g.dtype == w.dtype, (table, g.dtype, w.dtype)
        if np.issubdtype(w.dtype, np.integer):
            assert (g == w).all(), table
        else:
            np.testing.assert_allclose(g, w, rtol=1e-13, atol=1e-13, err_msg=table)
    assert got.root_id == want.root_id and got.total_bins == want.total_bins
    assert list(got.node_names) == list(want.node_names)
    assert list(got.component_names) == list(want.component_names)
    assert list(got.recorder_names) == list(want.recorder_names)


# ---- the oracle against the compiled reference kernel on the same random scenes -------------------------------------
# (tests/test_oracle_pinning.py does this on the six fixed scenes; here: overlapping and nested primitives at random
# poses, every component type, rays from all over the scene -- bit-identical histories are required)
INT_KEYS = ("counts", "kind", "hit", "container", "adjacent", "component", "source", "rec_distinct", "rec_crossings",
            "rec_bins")
FLOAT_KEYS = ("position", "direction", "normal", "wavelength", "travelled", "duration")


@pytest.mark.skipif(ref_loader.load_ref_kernel() is None, reason="oracle/_ref not built on this machine")
@pytest.mark.parametrize("seed", range(30))
def test_oracle_traces_random_scenes_like_the_reference_kernel(seed):
    from oracle import pvt_oracle
    from pvtrace_b200.engine import _cuda

    kernel = ref_loader.load_ref_kernel()
    compiled = ours().compile_scene(random_scene(ours(), seed))
    rng = np.random.default_rng(1000 + seed)
    n = 3000
    pos = rng.uniform(-6.0, 6.0, (n, 3))
    direction = rng.normal(size=(n, 3))
    direction /= np.linalg.norm(direction, axis=1)[:, None]
    wl = rng.uniform(350.0, 850.0, n)
    method = seed % 3  # kT, redshift, full
    want = kernel.trace_bundle(compiled, pos, direction, wl, 99 + seed, 300, 48, method, 2, 1)
    got = pvt_oracle.trace_bundle(compiled, pos, direction, wl, 99 + seed, 300, 48, method, 2, 1, rng_mode=_cuda.RNG_XOSHIRO)
    for key in INT_KEYS:
        assert (got[key] == want[key]).all(), key
    for key in FLOAT_KEYS:
        np.testing.assert_allclose(got[key], want[key], rtol=1e-12, atol=1e-22, err_msg=key)
    np.testing.assert_allclose(got["rec_sums"], want["rec_sums"], rtol=1e-10)
    assert got["counts"].max() >= 3


@pytest.mark.skipif(ref_loader.load_ref_kernel() is None, reason="oracle/_ref not built on this machine")
@pytest.mark.parametrize("seed", range(12))
def test_addressed_philox_draws_give_the_reference_statistics_on_random_scenes(seed):
    """The product's random numbers are Philox draws ADDRESSED by (photon, step, purpose) -- not the reference's
    sequential xoshiro stream -- so histories differ ray by ray and must agree in distribution: every recorder's
    distinct-ray and crossing counts of the oracle in Philox mode (the CUDA kernels' twin, paired with them ray by ray in
    tests/test_gpu_trace_parity.py) against the compiled reference kernel on the same rays, two independent samples of
    the same binomial: |a - b| <= 4 (sqrt(a + b) + 1) -- conservative, the two runs share their rays and much of a ray's
    fate is geometry (observed over the 12 scenes: max 1.2, rms 0.4).  Random scenes: every primitive, component type and
    phase function."""
    from oracle import pvt_oracle
    from pvtrace_b200.engine import _cuda

    kernel = ref_loader.load_ref_kernel()
    compiled = ours().compile_scene(random_scene(ours(), 100 + seed))
    rng = np.random.default_rng(2000 + seed)
    n = 60000
    pos = rng.uniform(-6.0, 6.0, (n, 3))
    direction = rng.normal(size=(n, 3))
    direction /= np.linalg.norm(direction, axis=1)[:, None]
    wl = rng.uniform(350.0, 850.0, n)
    method = seed % 3
    want = kernel.trace_bundle(compiled, pos, direction, wl, 7 + seed, 500, 48, method, 4, 0)
    got = pvt_oracle.trace_bundle(compiled, pos, direction, wl, 7 + seed, 500, 48, method, 4, 0, rng_mode=_cuda.RNG_PHILOX)
    for key in ("rec_distinct", "rec_crossings"):
        a, b = got[key].astype(np.float64), want[key].astype(np.float64)
        z = np.abs(a - b) / (np.sqrt(a + b) + 1.0)
        assert z.max() <= 4.0, (key, list(compiled.recorder_names)[int(z.argmax())], a[int(z.argmax())], b[int(z.argmax())])
    # histogram bins with enough counts, pooled the same way
    a, b = got["rec_bins"].astype(np.float64), want["rec_bins"].astype(np.float64)
    busy = a + b >= 60
    if busy.any():
        z = np.abs(a[busy] - b[busy]) / np.sqrt(a[busy] + b[busy])
        assert z.max() <= 4.5 and (z > 3.0).mean() < 0.02, (z.max(), (z > 3.0).mean(), int(busy.sum()))
    assert int(got["rec_distinct"].sum()) > n // 2  # the recorders really see the rays


# ---- the host API's geometry against the reference's (Scene.intersections / Node.intersections, the sphere
# and cylinder roots and normals) -- Box goes through trimesh in the reference, absent here, and is pinned through the
# compiled kernel instead (tests above, tests/test_oracle_pinning.py) -----------------------------------------------
def round_scene(ns, seed):
    rng = np.random.default_rng(seed)
    world = ns.Node(name="world", geometry=ns.Sphere(radius=30.0, material=ns.Material(refractive_index=1.0)))
    parents = [world]
    for k in range(int(rng.integers(2, 7))):
        mat = ns.Material(refractive_index=float(rng.uniform(1.1, 1.9)))
        geom = ns.Sphere(radius=float(rng.uniform(0.3, 2.0)), material=mat) if rng.random() < 0.5 else \
            ns.Cylinder(length=float(rng.uniform(0.5, 4.0)), radius=float(rng.uniform(0.2, 1.5)), material=mat)
        parent = parents[int(rng.integers(0, len(parents)))] if rng.random() < 0.4 else world
        node = ns.Node(name=f"n{k}", parent=parent, geometry=geom)
        node.translate(tuple(float(v) for v in rng.uniform(-3.0, 3.0, 3)))
        if rng.random() < 0.7:
            axis = rng.normal(size=3)
            node.rotate(float(rng.uniform(-3.0, 3.0)), tuple(float(v) for v in axis / np.linalg.norm(axis)))
        parents.append(node)
    return ns.Scene(world)


@pytest.mark.parametrize("seed", range(20))
def test_scene_intersections_equal_the_reference(seed):
    """Scene.intersections (pvtrace/scene/scene.py:153-195): the same hits, in the same order, at the same points and
    distances for rays from all over a random scene of nested, rotated spheres and cylinders; and the outward normals of
    both geometries at the hit points (sphere.py:63-74, cylinder.py:52-65)."""
    a, b = round_scene(ours(), 500 + seed), round_scene(reference(), 500 + seed)
    rng = np.random.default_rng(600 + seed)
    hits = 0
    for _ in range(120):
        origin = tuple(float(v) for v in rng.uniform(-5.0, 5.0, 3))
        d = rng.normal(size=3)
        direction = tuple(float(v) for v in d / np.linalg.norm(d))
        got, want = a.intersections(origin, direction), b.intersections(origin, direction)
        assert [i.hit.name for i in got] == [i.hit.name for i in want]
        for g, w in zip(got, want):
            np.testing.assert_allclose(g.point, w.point, rtol=0, atol=1e-9)
            assert abs(g.distance - w.distance) <= 1e-9
            assert g.coordsys.name == w.coordsys.name == "world"
            # outward normal of the hit geometry at the hit point, in the hit node's frame
            lp_g = a.root.point_to_node(g.point, g.hit)
            lp_w = b.root.point_to_node(w.point, w.hit)
            np.testing.assert_allclose(g.hit.geometry.normal(lp_g), w.hit.geometry.normal(lp_w), rtol=0, atol=1e-9)
        hits += len(got)
    assert hits > 125  # the world sphere is hit by every ray: some rays hit the objects too


# ---- the YAML front end on the reference's OWN scene files (examples/, tests/data/, the mkdocs tutorials) -------------
def reference_yaml_parser():
    """pvtrace.cli.parse of the reference, importable once the stub top-level package carries the names it imports."""
    import sys

    pkg = ref_loader.load_reference_package()
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

    return refparse


REFERENCE_YAML = ("examples/hello_world.yml", "examples/sphere_with_luminescent_dye.yml", "examples/studio_lsc.yml",
                  "mkdocs/source/lsc_lumogen_red_cli/tutorial001.yml", "mkdocs/source/lsc_lumogen_red_cli/tutorial002.yml",
                  "mkdocs/source/lsc_lumogen_red_cli/tutorial003.yml", "tests/data/simple_box.yml", "tests/data/simple_box2.yml")


@pytest.mark.parametrize("path", REFERENCE_YAML)
def test_yaml_front_end_on_the_reference_scene_files(path):
    """Every scene file of the reference tree that the reference's own parser + compiler accept (pvtrace/cli/parse.py:72-551
    incl. `record: true` desugaring, cli/parse.py:469-525): parsed by pvtrace_b200.cli.parse and flattened -- the same tables."""
    import os

    import pvtrace_b200 as pv
    from pvtrace_b200.cli import parse as ourparse

    full = os.path.join(ref_loader.REFERENCE_ROOT, path)
    want = reference().compile_scene(reference_yaml_parser().parse(full))
    got = pv.engine.compile_scene(ourparse.parse(full))
    for table in scenes.TABLES:
        g, w = np.asarray(getattr(got, table)), np.asarray(getattr(want, table))
        assert g.shape == w.shape and g.dtype == w.dtype, (table, g.shape, w.shape, g.dtype, w.dtype)
        np.testing.assert_allclose(g, w, rtol=1e-13, atol=1e-13, err_msg=table)
    assert list(got.node_names) == list(want.node_names)
    assert list(got.component_names) == list(want.component_names)
    assert list(got.recorder_names) == list(want.recorder_names)


def test_yaml_scene_with_histogram_sampled_spectra_is_rejected_like_the_reference():
    import os

    import pvtrace_b200 as pv
    from pvtrace_b200.cli import parse as ourparse

    full = os.path.join(ref_loader.REFERENCE_ROOT, "tests/data/lsc_scene.yml")
    with pytest.raises(Exception) as theirs:
        reference().compile_scene(reference_yaml_parser().parse(full))
    with pytest.raises(pv.engine.UnsupportedSceneError) as mine:
        pv.engine.compile_scene(ourparse.parse(full))
    assert type(theirs.value).__name__ == "UnsupportedSceneError"
    assert str(mine.value) == str(theirs.value)


@pytest.mark.skipif(ref_loader.load_ref_kernel() is None, reason="oracle/_ref not built on this machine")
@pytest.mark.parametrize("seed", range(10))
def test_host_result_classes_and_tally_restatement_on_the_reference_kernels_output(seed):
    """EngineResult / RecorderResult / tally_histories of this package (the host side of the drop-in: pvtrace/engine/api.py:
    26-194, tally.py:86-156) fed with what the COMPILED REFERENCE KERNEL returns for a random scene: the recorders read
    from its tallies must equal the tallies recomputed from its event log by this package's restatement -- the reference's
    own contract (tests/test_engine.py:204-318), across libraries."""
    import pvtrace_b200 as pv
    from pvtrace_b200.engine import tally_histories
    from pvtrace_b200.engine.api import EngineResult

    kernel = ref_loader.load_ref_kernel()
    scene = random_scene(ours(), 700 + seed)
    compiled = pv.engine.compile_scene(scene)
    rng = np.random.default_rng(800 + seed)
    n, m = 1500, 256
    pos = rng.uniform(-6.0, 6.0, (n, 3))
    direction = rng.normal(size=(n, 3))
    direction /= np.linalg.norm(direction, axis=1)[:, None]
    wl = rng.uniform(350.0, 850.0, n)
    data = kernel.trace_bundle(compiled, pos, direction, wl, 5 + seed, 1000, m, seed % 3, 2, 1)
    assert data["counts"].max() < m - 1
    result = EngineResult(compiled, data, ["light"] * n, m, 1, 0.0)
    got, want = result.recorders, tally_histories(scene, result.histories())
    assert set(got) == set(want)
    for key in got:
        assert got[key].rays == want[key].rays and got[key].crossings == want[key].crossings, key
        for i in range(len(got[key].spec.histograms)):
            assert (got[key].histogram(i)[-1] == want[key].histogram(i)[-1]).all(), (key, i)
        for prop in ("wavelength", "angle", "duration", "pathlength"):
            if got[key].rays:
                assert got[key].mean(prop) == pytest.approx(want[key].mean(prop), rel=1e-9, abs=1e-18), (key, prop)
    assert sum(r.rays for r in got.values()) > n // 2  # (absorbed rays never reach the world's exit recorder)
