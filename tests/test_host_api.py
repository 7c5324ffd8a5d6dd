"""Host-side mirror of the reference interface: the known answers of the reference's own unit tests
(tests/test_box.py, test_sphere.py, test_cylinder.py, test_node.py, test_distibution.py, test_frensel_*.py,
test_refractored_tracer.py:180-184) asserted against pvtrace_b200's classes."""
import functools
import math

import numpy as np
import pytest

import pvtrace_b200 as pv
from pvtrace_b200.geometry.transformations import rotation_matrix
from pvtrace_b200.geometry.utils import norm
from pvtrace_b200.material.utils import bandgap, fresnel_reflectivity, fresnel_refraction, specular_reflection


def test_box_known_answers():
    b = pv.Box(size=(1, 1, 1))
    for p in ((0.5, 0, 0), (0, 0.5, 0), (0, 0, 0.5), (-0.5, 0, 0), (0, -0.5, 0), (0, 0, -0.5)):
        assert b.is_on_surface(p)
        assert np.allclose(b.normal(p), np.sign(p))
    assert not b.is_on_surface((0, 0, 0)) and not b.is_on_surface((0.501, 0, 0))
    assert b.contains((0, 0, 0)) and not b.contains((0, 0, 0.5)) and not b.contains((0, 0, 1.0))
    assert np.allclose(b.intersections((-2.0, 0, 0), (1.0, 0, 0)), ((-0.5, 0, 0), (0.5, 0, 0)))
    assert b.is_entering((0.5, 0, 0), (-1.0, 0, 0)) and not b.is_entering((0.5, 0, 0), (1.0, 0, 0))
    # the two historical "bad" points of tests/test_box.py:25-36
    assert pv.Box(size=(1.0, 1.0, 0.02)).is_on_surface((0.06608370507653762, 0.5, -0.007798573829629238))
    assert pv.Box(size=(1.0, 3.0, 0.02)).is_on_surface((-0.5, -0.2415708917159319, -0.008736363958583498))


def test_sphere_known_answers():
    s = pv.Sphere(radius=1)
    assert s.is_on_surface((0, 0, 1.0)) and not s.is_on_surface((0, 0, 0))
    assert s.contains((0, 0, 0)) and not s.contains((0, 0, 1.0)) and not s.contains((0, 0, 2.0))
    assert np.allclose(s.intersections((-2.0, 0, 0), (1.0, 0, 0)), ((-1.0, 0, 0), (1.0, 0, 0)))
    assert np.allclose(s.normal((0, 0, 1.0)), (0, 0, 1.0))
    assert s.is_entering((-1.0, 0, 0), (1.0, 0, 0)) and not s.is_entering((-1.0, 0, 0), (-1.0, 0, 0))


def test_cylinder_known_answers():
    c = pv.Cylinder(length=1.0, radius=1.0)
    assert c.contains((0, 0, 0)) and c.contains((0.25, 0.25, 0.25))
    for p in ((0, 0, 0.5), (0, 0, -0.5), (0, 1.0, 0), (-1.0, 0, 0), (0, 0, 0.6), (0, 1.1, 0)):
        assert not c.contains(p)
    pts = c.intersections((-2, 0.2, 0.0), norm((1.0, 0.2, -0.2)))
    assert np.allclose(pts, ((-0.9082895433880116, 0.41834209132239775, -0.2183420913223977), (0.5, 0.7, -0.5)))
    assert np.allclose(c.normal((0, 0, 0.5)), (0, 0, 1)) and np.allclose(c.normal((0, 0, -0.5)), (0, 0, -1))
    assert np.allclose(c.normal((0, 1.0, 0)), (0, 1, 0)) and np.allclose(c.normal((0, -1.0, 0)), (0, -1, 0))
    assert c.is_entering((0, 0, 0.5), norm((1, 1, -1))) and not c.is_entering((0, 0, 0.5), norm((1, 1, 1)))
    assert c.is_entering((-1.0, 0, 0), norm((1, 1, 1))) and not c.is_entering((-1.0, 0, 0), norm((-1, 1, 1)))


def test_fresnel_known_answers():
    assert np.isclose(fresnel_reflectivity(0.0, 1.0, 1.5), 0.04)
    for normal in ((0, 0, 1.0), (0, 0, -1.0)):
        assert np.allclose(specular_reflection((0, 0, -1.0), normal), (0, 0, 1.0))
    assert np.allclose(fresnel_refraction((0, 0, -1.0), (0, 0, -1.0), 1.0, 1.5), (0, 0, -1.0))


def test_node_frames():
    """tests/test_node.py:38-63"""
    a = pv.Node(name="a")
    b = pv.Node(name="b", parent=a)
    c = pv.Node(name="c", parent=b)
    d = pv.Node(name="d", parent=a)
    b.translate((1, 1, 1)); c.translate((0, 1, 1)); d.translate((-1, -1, -1))
    theta = 0.5 * np.pi
    b.rotate(theta, (0, 0, 1)); c.rotate(theta, (1, 0, 0)); d.rotate(theta, (0, 1, 0))
    assert np.allclose(d.point_to_node((0, 0, 0), a), (-1, -1, -1))
    assert np.allclose(d.point_to_node((1, 1, 1), a), (0, 0, -2))
    assert np.allclose(d.vector_to_node((1, 0, 0), a), (0, 0, -1))
    assert np.allclose(d.vector_to_node((0, 0, 1), a), (1, 0, 0))
    assert np.allclose(c.point_to_node((0, 0, 0), d), (-3, 2, 1))
    assert np.allclose(c.point_to_node((1, 1, 1), d), (-4, 3, 2))
    assert np.allclose(c.vector_to_node((1, 0, 0), d), (0, 1, 0))
    assert np.allclose(c.vector_to_node((0, 1, 0), d), (-1, 0, 0))
    assert b.parent is a and a.parent is None and set(a.children) == {b, d}


def test_node_look_at_and_intersections():
    a = pv.Node(name="A")
    a.look_at([1, 0, 0])
    assert np.allclose(rotation_matrix(np.pi / 2, [0, 1, 0]), a.pose)
    a = pv.Node(name="A")
    a.look_at([0, 0, -1])
    assert np.allclose(rotation_matrix(np.pi, [0, 1, 0]), a.pose)
    a = pv.Node(name="A")
    b = pv.Node(name="B", parent=a, geometry=pv.Sphere(radius=1.0))
    b.translate((1.0, 0, 0))
    pts = np.array([x.to(a).point for x in a.intersections((-2.0, 0, 0), (1.0, 0, 0))])
    assert np.allclose(pts, ((0, 0, 0), (2.0, 0, 0)))
    scene = pv.Scene(a)
    assert [i.hit.name for i in scene.intersections((-2.0, 0, 0), (1.0, 0, 0))] == ["B", "B"]


def test_distribution_known_answers():
    """tests/test_distibution.py:9-41"""
    x = np.linspace(400, 1010, 2000)
    dist = pv.Distribution(x, np.exp(-((x - 700.0) / 50.0) ** 2))
    assert np.isclose(dist.sample(0), x.min()) and np.isclose(dist.sample(1), x.max())
    assert np.isclose(dist.lookup(x.min()), 0.0) and np.isclose(dist.lookup(x.max()), 1.0)
    xs = np.arange(400.0, 801.0, 1.0)
    step = pv.Distribution(xs, bandgap(xs, 600.0, 1.0), hist=True)
    assert np.isclose(step.sample(0), 400.0) and np.isclose(step.sample(1), 600.0)
    assert step.lookup(800.0) == 1.0
    values = step.sample(np.linspace(step.lookup(598.0), step.lookup(601.0), 10000))
    assert len(set(values.tolist())) == 3
    with pytest.raises(ValueError):
        dist(100.0)


def test_beer_lambert_known_answer():
    """tests/test_refractored_tracer.py:180-184: alpha = 10 cm^-1, np.random.seed(0); the first MT19937 draw is
    spent on the Fresnel test, the second gives the free path: z = -0.5 + depth = -0.3744069237034118."""
    np.random.seed(0)
    np.random.uniform()
    material = pv.Material(refractive_index=1.0, components=[pv.Absorber(coefficient=10.0)])
    assert -0.5 + material.penetration_depth(555.0) == pytest.approx(-0.3744069237034118, abs=1e-13)
    assert pv.Material(1.0).penetration_depth(555.0) == math.inf


def test_light_and_scene_emit():
    world = pv.Node(name="world", geometry=pv.Sphere(10.0, material=pv.Material(1.0)))
    a = pv.Node(name="a", parent=world, light=pv.Light(name="a"))
    b = pv.Node(name="b", parent=world, light=pv.Light(name="b", direction=functools.partial(pv.cone, 0.1)))
    b.location = (0, 0, 1.0)
    b.rotate(np.pi, (1, 0, 0))
    rays = list(pv.Scene(world).emit(4))
    assert [r.source for r in rays] == ["a", "b", "a", "b"]  # lights take turns, scene/scene.py:141-151
    assert rays[0].position == (0.0, 0.0, 0.0) and rays[0].direction == (0.0, 0.0, 1.0) and rays[0].wavelength == 555.0
    assert np.allclose(rays[1].position, (0, 0, 1.0)) and rays[1].direction[2] < -0.99
    with pytest.raises(ValueError):
        pv.cone(0.0)


def test_ray_propagate():
    ray = pv.Ray(position=(0, 0, 0), direction=(0, 0, 1.0), wavelength=555.0)
    moved = ray.propagate(3.0, 1.5)
    assert moved.position == (0, 0, 3.0) and moved.travelled == 3.0
    assert moved.duration == pytest.approx(3.0 * 1.5 / 2.99792458e10)


def test_emitter_lowering_recognises_partials():
    """functools.partial light delegates (hello_world, nested_cylinders, LSC default) must lower to device
    descriptors; the reference's vectorised emitter falls back to one Python call per ray for them (emit.py:67,116)."""
    from pvtrace_b200.device import configs
    from pvtrace_b200.engine import compiler

    for name, (build, _) in configs.CONFIGS.items():
        emitter = compiler.compile_emitter(build())
        assert emitter is not None, name
    e = compiler.compile_emitter(configs.validation())
    assert e.pos_kind[0] == compiler.LPOS_RECT and e.wl_kind[0] == compiler.LWL_SPECTRUM and e.wl_n[0] == 401
    world = pv.Node(name="w", geometry=pv.Sphere(5.0, material=pv.Material(1.0)))
    pv.Node(name="l", parent=world, light=pv.Light(direction=lambda: (0.0, 0.0, 1.0)))
    assert compiler.compile_emitter(pv.Scene(world)) is None  # unknown delegate => host emission


def test_lsc_delegate_lowering():
    """OptionalMirrorAndSolarCell (pvtrace/device/lsc.py:22-62) as data: facet table of the LSC node."""
    lsc = pv.LSC((5.0, 5.0, 1.0))
    lsc.add_solar_cell({"left", "right"})
    lsc.add_back_surface_mirror()
    compiled = pv.engine.compile_scene(lsc._make_scene())
    assert compiled.n_facets == 3
    node = compiled.node_names.index("LSC")
    rows = slice(compiled.facet_start[node], compiled.facet_start[node] + compiled.facet_count[node])
    table = {tuple(n): (r, f) for n, r, f in zip(compiled.facet_normal[rows].tolist(),
                                                 compiled.facet_reflectivity[rows], compiled.facet_flags[rows])}
    assert table[(0.0, 0.0, -1.0)] == (1.0, 0)          # mirror
    assert table[(-1.0, 0.0, 0.0)] == (0.0, 1) and table[(1.0, 0.0, 0.0)] == (0.0, 1)  # R = 0, straight through
    with pytest.raises(ValueError):
        lsc.add_solar_cell({"top"})
    # host restatement of the delegate for one ray
    delegate = lsc._scene.root.children[0].geometry.material.surface.delegate
    geometry = lsc._scene.root.children[0].geometry
    world, node = lsc._scene.root, lsc._scene.root.children[0]
    ray = pv.Ray(position=(0.0, 0.0, -0.5), direction=norm((0.3, 0.0, -1.0)), wavelength=600.0)
    assert delegate.reflectivity(None, ray, geometry, node, world) == 1.0
    ray = pv.Ray(position=(2.5, 0.0, 0.0), direction=norm((1.0, 0.2, 0.0)), wavelength=600.0)
    assert delegate.reflectivity(None, ray, geometry, node, world) == 0.0
    assert delegate.transmitted_direction(None, ray, geometry, node, world) == tuple(ray.direction)
    ray = pv.Ray(position=(0.0, 0.0, 0.5), direction=(0.0, 0.0, 1.0), wavelength=600.0)
    assert delegate.reflectivity(None, ray, geometry, node, world) == pytest.approx(0.04)


def test_auto_recorders_match_yaml_desugaring():
    """`record: true` on a 5x5x1 box => 1 lost + 6 per-face escaping recorders, 7808 bins."""
    from pvtrace_b200.device import configs

    compiled = pv.engine.compile_scene(configs.lsc_default())
    assert len(compiled.rec_node) == 8 and compiled.total_bins == 7808
