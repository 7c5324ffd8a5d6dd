"""The oracle's EXTENSIONS beyond the reference's compiled kernel, pinned to the reference itself
(tests/golden/make_reference_pins.py wrote the fixtures from the unmodified reference):

* seeded emission (oracle `emit_one` == device `emit_ray_value`) against pvtrace.engine.emit.emit_bundle,
  emit.py:22-134, for every built-in light delegate: distribution of every coordinate + second moments;
* the facet / coating table against the reference's SurfaceDelegate classes called directly: the LSC delegates
  of pvtrace/device/lsc.py:22-86 and two user delegates in the style of examples/006 Coatings.ipynb cell 3 (a
  partial mirror, a reflectivity spectrum) -- both the host restatement (FacetSurfaceDelegate) and the oracle's C;
* a disc carrying both coatings traced by the oracle against the reference's PYTHON tracer (per-ray event means).
"""
import functools
import os

import numpy as np
import pytest

import pvtrace_b200 as pv
from oracle import pvt_oracle
from pvtrace_b200.engine.compiler import EMIT_METHODS
from pvtrace_b200.light import light as light_module
from pvtrace_b200.light.event import Event
from pvtrace_b200.material import utils as material_utils
from pvtrace_b200.material.distribution import Distribution
from pvtrace_b200.material.surface import Facet, FacetSurfaceDelegate, Surface

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
LIGHT_LOCATION, LIGHT_ROTATION = (0.3, -0.2, 1.0), (0.6, (1.0, 1.0, 0.0))


def lamp(x):
    return np.exp(-(((x - 520.0) / 60.0) ** 2)) + 0.6 * np.exp(-(((x - 600.0) / 25.0) ** 2))


def emission_cases():
    x = np.linspace(400.0, 700.0, 121)
    return {
        "rect_cone": (light_module.RectangularMask(1.5, 0.7), material_utils.Cone(0.4),
                      light_module.ConstantWavelengthMask(600.0)),
        "circle_isotropic": (light_module.CircularMask(2.0), material_utils.isotropic,
                             light_module.SpectrumWavelengthMask(Distribution(x, lamp(x)))),
        "cube_lambertian": (light_module.CubeMask(1.0, 2.0, 0.5), material_utils.lambertian, None),
        "point_hg": (None, material_utils.HenyeyGreenstein(0.7), light_module.ConstantWavelengthMask(450.0)),
        "point_hg_zero": (None, material_utils.HenyeyGreenstein(0.0), None),
        "partial_forms": (functools.partial(light_module.rectangular_mask, 0.5, 2.5),
                          functools.partial(material_utils.cone, 0.9), None),
    }


@pytest.mark.parametrize("name", list(emission_cases()))
def test_seeded_emission_has_the_reference_delegates_distribution(name):
    golden = np.load(os.path.join(GOLDEN, "emission.npz"))
    position, direction, wavelength = emission_cases()[name]
    world = pv.Node(name="world", geometry=pv.Sphere(radius=50.0, material=pv.Material(refractive_index=1.0)))
    kwargs = {k: v for k, v in (("position", position), ("direction", direction), ("wavelength", wavelength)) if v is not None}
    node = pv.Node(name="light", parent=world, light=pv.Light(**kwargs))
    node.location = LIGHT_LOCATION
    node.rotate(LIGHT_ROTATION[0], np.asarray(LIGHT_ROTATION[1]) / np.linalg.norm(LIGHT_ROTATION[1]))
    emitter = pv.engine.compile_emitter(pv.Scene(world))
    assert emitter is not None, "built-in delegates (and their functools.partial forms) must lower to the device emitter"
    n = 200_000
    pos, dirs, wl = pvt_oracle.emit_bundle(emitter, n, seed=9)
    rows = np.column_stack((pos, dirs, wl))
    n_ref = int(golden[f"{name}_n"])
    probabilities, quantiles = golden["probabilities"], golden[f"{name}_quantiles"]
    np.testing.assert_allclose(np.linalg.norm(dirs, axis=1), 1.0, atol=1e-12)
    for col in range(7):
        q = quantiles[:, col]
        if q[-1] - q[0] < 1e-12:  # a constant coordinate (monochromatic light, ...)
            np.testing.assert_allclose(rows[:, col], q[0], atol=1e-9)
            continue
        # two-sample Kolmogorov-Smirnov at the reference's quantile points: |F_ours(q_k) - k/200| within 5 sigma
        ours = np.searchsorted(np.sort(rows[:, col]), q[1:-1], side="right") / n
        p = probabilities[1:-1]
        tol = 5.0 * np.sqrt(p * (1 - p) * (1.0 / n + 1.0 / n_ref)) + 2.0 / n_ref
        assert (np.abs(ours - p) <= tol).all(), (name, col, np.abs(ours - p).max())
    second = rows.T @ rows / n
    spread = np.sqrt(np.maximum((rows ** 2).T @ (rows ** 2) / n - second ** 2, 0.0) * (1.0 / n + 1.0 / n_ref))
    assert (np.abs(second - golden[f"{name}_second_moments"]) <= 6.0 * spread + 1e-9).all(), name


# ---- surface delegates -----------------------------------------------------------------------------------------

def _host_delegate(delegate, geometry, inner, outer, pos, dirs, wl, leaving):
    n = len(wl)
    refl, r_dir, t_dir = np.zeros(n), np.full((n, 3), np.nan), np.full((n, 3), np.nan)
    for i in range(n):
        ray = pv.Ray(position=tuple(pos[i]), direction=tuple(dirs[i]), wavelength=float(wl[i]))
        container, adjacent = (inner, outer) if leaving[i] else (outer, inner)
        refl[i] = delegate.reflectivity(None, ray, geometry, container, adjacent)
        r_dir[i] = delegate.reflected_direction(None, ray, geometry, container, adjacent)
        if refl[i] < 1.0:
            t_dir[i] = delegate.transmitted_direction(None, ray, geometry, container, adjacent)
    return refl, r_dir, t_dir


def _check_against_reference(golden, tag, scene, node_name):
    """Host FacetSurfaceDelegate and the oracle's facet table against the reference delegate's recorded answers."""
    pos, dirs, wl, leaving = (golden[f"{tag}_{k}"] for k in ("pos", "dir", "wl", "leaving"))
    want_R, want_r, want_t = (golden[f"{tag}_{k}"] for k in ("R", "reflected", "transmitted"))
    compiled = pv.engine.compile_scene(scene)
    inner = next(n for n in scene.root.children if n.name == node_name)
    outer = scene.root
    hit = compiled.node_names.index(node_name)
    world = compiled.node_names.index(outer.name)
    container = np.where(leaving > 0, hit, world).astype(np.int32)
    adjacent = np.where(leaving > 0, world, hit).astype(np.int32)
    # the reference's user delegate states R < 1 even where no refracted ray exists and then returns NaN directions
    # (its sqrt of a negative number): those rows reflect here (R = 1), everything else must agree
    no_ray = np.isnan(want_t).any(axis=1) & (want_R < 1.0)
    answers = {"host": _host_delegate(inner.geometry.material.surface.delegate, inner.geometry, inner, outer, pos, dirs,
                                      wl, leaving),
               "oracle": pvt_oracle.surface_event(compiled, hit, container, adjacent, pos, dirs, wl)}
    for who, (got_R, got_r, got_t) in answers.items():
        np.testing.assert_allclose(got_R[~no_ray], want_R[~no_ray], rtol=0, atol=1e-12, err_msg=f"{tag} {who} R")
        assert (got_R[no_ray] == 1.0).all(), (tag, who)
        np.testing.assert_allclose(got_r, want_r, rtol=0, atol=1e-12, err_msg=f"{tag} {who} reflected")
        open_rows = (want_R < 1.0) & ~no_ray
        np.testing.assert_allclose(got_t[open_rows], want_t[open_rows], rtol=0, atol=1e-12, err_msg=f"{tag} {who} transmitted")
    return no_ray.sum(), len(want_R)


@pytest.mark.parametrize("tag,cells,mirror", [("cells_and_mirror", {"left", "right", "near", "far"}, True),
                                              ("two_cells", {"left", "far"}, False), ("bare", set(), False)])
def test_lsc_facet_table_is_the_reference_delegate(tag, cells, mirror):
    """OptionalMirrorAndSolarCell (pvtrace/device/lsc.py:22-62) as data: every face, both sides."""
    golden = np.load(os.path.join(GOLDEN, "lsc_delegates.npz"))
    lsc = pv.LSC((5.0, 5.0, 1.0))
    if cells:
        lsc.add_solar_cell(cells)
    if mirror:
        lsc.add_back_surface_mirror()
    scene = lsc._make_scene(record=False)
    bad, total = _check_against_reference(golden, tag, scene, "LSC")
    assert bad == 0 and total == 240


def test_air_gap_mirror_reflects_everything():
    """AirGapMirror.reflectivity (lsc.py:65-71) == 1 on every face; same node size and place as the reference's."""
    golden = np.load(os.path.join(GOLDEN, "lsc_delegates.npz"))
    assert (golden["air_gap_R"] == 1.0).all()
    lsc = pv.LSC((5.0, 5.0, 1.0))
    lsc.add_air_gap_mirror(lambertian=False)
    scene = lsc._make_scene(record=False)
    mirror = next(n for n in scene.root.children if n.name == "Air Gap Mirror")
    np.testing.assert_allclose(mirror.geometry.size, golden["air_gap_size"])
    np.testing.assert_allclose(mirror.location, golden["air_gap_location"])
    compiled = pv.engine.compile_scene(scene)
    rows = slice(compiled.facet_start[compiled.node_names.index("Air Gap Mirror")], None)
    assert (compiled.facet_reflectivity[rows][:6] == 1.0).all()


def coated_facets(golden):
    """examples/006 Coatings.ipynb cell 3 as data (quarter mirror on the top face) + a reflectivity spectrum below."""
    return [Facet((0, 0, 1), reflectivity=1.0, region=((0.0, None), (0.0, None), None)),
            Facet((0, 0, -1), reflectivity=np.column_stack((golden["coating_x"], golden["coating_R"])))]


def test_coatings_are_the_reference_user_delegates():
    golden = np.load(os.path.join(GOLDEN, "coatings.npz"))
    world = pv.Node(name="world", geometry=pv.Box((15.0, 15.0, 15.0), material=pv.Material(refractive_index=1.0)))
    pv.Node(name="slab", parent=world, geometry=pv.Box((10.0, 10.0, 1.0), material=pv.Material(
        refractive_index=1.5, surface=Surface(delegate=FacetSurfaceDelegate(coated_facets(golden))))))
    bad, total = _check_against_reference(golden, "box", pv.Scene(world), "slab")
    assert 0 < bad < total // 4  # steep rays from inside onto the coated bottom: the rows where the rule above applies
    on_top = (golden["box_pos"][:, 2] == 0.5)
    mirrored = on_top & (golden["box_pos"][:, 0] > 0) & (golden["box_pos"][:, 1] > 0)
    assert mirrored.sum() > 5 and (golden["box_R"][mirrored] == 1.0).all()  # the fixture does exercise the quarter


def coated_disc_scene(golden):
    x = np.linspace(440.0, 660.0, 111)
    world = pv.Node(name="world", geometry=pv.Sphere(radius=10.0, material=pv.Material(refractive_index=1.0)))
    pv.Node(name="disc", parent=world, geometry=pv.Cylinder(length=1.0, radius=3.0, material=pv.Material(
        refractive_index=1.5, surface=Surface(delegate=FacetSurfaceDelegate(coated_facets(golden))),
        components=[pv.Absorber(coefficient=0.3, name="grey")])))
    light = pv.Node(name="lamp", parent=world, light=pv.Light(
        position=light_module.RectangularMask(2.0, 2.0), direction=material_utils.Cone(0.3),
        wavelength=light_module.SpectrumWavelengthMask(Distribution(x, lamp(x)))))
    light.location = (0.0, 0.0, 3.0)
    light.rotate(np.radians(180), (1, 0, 0))
    return pv.Scene(world)


DISC_EVENTS = [Event.REFLECT, Event.TRANSMIT, Event.ABSORB, Event.NONRADIATIVE, Event.EXIT, Event.KILL]


def assert_disc_statistics(golden, data, n, max_events):
    """Welch comparison of per-ray event means with the reference's Python tracer (tests/test_engine.py:117-128)."""
    kinds = data["kind"].reshape(n, max_events)
    valid = np.arange(max_events)[None, :] < data["counts"][:, None]
    assert data["counts"].max() < max_events - 1
    n_ref = int(golden["disc_n"])
    for event in DISC_EVENTS:
        ours = ((kinds == event.value) & valid).sum(axis=1).astype(float)
        se = np.sqrt(ours.var(ddof=1) / n + float(golden[f"disc_var_{event.name}"]) / n_ref)
        assert abs(ours.mean() - float(golden[f"disc_mean_{event.name}"])) <= 5.0 * se + 1e-9, event.name
    last = data["counts"] - 1
    rows = np.arange(n) * max_events + last
    down = ((data["kind"][rows] == Event.EXIT.value) & (data["direction"][rows, 2] < 0)).astype(float)
    p = float(golden["disc_exits_downwards"])
    assert abs(down.mean() - p) <= 5.0 * np.sqrt(p * (1 - p) * (1.0 / n + 1.0 / n_ref)), (down.mean(), p)


def test_oracle_traces_the_coated_disc_like_the_python_tracer():
    golden = np.load(os.path.join(GOLDEN, "coatings.npz"))
    scene = coated_disc_scene(golden)
    compiled, emitter = pv.engine.compile_scene(scene), pv.engine.compile_emitter(scene)
    n, m = 60_000, 64
    data = pvt_oracle.trace_bundle(compiled, None, None, None, 4, 1000, m, EMIT_METHODS["kT"], os.cpu_count() or 1, 1,
                                   emitter=emitter, n=n)
    assert_disc_statistics(golden, data, n, m)
