"""The flattener must produce the reference's table contract: every CompiledScene array built here from
pvtrace_b200 classes equals the one pvtrace.engine.compiler.compile_scene (pvtrace/engine/compiler.py:57-204)
built from the reference's classes for the same scene (fixtures: tests/golden/tables_*.npz)."""
import os

import numpy as np
import pytest

import pvtrace_b200 as pv
from tests import scenes

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("name", list(scenes.SCENES))
def test_tables_equal_reference(name):
    want = np.load(os.path.join(GOLDEN, f"tables_{name}.npz"))
    got = pv.engine.compile_scene(scenes.SCENES[name]())
    for table in scenes.TABLES:
        g, w = np.asarray(getattr(got, table)), want[table]
        assert g.shape == w.shape, (table, g.shape, w.shape)
        assert g.dtype == w.dtype, (table, g.dtype, w.dtype)
        if np.issubdtype(w.dtype, np.integer):
            assert (g == w).all(), table
        else:
            np.testing.assert_allclose(g, w, rtol=1e-13, atol=1e-13, err_msg=table)
    assert got.root_id == int(want["root_id"])
    assert got.total_bins == int(want["total_bins"])
    assert list(got.node_names) == list(want["node_names"])
    assert list(got.component_names) == list(want["component_names"])
    assert list(got.recorder_names) == list(want["recorder_names"])


def test_reference_compatible_scenes_have_no_facets():
    for name in scenes.SCENES:
        assert pv.engine.compile_scene(scenes.SCENES[name]()).n_facets == 0


def test_unsupported_scenes_raise():
    from pvtrace_b200.engine import UnsupportedSceneError
    from pvtrace_b200.material.surface import FresnelSurfaceDelegate

    class Custom(FresnelSurfaceDelegate):
        pass

    world = pv.Node(name="w", geometry=pv.Sphere(5.0, material=pv.Material(1.0)))
    pv.Node(name="b", parent=world, geometry=pv.Box((1, 1, 1), material=pv.Material(
        1.5, surface=pv.Surface(delegate=Custom()))))
    pv.Node(name="l", parent=world, light=pv.Light())
    with pytest.raises(UnsupportedSceneError):
        pv.engine.compile_scene(pv.Scene(world))
    # custom phase function (compiler.py:300-310)
    world = pv.Node(name="w", geometry=pv.Sphere(5.0, material=pv.Material(1.0, components=[
        pv.Scatterer(1.0, phase_function=lambda: (0.0, 0.0, 1.0))])))
    with pytest.raises(UnsupportedSceneError):
        pv.engine.compile_scene(pv.Scene(world))
    # facet filter on a volume event (compiler.py:140-155)
    world = pv.Node(name="w", geometry=pv.Sphere(5.0, material=pv.Material(1.0)))
    world.recorders.append(pv.engine.Recorder("bad", event="lost", facet=(0, 0, 1)))
    with pytest.raises(UnsupportedSceneError):
        pv.engine.compile_scene(pv.Scene(world))
    with pytest.raises(ValueError):
        pv.engine.simulate(scenes.fresnel(), 10, emit_method="nope")
