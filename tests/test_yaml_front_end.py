"""The YAML scene front end (pvtrace_b200/cli/parse.py) against the reference's own parser: the scenes
both build from the same files (tests/data/*.yml) flatten to identical tables (fixtures tests/golden/yaml_*.npz were
produced by running pvtrace/cli/parse.py + compile_scene of the unmodified reference), plus the reference's recorder
tests (tests/test_engine.py:353-415)."""
import os

import numpy as np
import pytest

import pvtrace_b200 as pv
from pvtrace_b200.cli.parse import SpecError, parse, parse_spec
from pvtrace_b200.engine import UnsupportedSceneError
from tests import scenes

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")


@pytest.mark.parametrize("stem", ["lsc_recorded", "primitives"])
def test_parsed_scene_flattens_like_the_reference(stem):
    want = np.load(os.path.join(GOLDEN, f"yaml_{stem}.npz"))
    scene = parse(os.path.join(HERE, "data", stem + ".yml"))
    got = pv.engine.compile_scene(scene)
    for table in scenes.TABLES:
        g, w = np.asarray(getattr(got, table)), want[table]
        assert g.shape == w.shape and g.dtype == w.dtype, table
        if np.issubdtype(w.dtype, np.integer):
            assert (g == w).all(), table
        else:
            np.testing.assert_allclose(g, w, rtol=1e-13, atol=1e-13, err_msg=table)
    for key in ("node_names", "component_names", "recorder_names"):
        assert list(getattr(got, key)) == list(want[key]), key
    lights = scene.light_nodes
    assert [n.name for n in lights] == list(want["light_names"])
    poses = np.array([np.asarray(n.transformation_to(scene.root)) for n in lights])
    np.testing.assert_allclose(poses, want["light_to_world"], atol=1e-13)
    assert pv.engine.compile_emitter(scene) is not None  # every mask the spec can express samples on the device


def test_yaml_recorders_parse():
    """tests/test_engine.py:353-371"""
    scene = parse(os.path.join(HERE, "data", "lsc_recorded.yml"))
    recorders = {r.name: r for node in scene.root.iter_preorder() for r in node.recorders}
    assert recorders["edge-escape"].facet == (1.0, 0.0, 0.0)
    assert {"lsc-top", "lsc-bottom", "lsc-east", "lsc-west", "lsc-north", "lsc-south", "lsc-lost"} <= set(recorders)
    assert len(recorders["lsc-top"].histograms) == 3


def test_record_shorthand_and_errors(tmp_path):
    """tests/test_engine.py:374-415 (the device run of the parsed scene is in tests/test_gpu_engine.py)"""
    spec = {"version": "1.0", "nodes": {
        "world": {"sphere": {"radius": 10.0, "material": {"refractive-index": 1.0}}},
        "slab": {"record": True, "box": {"size": [5, 5, 1], "material": {"refractive-index": 1.5}}},
        "laser": {"location": [0, 0, 3], "direction": [0, 0, -1], "light": {"wavelength": 555}}}}
    scene = parse_spec(spec)
    recorders = {r.name: r for node in scene.root.iter_preorder() for r in node.recorders}
    assert len(recorders) == 7 and recorders["slab-top"].facet == (0.0, 0.0, 1.0) and "slab-lost" in recorders
    with pytest.raises(ValueError):
        parse_spec(dict(spec, version="2.0"))
    bad = {"version": "1.0", "nodes": dict(spec["nodes"], cube={"mesh": {"file": "x.stl", "material": {"refractive-index": 1.5}}})}
    with pytest.raises(UnsupportedSceneError):
        parse_spec(bad)
    with pytest.raises(SpecError):
        parse_spec({"version": "1.0", "nodes": dict(spec["nodes"], thing={"parent": "world"})})
    with pytest.raises(SpecError):
        parse_spec(dict(spec, recorders={"r": {"node": "nowhere", "event": "exit"}}))
    # CSV spectra: index column, x, y (the reference reads them with pandas, usecols 0-2, index_col 0)
    csv = tmp_path / "dye.csv"
    x = np.arange(400, 701, 10)
    csv.write_text("i,nm,abs\n" + "\n".join(f"{i},{a},{np.exp(-((a - 550) / 40.0) ** 2)}" for i, a in enumerate(x)))
    spec2 = dict(spec, components={"dye": {"absorber": {"coefficient": 2.0, "spectrum": {"file": str(csv)}}}})
    spec2["nodes"] = dict(spec["nodes"])
    spec2["nodes"]["slab"] = {"box": {"size": [5, 5, 1], "material": {"refractive-index": 1.5, "components": ["dye"]}}}
    compiled = pv.engine.compile_scene(parse_spec(spec2, str(tmp_path)))
    assert compiled.abs_y.max() == pytest.approx(2.0) and len(compiled.abs_x) == len(x)
