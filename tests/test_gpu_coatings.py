"""Coatings on the device (examples/006 Coatings.ipynb): partial facets and reflectivity spectra, against
the oracle ray by ray and against the reference's Python tracer statistically (fixtures of
tests/golden/make_reference_pins.py), and the rule that a coating cannot transmit where no refracted ray exists."""
import os

import numpy as np
import pytest

import pvtrace_b200 as pv
from oracle import pvt_oracle
from pvtrace_b200.engine import _cuda
from pvtrace_b200.engine.compiler import EMIT_METHODS
from pvtrace_b200.engine.recorder import Recorder
from pvtrace_b200.light.event import Event
from pvtrace_b200.material.surface import Facet, FacetSurfaceDelegate, Surface
from tests.test_reference_pins import GOLDEN, assert_disc_statistics, coated_disc_scene

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("flags", [0, _cuda.FLAG_REGISTER_KERNEL])
def test_coated_disc_matches_oracle_and_python_tracer(gpu, flags):
    golden = np.load(os.path.join(GOLDEN, "coatings.npz"))
    scene = coated_disc_scene(golden)
    compiled, emitter = pv.engine.compile_scene(scene), pv.engine.compile_emitter(scene)
    n, m = 60_000, 64
    got = _cuda.trace_bundle(compiled, None, None, None, 4, 1000, m, EMIT_METHODS["kT"], 0, 1, emitter=emitter, n=n, flags=flags)
    assert_disc_statistics(golden, got, n, m)  # the reference's photon_tracer.follow with the user delegates
    want = pvt_oracle.trace_bundle(compiled, None, None, None, 4, 1000, m, EMIT_METHODS["kT"], os.cpu_count() or 1, 1,
                                   emitter=emitter, n=n)
    same = (got["counts"] == want["counts"]) & (got["kind"].reshape(n, m) == want["kind"].reshape(n, m)).all(axis=1)
    assert same.mean() >= 0.999, same.mean()
    rows = np.repeat(same, m)
    np.testing.assert_allclose(got["position"][rows], want["position"][rows], rtol=0, atol=1e-7)
    np.testing.assert_allclose(got["direction"][rows], want["direction"][rows], rtol=0, atol=1e-7)


def test_a_coating_cannot_transmit_where_no_refracted_ray_exists(gpu):
    """Round-1 advisor finding: Facet(reflectivity < 1, transmit="refract") on a high-index node sent rays beyond the
    critical angle through sqrt(negative): NaN directions, photons that vanished from the accounting.  Now they
    reflect.  An anti-reflection coat (R = 0) on every face of a glass block with an isotropic source inside."""
    world = pv.Node(name="world", geometry=pv.Sphere(radius=10.0, material=pv.Material(refractive_index=1.0)))
    faces = [(1, 0, 0), (-1, 0, 0), (0, 1, 0), (0, -1, 0), (0, 0, 1), (0, 0, -1)]
    block = pv.Node(name="block", parent=world, geometry=pv.Box((2.0, 2.0, 2.0), material=pv.Material(
        refractive_index=1.5, components=[pv.Absorber(coefficient=0.5, name="grey")],
        surface=Surface(delegate=FacetSurfaceDelegate([Facet(f, reflectivity=0.0) for f in faces])))))
    pv.Node(name="lamp", parent=world, light=pv.Light(direction=pv.isotropic))
    world.recorders = [Recorder("exit", event="exit")]
    block.recorders = [Recorder("lost", event="lost"), Recorder("escaping", event="escaping")]
    scene = pv.Scene(world)
    n = 200_000
    result = pv.engine.simulate(scene, n, seed=2, record_every=100, max_events=300)
    rec = result.recorders
    assert rec["exit"].rays + rec["lost"].rays == n           # nobody vanished
    assert np.isfinite(result.data["direction"]).all()
    # from inside, the reflections are exactly the total internal ones (R = 0 otherwise): with an isotropic source in a
    # cube of n = 1.5 a sizeable share of the first hits (the `reflected` selector only counts reflections from outside:
    # count the logged events)
    events = result.event_counts()
    assert 0.2 * result.num_recorded < events[Event.REFLECT]
    assert rec["escaping"].rays > 0.3 * n
    compiled, emitter = pv.engine.compile_scene(scene), pv.engine.compile_emitter(scene)
    want = pvt_oracle.trace_bundle(compiled, None, None, None, 2, 1000, 300, EMIT_METHODS["kT"], os.cpu_count() or 1, 100,
                                   emitter=emitter, n=n)
    for key in ("rec_distinct", "rec_crossings"):
        assert np.abs(result.data[key] - want[key]).max() <= 5 * np.sqrt(0.002 * n) + 3, key
