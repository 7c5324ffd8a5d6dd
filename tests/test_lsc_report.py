"""LSC.counts_table / summary / spectrum (pvtrace_b200/device/lsc.py) against a plain loop over the histories that
follows the reference's definitions line by line (pvtrace/device/lsc.py:338-364 end rays, :455-506 counts,
:577-607 summary).  The histories come from the CPU oracle, so this runs without a GPU."""
import os

import numpy as np
import pytest

import pvtrace_b200 as pv
from oracle import pvt_oracle
from pvtrace_b200.device.lsc import FACES, LSC
from pvtrace_b200.engine.api import EngineResult, LightNames
from pvtrace_b200.engine.compiler import EMIT_METHODS
from pvtrace_b200.light.event import Event


def traced_lsc(n=1500, seed=7, cells=("left", "right", "near", "far"), mirror=True):
    lsc = LSC((5.0, 5.0, 1.0))
    if cells:
        lsc.add_solar_cell(set(cells))
    if mirror:
        lsc.add_back_surface_mirror()
    scene = lsc._make_scene()
    compiled, emitter = pv.engine.compile_scene(scene), pv.engine.compile_emitter(scene)
    m = 256
    data = pvt_oracle.trace_bundle(compiled, None, None, None, seed, 1000, m, EMIT_METHODS["kT"], os.cpu_count() or 1, 1,
                                   emitter=emitter, n=n)
    lsc._result = EngineResult(compiled, data, LightNames(emitter.light_names, n, 0), m, 1, 0.0)
    lsc._store = lsc._rows_of(lsc._result)  # what LSC.simulate keeps after a run
    return lsc


def facet_of(position, size):
    label = None
    for name, normal in FACES.items():
        axis = int(np.argmax(np.abs(normal)))
        if np.isclose(position[axis], normal[axis] * 0.5 * size[axis], atol=2.220446049250313e-13):
            label = name
    return label


def reference_style_counts(lsc):
    lights = lsc.light_names()
    table = {c: {f: 0 for f in FACES} for c in ("Solar In", "Solar Out", "Luminescent Out", "Luminescent In")}
    lost = 0
    for history in lsc._result.histories():
        rays, events, _ = zip(*history)
        stored = [("first", rays[1], events[1])]
        if events[-1] in (Event.ABSORB, Event.NONRADIATIVE, Event.REACT, Event.KILL):
            stored.append(("last", rays[-1], events[-1]))
            lost += events[-1] != Event.KILL
        elif events[-1] == Event.EXIT:
            stored.append(("last", rays[-2], events[-2]))
        for kind, ray, _ in stored:
            facet = facet_of(ray.position, lsc.size)
            if facet is None:
                continue
            solar = ray.source in lights
            column = ("Solar " if solar else "Luminescent ") + ("In" if kind == "first" else "Out")
            table[column][facet] += 1
    return table, lost


@pytest.mark.parametrize("cells,mirror", [(("left", "right", "near", "far"), True), ((), False)])
def test_counts_and_summary_follow_the_reference_definitions(cells, mirror):
    lsc = traced_lsc(cells=cells, mirror=mirror)
    want, lost = reference_style_counts(lsc)
    got = lsc.counts_table()
    assert got == want
    assert sum(got["Solar In"].values()) == lsc._result.num_recorded  # every ray enters through a face
    if mirror:
        assert got["Luminescent Out"]["bottom"] == 0 and got["Solar Out"]["bottom"] == 0
    summary = lsc.summary()
    incident = sum(want["Solar In"].values())
    collected = sum(want["Luminescent Out"][f] for f in cells)
    escaped = sum(want["Luminescent Out"][f] for f in FACES if f not in cells)
    assert summary["Incident"] == incident
    assert summary["Non-radiative Loss (fraction):"] == pytest.approx(lost / incident)
    if cells:
        assert summary["Optical Efficiency"] == pytest.approx(collected / incident)
        assert summary["Waveguide Efficiency"] == pytest.approx(collected / (collected + escaped))
    assert summary["Geometric Concentration"] == pytest.approx(25.0 / 20.0)
    assert summary["Components"] == {"Lumogen F Red 305", "Background"} and summary["Lights"] == {"Light"}


def test_spectrum_filters(capsys):
    lsc = traced_lsc(n=800)
    everything = lsc.spectrum(kind=None)
    first, last = lsc.spectrum(kind="first"), lsc.spectrum(kind="last")
    assert len(everything) == len(first) + len(last) and len(first) == 800
    assert (lsc.spectrum(kind="first", source="light") == 555.0).all()
    # the reference's form: names of lights / components, alone or as a collection (lsc.py:520-531)
    assert len(lsc.spectrum(kind="first", source="Light")) == 800
    by_name = lsc.spectrum(kind="last", source={"Lumogen F Red 305"})
    assert len(by_name) == len(lsc.spectrum(kind="last", source="luminescent")) > 50
    assert len(lsc.spectrum(kind="last", source=["Light", "Lumogen F Red 305"])) == len(last)
    with pytest.raises(ValueError):
        lsc.spectrum(source="Unobtainium")
    assert lsc.counts() == lsc.counts_table()
    red = lsc.spectrum(kind="last", source="luminescent", facets={"left", "right", "near", "far"})
    assert len(red) > 50 and red.min() > 555.0  # kT emission: collected light is red-shifted
    by_event = {e: len(lsc.spectrum(kind="last", events={e})) for e in ("nonradiative", "transmit", "reflect", "kill")}
    assert sum(by_event.values()) == len(last) and by_event["nonradiative"] > 50 and by_event["transmit"] > 50
    with pytest.raises(ValueError):
        lsc.spectrum(kind="sideways")
    lsc.report()
    assert "Surface Counts:" in capsys.readouterr().out


def test_report_needs_logged_histories():
    lsc = traced_lsc(n=50)
    lsc._store = None  # a run with record_every=0 keeps nothing
    with pytest.raises(ValueError):
        lsc.counts_table()
    with pytest.raises(ValueError):
        LSC((5.0, 5.0, 1.0)).summary()
