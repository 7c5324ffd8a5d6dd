"""Edges of the device path that the five benchmark scenes never reach: more recorders than the pool kernel's 64-bit seen
mask holds, the node limit of the reference (128, _kernel.pyx:65-68), a scene whose tables do not fit into shared
memory, and ray arrays the intersect stage's bulk copies cannot start from.  Each against the oracle."""
import os

import numpy as np
import pytest

import pvtrace_b200 as pv
from oracle import pvt_oracle
from pvtrace_b200.engine import _cuda
from pvtrace_b200.engine.compiler import EMIT_METHODS
from pvtrace_b200.engine.recorder import Histogram, Recorder
from pvtrace_b200.material.utils import gaussian
from tests import scenes

pytestmark = pytest.mark.gpu
INTEGER_KEYS = ("rec_distinct", "rec_crossings", "rec_bins")


def _against_oracle(scene, n, seed=3, max_events=64, record_every=50):
    compiled, emitter = pv.engine.compile_scene(scene), pv.engine.compile_emitter(scene)
    got = _cuda.trace_bundle(compiled, None, None, None, seed, 1000, max_events, EMIT_METHODS["kT"], 0, record_every,
                             emitter=emitter, n=n)
    want = pvt_oracle.trace_bundle(compiled, None, None, None, seed, 1000, max_events, EMIT_METHODS["kT"],
                                   os.cpu_count() or 1, record_every, emitter=emitter, n=n)
    assert got["stats"][_cuda.STAT_RAYS] == n
    m = max_events
    same = (got["counts"] == want["counts"]) & (got["kind"].reshape(-1, m) == want["kind"].reshape(-1, m)).all(axis=1)
    assert same.mean() >= 0.995, same.mean()
    for key in INTEGER_KEYS:
        if got[key].size:
            p = np.clip(want[key] / n, 1e-9, 1 - 1e-9)
            assert (np.abs(got[key] - want[key]) <= 5 * np.sqrt(2 * 0.001 * n * p * (1 - p)) + 3).all(), key
    return compiled, got, want


def test_more_than_64_recorders_take_the_wide_mask_kernel(gpu):
    """The pool kernel keeps a ray's distinct-recorder mask in 64 bits; scenes with more recorders (up to the reference's
    256) go through trace_kernel<., 8>.  80 recorders on the slab, every one with a histogram."""
    scene = scenes.lsc()
    slab = next(n for n in scene.root.children if n.name == "slab")
    faces = [(1, 0, 0), (-1, 0, 0), (0, 1, 0), (0, -1, 0), (0, 0, 1), (0, 0, -1)]
    for k in range(76):
        slab.recorders.append(Recorder(f"extra-{k}", event=("escaping", "reflected", "entering")[k % 3], facet=faces[k % 6],
                                       histograms=[Histogram("wavelength", 300.0, 1000.0, 7)]))
    compiled, got, want = _against_oracle(scene, 60_000)
    assert len(compiled.rec_node) == 81
    dup = [i for i, name in enumerate(compiled.recorder_names) if name.startswith("extra-")]
    assert got["rec_distinct"][dup].sum() > 10_000  # the extra recorders do count
    # recorders with the same (event, facet) see the same rays
    assert got["rec_distinct"][dup[0]] == got["rec_distinct"][dup[18]]


def test_the_reference_node_limit(gpu):
    """128 geometry nodes trace (and agree with the oracle); 129 raise the reference's error (compiler.py:23, :80-81)."""
    def build(count):
        world = pv.Node(name="world", geometry=pv.Sphere(radius=50.0, material=pv.Material(refractive_index=1.0)))
        rng = np.random.default_rng(5)
        for k in range(count - 1):
            ball = pv.Node(name=f"ball-{k}", parent=world, geometry=pv.Sphere(radius=0.4, material=pv.Material(
                refractive_index=1.3 + 0.002 * k)))
            ball.location = tuple((np.array([k % 8, (k // 8) % 4, k // 32]) * 1.1 - np.array([3.85, 1.65, 1.65]) +
                                   0.02 * rng.normal(size=3)).tolist())
        light = pv.Node(name="lamp", parent=world, light=pv.Light(direction=pv.isotropic))
        light.location = (0.55, 0.55, 0.55)
        world.recorders = [Recorder("exit", event="exit")]
        return pv.Scene(world)

    compiled, got, want = _against_oracle(build(128), 40_000, max_events=96)
    assert len(compiled.geom_type) == 128 and got["rec_distinct"][0] >= 40_000 - 40
    with pytest.raises(ValueError, match="at most 128"):
        pv.engine.simulate(build(129), 10, seed=1)


def test_tables_too_large_for_shared_memory(gpu):
    """Spectra of 9000 knots each: a blob of ~300 KB stays in global memory (read through L1 / L2) and the one-photon-
    per-lane kernel traces the scene -- same histories as the oracle."""
    x = np.linspace(300.0, 1000.0, 9000)
    world = pv.Node(name="world", geometry=pv.Sphere(radius=10.0, material=pv.Material(refractive_index=1.0)))
    slab = pv.Node(name="slab", parent=world, geometry=pv.Box((5.0, 5.0, 1.0), material=pv.Material(
        refractive_index=1.5, components=[
            pv.Luminophore(coefficient=np.column_stack((x, 5.0 * gaussian(x, 1.0, 480.0, 40.0))),
                           emission=np.column_stack((x, gaussian(x, 1.0, 600.0, 40.0))), quantum_yield=0.9, name="dye"),
            pv.Absorber(coefficient=0.3, name="background")])))
    light = pv.Node(name="light", light=pv.Light(), parent=world)
    light.location = (0.0, 0.0, -3.0)
    world.recorders = [Recorder("exit", event="exit")]
    slab.recorders = [Recorder("lost", event="lost")]
    compiled, got, want = _against_oracle(pv.Scene(world), 50_000)
    assert len(compiled.abs_x) >= 9000 and got["rec_distinct"].sum() >= 50_000 - 5


@pytest.mark.parametrize("n", [1, 255, 256, 257, 100_001])
def test_intersect_stage_tail_and_unaligned_arrays(gpu, n):
    """The ring moves whole tiles of 256 rays with 16-byte aligned bulk copies: the last n % 256 rays and arrays that start
    8 bytes off a 16-byte boundary take plain loads -- same answers."""
    import torch

    from pvtrace_b200.device import configs

    scene = configs.lsc_default()
    compiled = pv.engine.compile_scene(scene)
    rng = np.random.default_rng(n)
    pos = rng.uniform(-3, 3, size=(n, 3))
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    w_t0, w_hit, w_cont, w_adj = pvt_oracle.intersect_bundle(compiled, pos, d)
    with _cuda.Context(compiled, None, 0) as ctx:
        for shift in (0, 1):  # 1: the arrays start one double into a 16-byte aligned allocation
            buf_p = torch.zeros(3 * n + 2, dtype=torch.float64, device="cuda")
            buf_d = torch.zeros(3 * n + 2, dtype=torch.float64, device="cuda")
            dp, dd = buf_p[shift:shift + 3 * n], buf_d[shift:shift + 3 * n]
            dp.copy_(torch.from_numpy(pos.ravel())); dd.copy_(torch.from_numpy(d.ravel()))
            assert (dp.data_ptr() % 16 == 0) == (shift == 0)
            t0 = torch.empty(n, dtype=torch.float64, device="cuda")
            ids = torch.empty(n, dtype=torch.int32, device="cuda")
            ctx.intersect_packed(dp.data_ptr(), dd.data_ptr(), n, t0.data_ptr(), ids.data_ptr())
            torch.cuda.synchronize()
            packed = ids.cpu().numpy().view(np.uint32).astype(np.int64)
            unpack = lambda sh: np.where(((packed >> sh) & 0xff) == 0xff, -1, (packed >> sh) & 0xff)  # noqa: E731
            ok = (unpack(0) == w_hit) & (unpack(8) == w_cont) & (unpack(16) == w_adj)
            assert ok.mean() >= 0.9999, (n, shift, ok.mean())
            np.testing.assert_allclose(t0.cpu().numpy()[ok], w_t0[ok], rtol=1e-10)
