"""The drop-in host call on hardware: constant columns of the caller's arrays never cross PCIe, a bundle spread over
several GPUs of one process or over several ranks gives the tallies of one GPU (integer fields identical,
sums rtol 1e-12; reference contract tests/test_engine.py:169-176, pvtrace/engine/api.py:252-262), and seeds are
independent runs."""
import os
import socket

import numpy as np
import pytest

import pvtrace_b200 as pv
from oracle import pvt_oracle
from pvtrace_b200.device import configs
from pvtrace_b200.engine import _cuda, distributed
from pvtrace_b200.engine.compiler import EMIT_METHODS
from tests import scenes

pytestmark = pytest.mark.gpu
INTEGER_KEYS = ("rec_distinct", "rec_crossings", "rec_bins")


def _lsc_bundle(n, seed=5):
    scene = configs.lsc_default()
    compiled, emitter = pv.engine.compile_scene(scene), pv.engine.compile_emitter(scene)
    pos, direction, wl = pvt_oracle.emit_bundle(emitter, n, seed=seed)
    return compiled, pos, direction, wl


def _trace(compiled, pos, direction, wl, elide, **kw):
    old = os.environ.get("PVT_ELIDE_CONSTANT")
    os.environ["PVT_ELIDE_CONSTANT"] = "1" if elide else "0"
    try:
        return _cuda.trace_bundle(compiled, pos, direction, wl, 5, 1000, 64, EMIT_METHODS["kT"], 0, 0, **kw)
    finally:
        if old is None:
            del os.environ["PVT_ELIDE_CONSTANT"]
        else:
            os.environ["PVT_ELIDE_CONSTANT"] = old


def _same(a, b):
    for key in INTEGER_KEYS:
        assert (a[key] == b[key]).all(), key
    np.testing.assert_allclose(a["rec_sums"], b["rec_sums"], rtol=1e-12, atol=0)


def test_constant_columns_are_not_uploaded(gpu):
    """Config 2's light is a point source of one wavelength: 24 of 56 bytes per ray are upload, the result is unchanged."""
    n = 600_000
    compiled, pos, direction, wl = _lsc_bundle(n)
    assert (pos == pos[0]).all() and (wl == wl[0]).all() and not (direction == direction[0]).all()
    plain = _trace(compiled, pos, direction, wl, elide=False)
    lean = _trace(compiled, pos, direction, wl, elide=True)
    _same(plain, lean)
    assert plain["stats"][_cuda.STAT_H2D_BYTES] == 56 * n
    assert lean["stats"][_cuda.STAT_H2D_BYTES] == 24 * n
    assert lean["stats"][_cuda.STAT_RAYS] == n and lean["rec_distinct"].sum() >= n


def test_a_column_that_differs_late_is_caught(gpu):
    """The check runs beside the trace: a column that only LOOKS constant costs a second trace, never a wrong result."""
    n = 600_000
    compiled, pos, direction, wl = _lsc_bundle(n)
    pos, wl = pos.copy(), wl.copy()
    pos[n - 7] += (0.5, -0.25, 0.0)  # one ray from somewhere else, beyond the rows probed up front
    wl[n // 2] = 470.0
    plain = _trace(compiled, pos, direction, wl, elide=False)
    lean = _trace(compiled, pos, direction, wl, elide=True)
    _same(plain, lean)
    assert lean["stats"][_cuda.STAT_H2D_BYTES] == (24 + 56) * n  # optimistic upload + the full one of the second trace


def test_all_columns_constant(gpu):
    """A collimated monochromatic beam from a point: nothing to upload, every ray has the same initial state."""
    n = 400_000
    compiled, pos, direction, wl = _lsc_bundle(n)
    direction = np.tile(direction[3], (n, 1))
    plain = _trace(compiled, pos, direction, wl, elide=False)
    lean = _trace(compiled, pos, direction, wl, elide=True)
    _same(plain, lean)
    assert lean["stats"][_cuda.STAT_H2D_BYTES] == 0


def test_seeds_are_independent_runs(gpu):
    """Seeds s and s + 1 must not share photons (round 1: id = seed + i made them share n - 1): the difference of two
    runs' counts scatters like two independent binomial samples, sqrt(2 n p q), not like ~0."""
    scene = configs.lsc_default()
    n = 400_000
    runs = [pv.engine.simulate(scene, n, seed=s, record_every=0).data["rec_distinct"].astype(float) for s in (1, 2, 3, 4)]
    p = np.clip(np.mean(runs, axis=0) / n, 1e-9, 1 - 1e-9)
    sigma = np.sqrt(2 * n * p * (1 - p))
    busy = sigma > 30  # recorders with enough counts for the normal approximation
    assert busy.sum() >= 4
    z = np.concatenate([((runs[a] - runs[b]) / sigma)[busy] for a, b in ((0, 1), (1, 2), (2, 3), (0, 3))])
    assert np.abs(z).max() < 5.0           # not correlated the other way either
    assert 0.5 < np.std(z) < 1.6, np.std(z)  # shared photons would give ~0.002
    # and one seed is one run: bit-identical on repetition
    again = pv.engine.simulate(scene, n, seed=2, record_every=0).data
    assert (again["rec_distinct"] == runs[1]).all()


# ---- several GPUs -------------------------------------------------------------------------------------------

def _need(gpu, count):
    if gpu < count:
        pytest.skip(f"needs {count} GPUs, have {gpu}")


@pytest.mark.parametrize("name", ["lsc", "mixed"])
def test_devices_of_one_process_give_the_single_device_result(gpu, name):
    """pvt_trace_bundle_devices: tallies, stats and the event log (rows of every slice in place) equal one device's."""
    _need(gpu, 2)
    scene = scenes.SCENES[name]()
    n = 200_003
    one = pv.engine.simulate(scene, n, seed=11, record_every=1000, max_events=256, device=0)
    for devices in ([0, 1], list(range(min(gpu, 4)))):
        many = pv.engine.simulate(scene, n, seed=11, record_every=1000, max_events=256, devices=devices)
        _same(one.data, many.data)
        for key in ("counts", "kind", "hit", "container", "adjacent", "component", "source", "position", "direction",
                    "wavelength", "travelled", "duration", "normal"):
            assert (one.data[key] == many.data[key]).all(), key
        assert many.stats["rays"] == n and many.stats["steps"] == one.stats["steps"]
    by_workers = pv.engine.simulate(scene, n, seed=11, record_every=0, workers=2)
    _same(one.data, by_workers.data)


def test_devices_with_host_rays_and_constant_columns(gpu):
    _need(gpu, 2)
    n = 1_000_000
    compiled, pos, direction, wl = _lsc_bundle(n)
    one = _trace(compiled, pos, direction, wl, elide=True)
    two = _trace(compiled, pos, direction, wl, elide=True, devices=[0, 1])
    _same(one, two)
    assert two["stats"][_cuda.STAT_H2D_BYTES] == 24 * n


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _rank(rank, world, port, n, seed, out_dir):
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    os.environ["LOCAL_RANK"] = str(rank)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    scene = scenes.lsc()
    result = pv.engine.simulate(scene, n, seed=seed, record_every=0)  # sharded: engine/distributed.py
    assert result.num_rays == n  # the global count the reduced tallies refer to (round 1 returned the shard's)
    unseeded = pv.engine.simulate(scene, 20_000, seed=None, record_every=0)  # rank 0's draw is broadcast
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), unseeded=unseeded.data["rec_distinct"],
             **{k: result.data[k] for k in INTEGER_KEYS + ("rec_sums",)})
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_ranks_give_the_single_gpu_result(gpu, world, tmp_path):
    """One process per GPU over NCCL: the all-reduced tallies of 2 / 4 / 8 ranks equal one GPU's, on every rank."""
    _need(gpu, world)
    import torch.multiprocessing as mp

    n, seed = 300_001, 17
    mp.spawn(_rank, args=(world, _free_port(), n, seed, str(tmp_path)), nprocs=world, join=True)
    want = pv.engine.simulate(scenes.lsc(), n, seed=seed, record_every=0, device=0).data
    assert want["rec_distinct"].sum() > n // 2
    first = np.load(tmp_path / "rank0.npz")
    for rank in range(world):
        got = np.load(tmp_path / f"rank{rank}.npz")
        _same(want, got)
        assert (got["unseeded"] == first["unseeded"]).all()
        assert got["unseeded"].sum() >= 20_000  # every ray of the unseeded run exits or is lost, counted once
