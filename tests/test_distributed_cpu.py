"""Host logic of the multi-GPU path on CPU: photon-range sharding and the single all-reduce of the packed tally
buffer, world_size 2 over gloo.  (The device trace itself is covered by the gpu tests; here each rank contributes
tallies computed by the CPU oracle for its own index range and the reduced result must equal one full run.)"""
import os
import socket

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

import pvtrace_b200 as pv
from pvtrace_b200.engine import distributed
from tests import scenes


def test_shard_ranges_partition_the_bundle():
    for n in (0, 1, 7, 1000, 10 ** 7 + 3):
        for world in (1, 2, 3, 8):
            spans = [distributed.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == n
            for (s0, c0), (s1, _) in zip(spans, spans[1:]):
                assert s0 + c0 == s1
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1


def test_pack_unpack_roundtrip():
    data = {"rec_distinct": np.array([3, 2 ** 40], dtype=np.int64), "rec_crossings": np.array([5, 7], dtype=np.int64),
            "rec_sums": np.arange(16, dtype=np.float64).reshape(2, 4, 2) * 0.1, "rec_bins": np.arange(9, dtype=np.int64)}
    back = distributed.unpack_host(distributed.pack_host(data), {k: v.copy() for k, v in data.items()})
    for key in data:
        assert (back[key] == data[key]).all() and back[key].dtype == data[key].dtype


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, seed, out_path):
    from oracle import pvt_oracle

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    assert distributed.is_active()
    scene = scenes.lsc()
    compiled, emitter = pv.engine.compile_scene(scene), pv.engine.compile_emitter(scene)
    start, count = distributed.shard_range(n, rank, world)
    data = pvt_oracle.trace_bundle(compiled, None, None, None, seed, 1000, 64, 0, 1, 0, emitter=emitter, n=count,
                                   first_index=start)
    data = distributed.all_reduce_tallies(data)
    if rank == 0:
        np.savez(out_path, **{k: data[k] for k in ("rec_distinct", "rec_crossings", "rec_sums", "rec_bins")})
    dist.destroy_process_group()


def test_two_rank_gloo_reduce_equals_single_run(tmp_path):
    from oracle import pvt_oracle

    n, seed = 6001, 21
    out_path = str(tmp_path / "reduced.npz")
    mp.spawn(_worker, args=(2, _free_port(), n, seed, out_path), nprocs=2, join=True)
    got = np.load(out_path)
    scene = scenes.lsc()
    compiled, emitter = pv.engine.compile_scene(scene), pv.engine.compile_emitter(scene)
    want = pvt_oracle.trace_bundle(compiled, None, None, None, seed, 1000, 64, 0, 1, 0, emitter=emitter, n=n)
    for key in ("rec_distinct", "rec_crossings", "rec_bins"):
        assert (got[key] == want[key]).all(), key
    np.testing.assert_allclose(got["rec_sums"], want["rec_sums"], rtol=1e-12)
    assert want["rec_distinct"].sum() > n // 2
