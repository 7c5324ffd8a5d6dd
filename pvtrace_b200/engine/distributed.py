"""Photon sharding over GPUs: one process per GPU (torch.distributed), contiguous photon-index slices, ONE
all-reduce of the packed tally buffer at the end.  There is no exchange step inside the trace: ray i
depends only on (seed + i, scene), exactly like the reference's `prange` over rays (_kernel.pyx:1075) and its
consecutive-seed bundles (api.py:252-262).

Backends: with NCCL the packed float64 buffer is reduced in place on the device (NVLink / NVSwitch); with gloo
(CPU tests, world_size 2) the same buffer layout is reduced as a host tensor.
"""
import numpy as np


def is_active():
    try:
        import torch.distributed as dist
    except ImportError:  # pragma: no cover
        return False
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def shard_range(num_rays, rank, world_size):
    """Contiguous slice [start, start + count) of the photon index range owned by `rank`."""
    base, extra = divmod(int(num_rays), int(world_size))
    start = rank * base + min(rank, extra)
    return start, base + (1 if rank < extra else 0)


def pack_host(data):
    """Host arrays -> one float64 vector [distinct | crossings | sums | bins] (integers stay exact below 2^53)."""
    return np.concatenate([data["rec_distinct"].astype(np.float64), data["rec_crossings"].astype(np.float64),
                           np.asarray(data["rec_sums"], dtype=np.float64).ravel(),
                           data["rec_bins"].astype(np.float64)])


def unpack_host(packed, data):
    r, b = len(data["rec_distinct"]), len(data["rec_bins"])
    data["rec_distinct"] = np.rint(packed[:r]).astype(np.int64)
    data["rec_crossings"] = np.rint(packed[r:2 * r]).astype(np.int64)
    data["rec_sums"] = packed[2 * r:10 * r].reshape(r, 4, 2).copy()
    data["rec_bins"] = np.rint(packed[10 * r:10 * r + b]).astype(np.int64)
    return data


def all_reduce_tallies(data):
    """Sum the tallies of every rank (any backend); every rank gets the total."""
    import torch
    import torch.distributed as dist

    packed = torch.from_numpy(pack_host(data))
    if dist.get_backend() == "nccl":
        packed = packed.cuda()
    dist.all_reduce(packed, op=dist.ReduceOp.SUM)
    return unpack_host(packed.cpu().numpy(), data)


def simulate_sharded(scene, compiled, num_rays, seed, maxsteps, max_events, emit_method, record_every, *,
                     rng="philox", first_index=0):
    """Body of `engine.simulate` under an initialised process group: every rank traces its slice on its own GPU
    (LOCAL_RANK), tallies are all-reduced, and each rank keeps the event log of its own slice."""
    import os

    import torch.distributed as dist

    from pvtrace_b200.engine import _cuda
    from pvtrace_b200.engine.api import EngineResult
    from pvtrace_b200.engine.compiler import EMIT_METHODS, compile_emitter
    from pvtrace_b200.engine.emit import LightNames, emit_bundle_host

    rank, world = dist.get_rank(), dist.get_world_size()
    start, count = shard_range(num_rays, rank, world)
    device = int(os.environ.get("LOCAL_RANK", rank)) % max(_cuda.device_count(), 1)
    emitter = compile_emitter(scene)
    if emitter is not None:
        positions = directions = wavelengths = None
        sources = LightNames(emitter.light_names, count, first_index + start)
    else:
        # host delegates: every rank draws the whole bundle from the same numpy state and keeps its slice
        positions, directions, wavelengths, sources = emit_bundle_host(scene, num_rays)
        positions, directions = positions[start:start + count], directions[start:start + count]
        wavelengths, sources = wavelengths[start:start + count], sources[start:start + count]
    data, elapsed = _cuda.trace_bundle(
        compiled, positions, directions, wavelengths, seed, int(maxsteps), int(max_events), EMIT_METHODS[emit_method],
        0, int(record_every), emitter=emitter, n=count, first_index=first_index + start,
        rng_mode=_cuda.RNG_MODES[rng], device=device, return_elapsed=True)
    data = all_reduce_tallies(data)
    return EngineResult(compiled, data, sources, max_events, record_every, elapsed)
