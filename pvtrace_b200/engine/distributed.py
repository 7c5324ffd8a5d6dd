"""Photon sharding over GPUs: one process per GPU (torch.distributed), contiguous photon-index slices, ONE
all-reduce of the packed tally buffer at the end.  There is no exchange step inside the trace: ray i
depends only on (seed, i, scene), exactly like the reference's `prange` over rays (_kernel.pyx:1075) and its
consecutive-seed bundles (api.py:252-262).

Backends: with NCCL the context's packed float64 DEVICE buffer is reduced in place (NVLink / NVSwitch) and unpacked
into the accumulators before the one read-back; with gloo (CPU tests, world_size 2) the same buffer layout is reduced
as a host tensor.  (Several GPUs of ONE process need none of this: `engine.simulate(..., workers=k)` ->
`pvt_trace_bundle_devices`.)
"""
import hashlib
import os

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


def local_device():
    """CUDA ordinal this rank traces on: LOCAL_RANK (torchrun), else the rank, modulo the devices of the node."""
    import torch.distributed as dist

    from pvtrace_b200.engine import _cuda

    rank = dist.get_rank() if dist.is_initialized() else 0
    return int(os.environ.get("LOCAL_RANK", rank)) % max(_cuda.device_count(), 1)


def all_reduce_tallies(data, device=None):
    """Sum host-side tallies of every rank (any backend); every rank gets the total.  Used after the drop-in host call
    (`_cuda.trace_bundle`), whose tallies are already on the host; `simulate_sharded` reduces on the device instead."""
    import torch
    import torch.distributed as dist

    packed = torch.from_numpy(pack_host(data))
    if dist.get_backend() == "nccl":
        packed = packed.to(torch.device("cuda", local_device() if device is None else int(device)))
    dist.all_reduce(packed, op=dist.ReduceOp.SUM)
    return unpack_host(packed.cpu().numpy(), data)


class _DevicePacked:
    """torch view (no copy) of the library's packed tally buffer."""

    def __init__(self, ptr, count):
        self.__cuda_array_interface__ = {"shape": (count,), "typestr": "<f8", "data": (ptr, False), "version": 3}


_contexts = {}  # (device, digest of the tables) -> _cuda.Context: repeated sharded runs of one scene reuse the upload


def _context_for(compiled, emitter, device):
    from pvtrace_b200.engine import _cuda

    digest = hashlib.sha1()
    for _, attr, dtype in _cuda._SCENE_TABLES:
        value = getattr(compiled, attr, None)
        if value is not None:
            digest.update(np.ascontiguousarray(value, dtype=dtype).tobytes())
    if emitter is not None:
        for field, dtype in _cuda._EMIT_TABLES:
            digest.update(np.ascontiguousarray(getattr(emitter, field), dtype=dtype).tobytes())
    key = (device, digest.hexdigest())
    if key not in _contexts:
        for stale in [k for k in _contexts if k[0] == device]:
            _contexts.pop(stale).close()
        _contexts[key] = _cuda.Context(compiled, emitter, device)
    return _contexts[key]


def simulate_sharded(scene, compiled, num_rays, seed, maxsteps, max_events, emit_method, record_every, *,
                     rng="philox", first_index=0):
    """Body of `engine.simulate` under an initialised process group: every rank traces its slice on its own GPU
    (LOCAL_RANK), the tallies are all-reduced, and each rank keeps the event log of its own slice.

    `seed` None: rank 0's draw is broadcast, so every rank traces the same run.  The result's `num_rays` is the
    GLOBAL count (what the reduced tallies refer to); `sources` / `recorded_indices` cover this rank's slice, whose first
    ray has global index `result.first_index`."""
    import torch
    import torch.distributed as dist

    from pvtrace_b200.engine import _cuda
    from pvtrace_b200.engine.api import EngineResult
    from pvtrace_b200.engine.compiler import EMIT_METHODS, compile_emitter
    from pvtrace_b200.engine.emit import LightNames, emit_bundle_host

    rank, world = dist.get_rank(), dist.get_world_size()
    nccl = dist.get_backend() == "nccl"
    device = local_device()
    if seed is None:
        box = torch.tensor([np.random.randint(0, 2 ** 31 - 1) if rank == 0 else 0], dtype=torch.int64)
        if nccl:
            box = box.to(torch.device("cuda", device))
        dist.broadcast(box, src=0)
        seed = int(box.item())
    start, count = shard_range(num_rays, rank, world)
    emitter = compile_emitter(scene)
    if emitter is not None:
        positions = directions = wavelengths = None
        sources = LightNames(emitter.light_names, count, first_index + start)
    else:
        # host delegates: every rank draws the whole bundle from the same numpy state and keeps its slice
        positions, directions, wavelengths, sources = emit_bundle_host(scene, num_rays)
        positions, directions = positions[start:start + count], directions[start:start + count]
        wavelengths, sources = wavelengths[start:start + count], sources[start:start + count]
    if nccl and emitter is not None:
        # device path: trace into the context's accumulators, reduce its packed DEVICE buffer in place, read once
        ctx = _context_for(compiled, emitter, device)
        with torch.cuda.device(device):
            stream = torch.cuda.current_stream().cuda_stream
            tic, toc = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            tic.record()
            ctx.reset(stream=stream)
            ctx.trace(count, int(seed), first_index=first_index + start, maxsteps=int(maxsteps), max_events=int(max_events),
                      emit_method=EMIT_METHODS[emit_method], record_every=int(record_every),
                      rng_mode=_cuda.RNG_MODES[rng], stream=stream)
            ptr, n_packed = ctx.pack_tallies(stream=stream)
            if n_packed:
                packed = torch.as_tensor(_DevicePacked(ptr, n_packed), device=torch.device("cuda", device))
                dist.all_reduce(packed, op=dist.ReduceOp.SUM)
                ctx.unpack_tallies(stream=stream)
            toc.record()
            data = ctx.read(stream=stream)
            elapsed = tic.elapsed_time(toc) * 1e-3
    else:
        data, elapsed = _cuda.trace_bundle(
            compiled, positions, directions, wavelengths, int(seed), int(maxsteps), int(max_events),
            EMIT_METHODS[emit_method], 0, int(record_every), emitter=emitter, n=count, first_index=first_index + start,
            rng_mode=_cuda.RNG_MODES[rng], device=device, return_elapsed=True)
        data = all_reduce_tallies(data, device)
    return EngineResult(compiled, data, sources, max_events, record_every, elapsed, num_rays=int(num_rays),
                        first_index=first_index + start)
