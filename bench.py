#!/usr/bin/env python
"""Benchmark of the photon-tracing hot path: photons/s on the 5x5x1 cm LSC (BASELINE.json `metric`).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--photons P] [--config NAME] [--impl reference]

One "step" = one pass of the hot path over one bundle of P photons (default 10^7, BASELINE.json configs[1]:
LSC((5,5,1)) default Lumogen-F-Red scene, emit_method kT, maxsteps 1000, record: true-style recorders, no event
log).  Under torchrun every rank traces its own P photons (weak scaling, disjoint photon-index ranges of one
global run) and the step ends with ONE all-reduce of the packed tally buffer; the reduced tallies are CHECKED
(every photon of every rank ends exactly once: exit + lost == world * P).

Printed JSON (rank 0):
  value      photons/s with the initial rays ALREADY RESIDENT in HBM (device time, CUDA events, max over ranks)
  e2e        photons/s through the reference-facing call `_cuda.trace_bundle` == pvt_trace_bundle (C ABI) with
             HOST (pinned) ray arrays: H2D of the ray columns + trace + D2H of the tallies inside the timed region
             (h2d_bytes_per_step is what crossed PCIe: columns that hold one value for every ray are not uploaded)
  roofline   dominant kernel vs the HBM roofline: achieved = photon steps (device counter) x 192 B / kernel time,
             peak = MEASURED_PEAKS.json hbm_gbs; `traffic` = DRAM bytes per launch from the committed ncu capture
  roofline_fp64  the same kernel's fp64 rate against the chip's measured fp64 FMA throughput (the bound that is
             physically there for it: the photon state never leaves the SM)
  intersect_stage  the ray/primitive stage alone against the HBM peak (north_star's named roofline), 60 B per ray
  extra      the other BASELINE configs (device-resident rays), each with its steps per photon and CPU baseline
  cpu_baseline  the reference's own compiled kernel (oracle/_ref, built from /root/reference) on the host cores
`--impl reference` times that CPU kernel alone: same metric, same config, same photons per step.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ALGORITHMIC_BYTES_PER_STEP = 192  # 96 B photon state read + written once per photon step
INTERSECT_BYTES_PER_RAY = 60      # position + direction in, t0 + one word of packed ids out
WORKLOADS = {
    "hello_world": "glass sphere (n 1.5, r 1) in an air sphere, 555 nm cone(pi/8) light (examples/hello_world.py)",
    "lsc_default": "LSC((5,5,1)) Lumogen F Red 305 x10 cm^-1 + 0.1 cm^-1 background, 555 nm cone(20 deg) light, kT "
                   "emission, record:true recorders (7808 bins)",
    "nested_cylinders": "rotated cylinder with a protruding child cylinder in an air sphere, cone(30 deg) light "
                        "(examples/nested_cylinders.py)",
    "lsc_coated": "LSC((5,5,1)) + solar cells on the four edges + back-surface mirror (pvtrace/device/lsc.py:280-291)",
    "validation": "4.8x1.8x0.26 cm Fluro Red LSC under the Oriel lamp spectrum, redshift emission "
                  "(examples/Validation.ipynb)",
}
EXTRA_PHOTONS = {"hello_world": 10 ** 6, "nested_cylinders": 10 ** 6, "lsc_coated": 12_500_000, "validation": 10 ** 7,
                 "lsc_default": 10 ** 7}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--photons", type=float, default=1e7, help="photons per GPU per step")
    ap.add_argument("--config", default="lsc_default", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-sample", type=float, default=2e6, help="photons per CPU-baseline repetition (b200 arm)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the other BASELINE configs")
    return ap.parse_args()


def build_scene(name):
    import pvtrace_b200 as pv
    from pvtrace_b200.device import configs
    from pvtrace_b200.engine.compiler import EMIT_METHODS

    build, kw = configs.CONFIGS[name]
    scene = build()
    return scene, pv.engine.compile_scene(scene), pv.engine.compile_emitter(scene), EMIT_METHODS[kw["emit_method"]]


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "MEASURED_PEAKS.json"
    return 6650.0, "fallback (B200_PROFILING.md)"


def profile_constants(config):
    """Per-launch figures that only a profiler can give -- DRAM bytes and fp64 instructions per photon step of the trace
    kernel -- from the committed ncu capture (profiles/r2_traffic.json names the capture it was read from)."""
    path = os.path.join(ROOT, "profiles", "r2_traffic.json")
    if not os.path.exists(path):
        return {}
    with open(path) as fh:
        table = json.load(fh)
    entry = dict(table.get(config, {}))
    entry["source"] = table.get("source")
    return entry


def bind_to_gpu_numa_node(index):
    """Run this rank (and first-touch its pinned buffers) on the CPUs next to its GPU.  Returns what was done."""
    try:
        import pynvml

        pynvml.nvmlInit()
        handle = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(handle, words)
        cpus = {64 * w + b for w, word in enumerate(mask) for b in range(64) if (word >> b) & 1}
        allowed = cpus & os.sched_getaffinity(0)
        if allowed:
            os.sched_setaffinity(0, allowed)
            return f"{len(allowed)} cpus near gpu {index}"
    except Exception as exc:  # the binding is an optimisation, never a requirement
        return f"not bound ({type(exc).__name__})"
    return "not bound"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons while the timed region runs."""

    QUERY = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self._halt = index, [], threading.Event()

    def run(self):
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                      "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL,
                                     text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([f.strip() for f in out.split(",")])
            except Exception:
                pass
            self._halt.wait(0.02)

    def stop(self):
        self._halt.set()
        self.join(timeout=5)
        sm = [float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if s[1].replace(".", "").isdigit()]
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        reasons = sorted({n for s in self.samples for n, v in zip(names, s[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.samples)}


def time_cpu_kernel(compiled, emitter, method, photons, reps, threads):
    """photons/s of the reference's compiled CPU kernel (oracle/_ref) on `threads` host threads, best of reps.  Scenes
    the reference's engine cannot express (facet tables: config 4) go through the oracle port."""
    from oracle import pvt_oracle, ref_loader

    kernel = ref_loader.load_ref_kernel()
    if int(getattr(compiled, "n_facets", 0)) > 0:
        kernel = None
    kind = "reference" if kernel is not None else "port"
    n = int(photons)
    pos, direction, wl = pvt_oracle.emit_bundle(emitter, n, seed=1)
    best = None
    for _ in range(reps):
        tic = time.perf_counter()
        if kernel is not None:
            kernel.trace_bundle(compiled, pos, direction, wl, 1, 1000, 128, method, threads, 0)
        else:
            pvt_oracle.trace_bundle(compiled, pos, direction, wl, 1, 1000, 128, method, threads, 0, rng_mode=1)
        dt = time.perf_counter() - tic
        best = dt if best is None else min(best, dt)
    return n / best, kind, best


def cpu_sample_text(kind, photons, threads, reps):
    what = "pvtrace/engine/_kernel.pyx compiled -O3 -fopenmp (oracle/_ref)" if kind == "reference" else \
        "oracle/pvt_oracle.c (the reference kernel was not built here, or cannot express the scene)"
    return f"{int(photons)} photons of the same scene, best of {reps}, {threads} OpenMP threads, {what}"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    scene, compiled, emitter, method = build_scene(args.config)
    threads = os.cpu_count() or 1
    photons = int(args.photons)  # the same bundle as the b200 arm: same config, same photons per step
    kind = "reference"
    for _ in range(max(args.warmup, 0)):
        time_cpu_kernel(compiled, emitter, method, min(photons, 500000), 1, threads)
    times = []
    for _ in range(args.steps):
        rate, kind, dt = time_cpu_kernel(compiled, emitter, method, photons, 1, threads)
        times.append(dt)
    total = sum(times)
    value = photons * args.steps / total
    line = {
        "impl": "reference", "metric": "photons/sec on 5x5x1 cm LSC", "value": value, "unit": "photons/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{args.config}: {WORKLOADS[args.config]}, maxsteps 1000, record_every 0",
                   "photons_per_gpu_per_step": photons,
                   "note": "CPU arm: rank 0 only, the whole bundle per step (warm-up steps are shorter)"},
        "cpu_baseline": {"value": value, "unit": "photons/s", "cores": threads, "kind": kind,
                         "sample": f"{photons} photons per step, " + cpu_sample_text(kind, photons, threads, 1)},
        "e2e": {"value": value, "unit": "photons/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def time_resident(ctx, torch, n, seed, first_index, method, reps):
    """ms per trace of n device-emitted, device-resident rays (CUDA events) and the photon steps of one bundle."""
    pos = torch.empty((n, 3), dtype=torch.float64, device="cuda")
    dirs = torch.empty((n, 3), dtype=torch.float64, device="cuda")
    wl = torch.empty(n, dtype=torch.float64, device="cuda")
    stream = torch.cuda.current_stream()
    sptr = stream.cuda_stream
    ctx.emit(pos.data_ptr(), dirs.data_ptr(), wl.data_ptr(), n, seed=seed, first_index=first_index, stream=sptr)
    for _ in range(2):
        ctx.reset(stream=sptr)
        ctx.trace(n, seed, d_positions=pos.data_ptr(), d_directions=dirs.data_ptr(), d_wavelengths=wl.data_ptr(),
                  first_index=first_index, emit_method=method, record_every=0, stream=sptr)
    times = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ctx.reset(stream=sptr)
        a.record(stream)
        ctx.trace(n, seed, d_positions=pos.data_ptr(), d_directions=dirs.data_ptr(), d_wavelengths=wl.data_ptr(),
                  first_index=first_index, emit_method=method, record_every=0, stream=sptr)
        b.record(stream)
        torch.cuda.synchronize()
        times.append(a.elapsed_time(b))
    data = ctx.read(stream=sptr)
    return float(np.mean(times)), data


def run_b200(args):
    import torch
    import torch.distributed as dist

    from pvtrace_b200.engine import _cuda

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (pvtrace_b200 has no CPU fallback)")
    binding = bind_to_gpu_numa_node(local) if world > 1 else "single rank"
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    scene, compiled, emitter, method = build_scene(args.config)
    names = list(compiled.recorder_names)
    n = int(args.photons)
    seed = 1
    first_index = rank * n  # disjoint photon-index ranges of one global run of world * n photons
    ctx = _cuda.Context(compiled, emitter, local)
    stream = torch.cuda.current_stream()
    sptr = stream.cuda_stream

    # ---- device-resident leg: initial rays live in HBM before the timed region ---------------------------
    pos = torch.empty((n, 3), dtype=torch.float64, device="cuda")
    dirs = torch.empty((n, 3), dtype=torch.float64, device="cuda")
    wl = torch.empty(n, dtype=torch.float64, device="cuda")
    ctx.emit(pos.data_ptr(), dirs.data_ptr(), wl.data_ptr(), n, seed=seed, first_index=first_index, stream=sptr)

    class _Packed:  # zero-copy torch view of the library's packed tally buffer (for the NCCL all-reduce)
        def __init__(self, ptr, count):
            self.__cuda_array_interface__ = {"shape": (count,), "typestr": "<f8", "data": (ptr, False), "version": 3}

    def one_step(events=None):
        ctx.reset(stream=sptr)
        if events is not None:
            events[0].record(stream)
        ctx.trace(n, seed, d_positions=pos.data_ptr(), d_directions=dirs.data_ptr(), d_wavelengths=wl.data_ptr(),
                  first_index=first_index, emit_method=method, record_every=0, stream=sptr)
        if events is not None:
            events[1].record(stream)
        if world > 1:
            ptr, count = ctx.pack_tallies(stream=sptr)
            packed = torch.as_tensor(_Packed(ptr, count), device="cuda")
            dist.all_reduce(packed, op=dist.ReduceOp.SUM)
            ctx.unpack_tallies(stream=sptr)

    def fence():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def ended(data):
        """Rays that ended in the `exit` recorder of the world or a `lost` recorder: every ray does, exactly once,
        unless the step budget cut it (maxsteps 1000: a handful of TIR-trapped rays in lossless scenes)."""
        return int(sum(int(data["rec_distinct"][k]) for k, name in enumerate(names) if name == "exit" or name.endswith("lost")))

    warmup = max(args.warmup, 3)
    for _ in range(warmup):
        one_step()
    fence()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    kernel_events = []
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record(stream)
    for _ in range(args.steps):
        pair = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        one_step(pair)
        kernel_events.append(pair)
    stop.record(stream)
    fence()
    clocks = sampler.stop() if sampler else None
    total_ms = start.elapsed_time(stop)
    kernel_ms = float(np.mean([a.elapsed_time(b) for a, b in kernel_events]))
    data = ctx.read(stream=sptr)  # tallies of the LAST step: reduced over the ranks when world > 1
    steps_per_bundle = int(data["stats"][_cuda.STAT_STEPS])  # local device counter == one bundle of this rank
    reduced_ended = ended(data)
    if "exit" in names:  # SCALE checks results, not only time
        assert world * n - reduced_ended <= max(3, int(2e-3 * world * n)) and reduced_ended <= world * n, \
            f"reduced tallies account for {reduced_ended} of {world * n} photons"
    launches_per_step = 1 + (2 if world > 1 else 0)

    # ---- the intersect stage on its own (north_star: "HBM roofline for the intersect kernel") ----------------
    t0 = torch.empty(n, dtype=torch.float64, device="cuda")
    ids = torch.empty(n, dtype=torch.int32, device="cuda")
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device="cuda")  # 256 MB > L2
    for _ in range(3):
        ctx.intersect_packed(pos.data_ptr(), dirs.data_ptr(), n, t0.data_ptr(), ids.data_ptr(), stream=sptr)
    # Two flushes.  Writing 256 MB evicts the rays but leaves the L2 full of DIRTY lines of the flush buffer, whose
    # write-back (up to 126 MB) then shares the DRAM with the kernel that is being timed; reading the buffer back
    # afterwards leaves the L2 cold AND clean -- what ncu's own cache control gives (profiles/r2_intersect_kernel.txt).
    # Both are reported; the roofline fraction is the cold-and-clean one.
    intersect_times = {"clean": [], "dirty": []}
    for rep in range(10):
        kind = "clean" if rep % 2 == 0 else "dirty"
        flush.zero_()
        if kind == "clean":
            flush.sum()
        ia, ib = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ia.record(stream)
        ctx.intersect_packed(pos.data_ptr(), dirs.data_ptr(), n, t0.data_ptr(), ids.data_ptr(), stream=sptr)
        ib.record(stream)
        torch.cuda.synchronize()
        intersect_times[kind].append(ia.elapsed_time(ib))
    intersect_ms = float(np.mean(intersect_times["clean"]))
    intersect_dirty_ms = float(np.mean(intersect_times["dirty"]))
    del t0, ids, flush

    # ---- the user-facing call with on-device emission (no ray arrays cross PCIe) --------------------------------
    def emit_step():
        return _cuda.trace_bundle(compiled, None, None, None, seed, 1000, 128, method, 0, 0, emitter=emitter, n=n,
                                  first_index=first_index, device=local, return_elapsed=True)

    emit_step()
    fence()
    tic = time.perf_counter()
    for _ in range(3):
        emit_step()
    fence()
    emit_s = (time.perf_counter() - tic) / 3

    # ---- end-to-end leg: host (pinned) rays through the drop-in call -----------------------------------------
    h_pos = torch.empty((n, 3), dtype=torch.float64).pin_memory()
    h_dir = torch.empty((n, 3), dtype=torch.float64).pin_memory()
    h_wl = torch.empty(n, dtype=torch.float64).pin_memory()
    h_pos.copy_(pos); h_dir.copy_(dirs); h_wl.copy_(wl)
    torch.cuda.synchronize()
    np_pos, np_dir, np_wl = h_pos.numpy(), h_dir.numpy(), h_wl.numpy()

    def e2e_step():
        out, elapsed = _cuda.trace_bundle(compiled, np_pos, np_dir, np_wl, seed, 1000, 128, method, 0, 0,
                                          first_index=first_index, device=local, return_elapsed=True)
        if world > 1:
            from pvtrace_b200.engine import distributed

            out = distributed.all_reduce_tallies(out, local)
        return out, elapsed

    e2e_steps = max(3, min(args.steps, 5))
    for _ in range(2):
        e2e_step()
    fence()
    tic = time.perf_counter()
    for _ in range(e2e_steps):
        out, _ = e2e_step()
    fence()
    e2e_s = (time.perf_counter() - tic) / e2e_steps
    h2d = int(out["stats"][_cuda.STAT_H2D_BYTES])  # what crossed PCIe: columns that are constant over the bundle do not
    d2h = int(sum(out[k].nbytes for k in ("rec_distinct", "rec_crossings", "rec_sums", "rec_bins", "stats")))
    e2e_ended = ended(out)
    if "exit" in names:
        assert world * n - e2e_ended <= max(3, int(2e-3 * world * n)) and e2e_ended <= world * n, (e2e_ended, world * n)

    # ---- reduce timings over ranks (max) ---------------------------------------------------------------------
    times = torch.tensor([total_ms, kernel_ms, e2e_s, emit_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    total_ms, kernel_ms, e2e_s, emit_s = (float(v) for v in times.tolist())

    if rank == 0:
        peak, peak_src = measured_peaks()
        ms_per_step = total_ms / args.steps
        value = world * n / (ms_per_step * 1e-3)
        achieved = steps_per_bundle * ALGORITHMIC_BYTES_PER_STEP / (kernel_ms * 1e-3) / 1e9
        prof = profile_constants(args.config)
        fp64_peak = _cuda.measure_fp64_peak(local)
        line = {
            "metric": "photons/sec on 5x5x1 cm LSC", "value": value, "unit": "photons/s", "n_gpus": world,
            "steps": args.steps, "warmup": warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{args.config}: {WORKLOADS[args.config]}, maxsteps 1000, record_every 0",
                       "photons_per_gpu_per_step": n, "photon_steps_per_gpu_per_step": steps_per_bundle,
                       "l2": "initial rays 56 B/photon (560 MB at 1e7) exceed the 126 MB L2",
                       "sharding": f"photon index ranges, {world} rank(s), one all-reduce of the packed tallies per step",
                       "tallies_checked": f"exit + lost == {reduced_ended} of {world * n} photons after the all-reduce",
                       "cpu_binding": binding},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": prof.get("dram_bytes_per_launch"), "traffic_source": prof.get("source"),
                         "peak_source": peak_src, "kernel": "wavefront_kernel",
                         "kernel_ms": kernel_ms, "bytes_per_step": ALGORITHMIC_BYTES_PER_STEP,
                         "note": "the photon state lives in shared memory: DRAM traffic is the ray read; what bounds the "
                                 "kernel is dependent-issue latency (DESIGN.md section 7), see roofline_fp64"},
            "e2e": {"value": world * n / e2e_s, "unit": "photons/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": 1e3 * e2e_s,
                    "host_bytes_per_step": int(np_pos.nbytes + np_dir.nbytes + np_wl.nbytes)},
            "e2e_device_emission": {"value": world * n / emit_s, "unit": "photons/s", "ms_per_step": 1e3 * emit_s,
                                    "note": "engine.simulate path for built-in lights: rays sampled in the kernel, "
                                            "0 B H2D, tallies D2H"},
            "intersect_stage": {"kernel": "intersect_ring_kernel", "ms": intersect_ms, "bytes_per_ray": INTERSECT_BYTES_PER_RAY,
                                "achieved_gbs": INTERSECT_BYTES_PER_RAY * n / (intersect_ms * 1e-3) / 1e9,
                                "frac_of_hbm_peak": INTERSECT_BYTES_PER_RAY * n / (intersect_ms * 1e-3) / 1e9 / peak,
                                "l2": "flushed between repetitions: 256 MB written, then read back (cold and clean)",
                                "ms_dirty_l2": intersect_dirty_ms,
                                "frac_dirty_l2": INTERSECT_BYTES_PER_RAY * n / (intersect_dirty_ms * 1e-3) / 1e9 / peak,
                                "dirty_l2": "256 MB written and not read back: the flush buffer's dirty lines are "
                                            "written to DRAM while the kernel runs (round 1 and 2's earlier protocol)"},
            "gpu_launches": launches_per_step * args.steps,
            "clocks": clocks,
        }
        if prof.get("fp64_flops_per_photon_step"):
            rate = prof["fp64_flops_per_photon_step"] * steps_per_bundle / (kernel_ms * 1e-3) * 1e-12
            line["roofline_fp64"] = {"achieved": rate, "peak": fp64_peak, "unit": "TFLOP/s", "frac": rate / fp64_peak,
                                     "flops_per_photon_step": prof["fp64_flops_per_photon_step"],
                                     "peak_source": "pvt_measure_fp64_peak (DFMA micro-benchmark, this run)",
                                     "flops_source": prof.get("source")}
        else:
            line["roofline_fp64"] = {"peak": fp64_peak, "unit": "TFLOP/s",
                                     "peak_source": "pvt_measure_fp64_peak (DFMA micro-benchmark, this run)"}
        threads = os.cpu_count() or 1
        if not args.no_cpu_baseline:
            sample = int(min(args.cpu_sample, n))
            rate, kind, dt = time_cpu_kernel(compiled, emitter, method, sample, 3, threads)
            line["cpu_baseline"] = {"value": rate, "unit": "photons/s", "cores": threads, "kind": kind,
                                    "sample": cpu_sample_text(kind, sample, threads, 3)}
        if not args.no_extra and world == 1:
            extra = {}
            for name in WORKLOADS:
                if name == args.config:
                    continue
                _, c2, e2, m2 = build_scene(name)
                photons = EXTRA_PHOTONS[name]
                with _cuda.Context(c2, e2, local) as other:
                    ms, d2 = time_resident(other, torch, photons, seed, 0, m2, 3)
                entry = {"workload": WORKLOADS[name], "photons": photons, "ms": ms, "photons_per_s": photons / (ms * 1e-3),
                         "steps_per_photon": float(d2["stats"][_cuda.STAT_STEPS]) / photons}
                if not args.no_cpu_baseline:
                    sample = int(min(1e6, photons))
                    rate, kind, dt = time_cpu_kernel(c2, e2, m2, sample, 2, threads)
                    entry["cpu_baseline"] = {"value": rate, "unit": "photons/s", "cores": threads, "kind": kind,
                                             "sample": cpu_sample_text(kind, sample, threads, 2)}
                extra[name] = entry
            line["extra"] = extra
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
