"""Run one config's trace from device-resident rays a few times (target of the ncu captures in profiles/)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pvtrace_b200 as pv
from pvtrace_b200.device import configs
from pvtrace_b200.engine import _cuda
from pvtrace_b200.engine.compiler import EMIT_METHODS

name = sys.argv[1] if len(sys.argv) > 1 else "lsc_default"
n = int(float(sys.argv[2])) if len(sys.argv) > 2 else 10_000_000
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
build, kw = configs.CONFIGS[name]
scene = build()
ctx = _cuda.Context(pv.engine.compile_scene(scene), pv.engine.compile_emitter(scene), 0)
pos = torch.empty((n, 3), dtype=torch.float64, device="cuda")
dirs = torch.empty((n, 3), dtype=torch.float64, device="cuda")
wl = torch.empty(n, dtype=torch.float64, device="cuda")
ctx.emit(pos.data_ptr(), dirs.data_ptr(), wl.data_ptr(), n, seed=1)
torch.cuda.synchronize()
for rep in range(reps):
    ctx.reset()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    ctx.trace(n, 1, d_positions=pos.data_ptr(), d_directions=dirs.data_ptr(), d_wavelengths=wl.data_ptr(),
              emit_method=EMIT_METHODS[kw["emit_method"]])
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    d = ctx.read()
    prof = d["stats"][4:8]
    extra = ""
    if int(prof[2]) > 10**15:  # -DPVT_PROFILE_STAGES=2: globaltimer stamps
        M = (1 << 64) - 1
        t0, e_min, e_max, dur_sum = ~int(prof[0]) & M, ~int(prof[1]) & M, int(prof[2]), int(prof[3])
        ctas = 148
        extra = f"  CTA ends: first {(e_min - t0) / 1e6:.3f} ms, last {(e_max - t0) / 1e6:.3f} ms, mean CTA duration {dur_sum / ctas / 1e6:.3f} ms"
    elif prof.sum() > 0:  # library built with PVT_NVCC_FLAGS=-DPVT_PROFILE_STAGES: warp 0's cycles per stage, summed over CTAs
        extra = "  stage1/bar1/stage2/bar2 % of warp-0 time: " + " ".join(f"{100 * v / prof.sum():.1f}" for v in prof)
    print(f"{name} n={n} rep={rep} {ms:.3f} ms {n / ms / 1e3:.1f} Mphot/s steps/photon {d['stats'][0] / n:.3f}{extra}", flush=True)
