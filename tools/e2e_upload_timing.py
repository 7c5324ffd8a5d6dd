"""Host-ray end to end: PVT_DEBUG_TIMING=1 python tools/e2e_upload_timing.py prints when the upload and the trace ended
(library's stderr line) beside the wall time of the call and the device time the library reports."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pvtrace_b200 as pv
from pvtrace_b200.device import configs
from pvtrace_b200.engine import _cuda
n = 10_000_000
scene = configs.lsc_default()
compiled, emitter = pv.engine.compile_scene(scene), pv.engine.compile_emitter(scene)
h = [torch.empty((n, 3), dtype=torch.float64).pin_memory(), torch.empty((n, 3), dtype=torch.float64).pin_memory(),
     torch.empty(n, dtype=torch.float64).pin_memory()]
d = [torch.empty_like(t, device="cuda") for t in h]
ctx = _cuda.Context(compiled, emitter, 0)
ctx.emit(d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), n, seed=1)
for a, b in zip(h, d): a.copy_(b)
torch.cuda.synchronize()
arrs = [t.numpy() for t in h]
for elide in ("1", "0"):
    os.environ["PVT_ELIDE_CONSTANT"] = elide
    for rep in range(4):
        t0 = time.perf_counter()
        out, el = _cuda.trace_bundle(compiled, arrs[0], arrs[1], arrs[2], 1, 1000, 128, 0, 0, 0, return_elapsed=True)
        dt = time.perf_counter() - t0
    print(f"constant columns elided={elide}: wall {dt*1e3:.2f} ms  device-elapsed {el*1e3:.2f} ms  h2d {out['stats'][_cuda.STAT_H2D_BYTES]/1e6:.0f} MB", flush=True)
