"""Host-ray end to end against the streaming upload's chunk schedule (PVT_UPLOAD_MIN_CHUNK / PVT_UPLOAD_SHRINK)."""
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
for mc, sh in ((65536, 5), (131072, 5), (131072, 4), (262144, 4), (262144, 3), (524288, 3), (32768, 6)):
    os.environ["PVT_UPLOAD_MIN_CHUNK"] = str(mc); os.environ["PVT_UPLOAD_SHRINK"] = str(sh)
    best = 1e9
    for rep in range(4):
        out, el = _cuda.trace_bundle(compiled, arrs[0], arrs[1], arrs[2], 1, 1000, 128, 0, 0, 0, return_elapsed=True)
        best = min(best, el)
    print(f"min_chunk={mc} shrink=1/{sh}: device-elapsed best {best*1e3:.2f} ms", flush=True)
