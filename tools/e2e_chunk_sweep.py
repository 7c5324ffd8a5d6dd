"""Host-ray end to end against the streaming upload's chunk schedule (PVT_UPLOAD_MIN_CHUNK / _SHRINK / _GROWTH_PCT)."""
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
grid = [(65536, 5, g) for g in (105, 110, 115, 125, 150)] + [(32768, 5, 110), (131072, 5, 110), (65536, 8, 110), (65536, 4, 110)]
if len(sys.argv) > 1 and sys.argv[1] == "chunks":
    grid = [(mc, sh, 125) for mc, sh in ((65536, 5), (131072, 5), (131072, 4), (262144, 4), (262144, 3), (524288, 3), (32768, 6))]
for pass_ in range(2):  # interleaved: two passes over the grid
    for mc, sh, g in grid:
        os.environ["PVT_UPLOAD_MIN_CHUNK"] = str(mc); os.environ["PVT_UPLOAD_SHRINK"] = str(sh)
        os.environ["PVT_UPLOAD_GROWTH_PCT"] = str(g)
        times, walls = [], []
        for rep in range(5):
            tic = time.perf_counter()
            out, el = _cuda.trace_bundle(compiled, arrs[0], arrs[1], arrs[2], 1, 1000, 128, 0, 0, 0, return_elapsed=True)
            walls.append(time.perf_counter() - tic); times.append(el)
        print(f"min_chunk={mc} shrink=1/{sh} growth={g}%: device-elapsed best {min(times)*1e3:.3f} median {sorted(times)[2]*1e3:.3f} ms, "
              f"wall median {sorted(walls)[2]*1e3:.3f} ms", flush=True)
