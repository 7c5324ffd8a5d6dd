"""Where does the host-ray end-to-end time go?  H2D bandwidth of the pinned arrays and pvt_trace_bundle vs chunk count."""
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
for rep in range(3):
    t0 = time.perf_counter()
    for a, b in zip(h, d): b.copy_(a, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"H2D 560 MB pinned: {dt*1e3:.2f} ms  {0.56/dt:.1f} GB/s")
arrs = [t.numpy() for t in h]
for mode in ("stream", "copy"):
    os.environ["PVT_STREAM_UPLOAD"] = "1" if mode == "stream" else "0"
    for chunks in ((16, 32, 64) if mode == 'stream' else (4,)):
        os.environ["PVT_UPLOAD_CHUNKS"] = str(chunks)
        for rep in range(3):
            t0 = time.perf_counter()
            out, el = _cuda.trace_bundle(compiled, arrs[0], arrs[1], arrs[2], 1, 1000, 128, 0, 0, 0, return_elapsed=True)
            dt = time.perf_counter() - t0
        print(f"{mode} chunks={chunks:2d}: wall {dt*1e3:.2f} ms  device-elapsed {el*1e3:.2f} ms  exit {out['rec_distinct'][0]} rays {out['stats'][1]}")
os.environ.pop("PVT_UPLOAD_CHUNKS"); os.environ["PVT_STREAM_UPLOAD"] = "1"
pageable = [np.array(a) for a in arrs]
for n_small in (1000, 12345, 2_000_003):
    out = _cuda.trace_bundle(compiled, pageable[0][:n_small], pageable[1][:n_small], pageable[2][:n_small], 1, 1000, 128, 0, 0, 0)
    ref = _cuda.trace_bundle(compiled, pageable[0][:n_small], pageable[1][:n_small], pageable[2][:n_small], 1, 1000, 128, 0, 0, 0, flags=1)
    print(n_small, "stream == register kernel:", (out["rec_distinct"] == ref["rec_distinct"]).all(), out["stats"][1])
for rep in range(3):
    t0 = time.perf_counter()
    out, el = _cuda.trace_bundle(compiled, pageable[0], pageable[1], pageable[2], 1, 1000, 128, 0, 0, 0, return_elapsed=True)
    dt = time.perf_counter() - t0
print(f"pageable numpy arrays: wall {dt*1e3:.2f} ms  exit {out['rec_distinct'][0]}")
