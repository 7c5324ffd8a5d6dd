"""Does the streaming-upload code path itself cost time?  A bundle whose three columns are constant needs no upload: the
host call then runs the kernel in 'arrival mark' mode with the mark already at n, and can be compared with the same
bundle traced from device-resident arrays."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pvtrace_b200 as pv
from pvtrace_b200.device import configs
from pvtrace_b200.engine import _cuda
n = 10_000_000
scene = configs.lsc_default()
compiled, emitter = pv.engine.compile_scene(scene), pv.engine.compile_emitter(scene)
pos = np.tile(np.array([0.0, 0.0, 5.0]), (n, 1)); d = np.tile(np.array([0.05, 0.02, -1.0]) / np.linalg.norm([0.05, 0.02, -1.0]), (n, 1))
wl = np.full(n, 555.0)
h = [torch.from_numpy(a).pin_memory() for a in (pos, d, wl)]
arrs = [t.numpy() for t in h]
for rep in range(4):
    out, el = _cuda.trace_bundle(compiled, arrs[0], arrs[1], arrs[2], 1, 1000, 128, 0, 0, 0, return_elapsed=True)
print(f"host call, all columns constant (no upload): device-elapsed {el*1e3:.3f} ms  h2d {out['stats'][_cuda.STAT_H2D_BYTES]} steps {out['stats'][0]}")
ctx = _cuda.Context(compiled, emitter, 0)
dev = [t.cuda() for t in h]
for rep in range(4):
    ctx.reset()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); ctx.trace(n, 1, d_positions=dev[0].data_ptr(), d_directions=dev[1].data_ptr(), d_wavelengths=dev[2].data_ptr()); b.record()
    torch.cuda.synchronize()
print(f"device-resident arrays: {a.elapsed_time(b):.3f} ms steps {ctx.read()['stats'][0]}")
