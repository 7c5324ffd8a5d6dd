"""Time the intersect stage alone on device-resident rays against the measured HBM peak: the packed form (60 B per ray,
SURVEY 8d) and the three-array form of the host call (68 B per ray).  PVT_INTERSECT_VARIANT picks the ring shape."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pvtrace_b200 as pv
from pvtrace_b200.device import configs
from pvtrace_b200.engine import _cuda
name = sys.argv[1] if len(sys.argv) > 1 else "lsc_default"
n = int(float(sys.argv[2])) if len(sys.argv) > 2 else 20_000_000
spread = float(sys.argv[3]) if len(sys.argv) > 3 else 2.0
scene = configs.CONFIGS[name][0]()
ctx = _cuda.Context(pv.engine.compile_scene(scene), pv.engine.compile_emitter(scene), 0)
pos = torch.empty((n, 3), dtype=torch.float64, device="cuda"); d = torch.empty_like(pos)
wl = torch.empty(n, dtype=torch.float64, device="cuda")
ctx.emit(pos.data_ptr(), d.data_ptr(), wl.data_ptr(), n, seed=1)
if spread > 0:
    pos += torch.randn_like(pos) * spread  # rays all over the scene, not just at the lamp
t0 = torch.empty(n, dtype=torch.float64, device="cuda"); ids = torch.empty((3, n), dtype=torch.int32, device="cuda")
t0p = torch.empty(n, dtype=torch.float64, device="cuda"); packed = torch.empty(n, dtype=torch.int32, device="cuda")
def run3():
    ctx.intersect(pos.data_ptr(), d.data_ptr(), n, t0.data_ptr(), ids[0].data_ptr(), ids[1].data_ptr(), ids[2].data_ptr())
def runp():
    ctx.intersect_packed(pos.data_ptr(), d.data_ptr(), n, t0p.data_ptr(), packed.data_ptr())
peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device="cuda")  # 256 MB > L2
for label, fn, bytes_per_ray in (("packed", runp, 60), ("3 arrays", run3, 68)):
    for _ in range(3): fn()
    # "dirty": 256 MB written before the kernel (its dirty lines are written back while the kernel runs);
    # "clean": written, then read back -- a cold L2 with nothing to write back, like ncu's cache control
    for kind in ("clean", "dirty"):
        times = []
        for _ in range(10):
            flush.zero_()
            if kind == "clean": flush.sum()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record(); torch.cuda.synchronize()
            times.append(a.elapsed_time(b))
        ms = sum(times) / len(times)
        print(f"{name} n={n} variant={os.environ.get('PVT_INTERSECT_VARIANT', '0')} {label} [{kind} L2]: {ms:.4f} ms (min {min(times):.4f})  "
              f"{bytes_per_ray * n / ms / 1e6:.0f} GB/s  frac {bytes_per_ray * n / ms / 1e6 / peak:.3f}", flush=True)
p = packed.to(torch.int64) & 0xffffffff
unpack = lambda sh: torch.where(((p >> sh) & 0xff) == 0xff, torch.full_like(p, -1), (p >> sh) & 0xff).to(torch.int32)
ok = bool((unpack(0) == ids[0]).all() and (unpack(8) == ids[1]).all() and (unpack(16) == ids[2]).all() and (t0 == t0p).all())
print("packed == 3 arrays:", ok)
