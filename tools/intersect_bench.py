"""Time intersect_kernel alone on device-resident rays: GB/s of its 68 B/ray against the measured HBM peak."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pvtrace_b200 as pv
from pvtrace_b200.device import configs
from pvtrace_b200.engine import _cuda
name = sys.argv[1] if len(sys.argv) > 1 else "lsc_default"
n = int(float(sys.argv[2])) if len(sys.argv) > 2 else 20_000_000
scene = configs.CONFIGS[name][0]()
ctx = _cuda.Context(pv.engine.compile_scene(scene), pv.engine.compile_emitter(scene), 0)
pos = torch.empty((n, 3), dtype=torch.float64, device="cuda"); d = torch.empty_like(pos)
wl = torch.empty(n, dtype=torch.float64, device="cuda")
ctx.emit(pos.data_ptr(), d.data_ptr(), wl.data_ptr(), n, seed=1)
pos += torch.randn_like(pos) * 2.0  # rays all over the scene, not just at the lamp
t0 = torch.empty(n, dtype=torch.float64, device="cuda"); ids = torch.empty((3, n), dtype=torch.int32, device="cuda")
def run():
    ctx.intersect(pos.data_ptr(), d.data_ptr(), n, t0.data_ptr(), ids[0].data_ptr(), ids[1].data_ptr(), ids[2].data_ptr())
for _ in range(3): run()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10): run()
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 10
peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
print(f"{name} n={n} ctas={os.environ.get('PVT_INTERSECT_CTAS', '3')}: {ms:.4f} ms  {68 * n / ms / 1e6:.0f} GB/s  frac {68 * n / ms / 1e6 / peak:.3f}")
