"""Small runs of every kernel family through the C ABI (host arrays in, host arrays out; no torch), the target of
tools/sanitize.sh:  compute-sanitizer --tool {memcheck,racecheck,synccheck} python tools/sanitize_run.py [n]

Covers: wavefront_kernel with service warps (box scene, recorders) and without (no recorders), with and without the
event log, host rays (streaming upload, constant columns) and on-device emission; trace_kernel (Philox and the
reference's xoshiro stream); the non-box wavefront instantiation (spheres, cylinders); intersect ring kernel; emit_kernel.  Prints one line per case with the bookkeeping invariant exit + lost == n."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import pvtrace_b200 as pv
from pvtrace_b200.device import configs
from pvtrace_b200.engine import _cuda, emit
from pvtrace_b200.engine.compiler import EMIT_METHODS

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 20000
lib = _cuda.load_library()


def ended(compiled, data):
    names = list(compiled.recorder_names)
    return int(sum(int(data["rec_distinct"][k]) for k, nm in enumerate(names) if nm == "exit" or nm.endswith("lost")))


for name in ("lsc_default", "lsc_coated", "hello_world", "nested_cylinders", "validation"):
    build, kw = configs.CONFIGS[name]
    scene = build()
    compiled = pv.engine.compile_scene(scene)
    emitter = pv.engine.compile_emitter(scene)
    method = EMIT_METHODS[kw["emit_method"]]
    pos, dirs, wl, _ = emit.emit_bundle(scene, n, seed=3)  # emit_kernel
    cases = [("host rays", dict(), (pos, dirs, wl)),
             ("host rays, varying columns", dict(), (pos + 1e-9 * np.arange(n)[:, None], dirs, wl + 1e-9 * np.arange(n))),
             ("device emission", dict(emitter=emitter, n=n), (None, None, None)),
             ("event log", dict(record_every=7), (pos, dirs, wl)),
             ("register kernel", dict(flags=_cuda.FLAG_REGISTER_KERNEL, record_every=11), (pos, dirs, wl)),
             ("xoshiro", dict(rng_mode=_cuda.RNG_XOSHIRO), (pos, dirs, wl))]
    for label, extra, rays in cases:
        record_every = extra.pop("record_every", 0)
        data = _cuda.trace_bundle(compiled, rays[0], rays[1], rays[2], 5, 1000, 64, method, 0, record_every, **extra)
        print(f"{name:17s} {label:28s} steps {int(data['stats'][_cuda.STAT_STEPS]):8d}  ended {ended(compiled, data)} of {n}", flush=True)
    # the intersect stage (TMA ring; the plain kernel for unaligned device arrays is covered by tests/test_gpu_edge_cases.py)
    scene_struct, keep = _cuda.marshal_scene(compiled)
    t0 = np.zeros(n)
    hit, cont, adj = (np.zeros(n, dtype=np.int32) for _ in range(3))
    elapsed = C.c_double()
    _cuda.check(lib.pvt_intersect_bundle(C.byref(scene_struct), _cuda._vp(pos), _cuda._vp(dirs), n, _cuda._vp(t0), _cuda._vp(hit),
                                         _cuda._vp(cont), _cuda._vp(adj), 0, C.byref(elapsed)), "intersect_bundle")
    print(f"{name:17s} intersect stage              hits {int((hit >= 0).sum())} of {n}", flush=True)
print("done")
