import sys
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import numpy as np
import pvtrace_b200 as pv
from oracle import pvt_oracle
from pvtrace_b200.device import configs
from pvtrace_b200.engine import _cuda
from pvtrace_b200.engine.compiler import EMIT_METHODS
name = sys.argv[1] if len(sys.argv) > 1 else 'nested_cylinders'
n, m = 20000, 64
build, kw = configs.CONFIGS[name]
scene = build()
c, e = pv.engine.compile_scene(scene), pv.engine.compile_emitter(scene)
method = EMIT_METHODS[kw["emit_method"]]
got = _cuda.trace_bundle(c, None, None, None, 7, 1000, m, method, 0, 1, emitter=e, n=n)
want = pvt_oracle.trace_bundle(c, None, None, None, 7, 1000, m, method, 8, 1, emitter=e, n=n)
kg, kw_ = got["kind"].reshape(n, m), want["kind"].reshape(n, m)
bad = np.where((got["counts"] != want["counts"]) | (kg != kw_).any(axis=1))[0]
print("mismatching rays", len(bad), "of", n)
np.set_printoptions(precision=17, linewidth=200)
for i in bad[:6]:
    print("=== ray", i, "counts", got["counts"][i], want["counts"][i])
    for k in range(max(got["counts"][i], want["counts"][i])):
        r = i * m + k
        for tag, d in (("gpu", got), ("cpu", want)):
            print(tag, k, "kind", d["kind"][r], "hit", d["hit"][r], "cont", d["container"][r], "adj", d["adjacent"][r],
                  "pos", d["position"][r], "dir", d["direction"][r], "nrm", d["normal"][r])
        if got["kind"][r] != want["kind"][r] or got["hit"][r] != want["hit"][r]:
            break
