import time, numpy as np, sys
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import pvtrace_b200 as pv
from pvtrace_b200.device import configs
from pvtrace_b200.engine import _cuda
from pvtrace_b200.engine.compiler import EMIT_METHODS
print('devices', _cuda.device_count())
for name,(fn,kw) in configs.CONFIGS.items():
    s=fn(); c=pv.engine.compile_scene(s); e=pv.engine.compile_emitter(s)
    m=EMIT_METHODS[kw['emit_method']]
    for n in (1000, 1000000, 10000000):
        for rep in range(2):
            d,el=_cuda.trace_bundle(c,None,None,None,1,1000,128,m,0,0,emitter=e,n=n,return_elapsed=True)
        print(name, n, 'elapsed', el, 'Mphot/s', n/el/1e6, 'steps/photon', d['stats'][0]/n, flush=True)
    print('  ', dict(zip(c.recorder_names, d['rec_distinct'].tolist())))
