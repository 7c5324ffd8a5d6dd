#!/bin/bash
# tools/multi_gpu_check.sh N -- on an N-GPU box: topology, the multi-device / multi-rank invariance tests, the scaling
# bench at N ranks, and several GPUs driven from ONE process (engine.simulate(workers=N))
cd "$(dirname "$0")/.."
N=${1:-2}
nvidia-smi topo -m 2>&1 | head -14
python - <<PY
import os
print("cpus", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
try:
    nodes = sorted(d for d in os.listdir("/sys/devices/system/node") if d.startswith("node"))
    print("numa nodes", nodes, [open(f"/sys/devices/system/node/{d}/cpulist").read().strip() for d in nodes])
except Exception as exc:
    print("numa:", exc)
PY
timeout 900 python -m pytest tests/test_gpu_host_path.py -q 2>&1 | tail -5
for n in 1 $N; do
  if [ $n -eq 1 ]; then
    timeout 600 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline --no-extra 2>&1 | tail -1 > gpurun_out/mg_bench_n1.json
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $n --steps 5 --warmup 3 --no-cpu-baseline --no-extra 2>&1 | tail -1 > gpurun_out/mg_bench_n$n.json
  fi
  python - <<PY
import json
d = json.loads(open("gpurun_out/mg_bench_n$n.json").read())
print("N=$n value %.3e ms %.3f | e2e %.3e ms %.3f h2d %d | emission e2e ms %.3f | %s | %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["h2d_bytes_per_step"], d["e2e_device_emission"]["ms_per_step"], d["config"]["tallies_checked"], d["config"]["cpu_binding"]))
PY
done
timeout 600 python - <<PY
import time, numpy as np
import pvtrace_b200 as pv
from pvtrace_b200.device import configs
scene = configs.lsc_default()
n = 10_000_000
for workers in (1, $N):
    total = n * workers
    for rep in range(3):
        tic = time.perf_counter()
        r = pv.engine.simulate(scene, total, seed=3, record_every=0, workers=workers)
        dt = time.perf_counter() - tic
    print(f"one process, workers={workers}: {total} photons in {dt*1e3:.2f} ms wall ({total/dt/1e9:.3f} G photons/s), device time {r.elapsed*1e3:.2f} ms, exit+lost {r.recorders['exit'].rays + r.recorders['LSC-lost'].rays}")
PY
if [ $N -ge 8 ]; then
  # BASELINE config 4: LSC + edge solar cells + back mirror, 10^8 photons over 8 B200 -- as ranks and from one process
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu-baseline --no-extra --config lsc_coated --photons 1.25e7 2>&1 | tail -1 > gpurun_out/mg_bench_coated_n8.json
  python - <<PY
import json
d = json.loads(open("gpurun_out/mg_bench_coated_n8.json").read())
print("config 4, 8 ranks x 1.25e7: value %.3e ms %.3f | e2e %.3e ms %.3f | %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["config"]["tallies_checked"]))
PY
  timeout 600 python - <<PY
import time
import pvtrace_b200 as pv
from pvtrace_b200.device import configs
for name in ("lsc_default", "lsc_coated"):
    scene = configs.CONFIGS[name][0]()
    one = pv.engine.simulate(scene, 10**7, seed=5, record_every=0, workers=1)
    for rep in range(2):
        tic = time.perf_counter()
        r = pv.engine.simulate(scene, 10**8, seed=5, record_every=0, workers=8)
        dt = time.perf_counter() - tic
    rec = r.recorders
    print(f"{name}: engine.simulate(scene, 10**8, workers=8) from one python process: {dt*1e3:.2f} ms wall ({1e8/dt/1e9:.2f} G photons/s), device {r.elapsed*1e3:.2f} ms; exit+lost {rec['exit'].rays + rec['LSC-lost'].rays}; bottom {rec['LSC-bottom'].rays}")
    again = pv.engine.simulate(scene, 10**8, seed=5, record_every=0, devices=[0, 1, 2, 3])
    same = all((again.data[k] == r.data[k]).all() for k in ("rec_distinct", "rec_crossings", "rec_bins"))
    print(f"   4 devices give the tallies of 8: {same}")
PY
fi
