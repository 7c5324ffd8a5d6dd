#!/bin/bash
# what the driver runs at round end, in one call: smoke, the gpu test suite, the reference arm and the bench line
cd "$(dirname "$0")/.."
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -3
timeout 900 python bench.py --impl reference --gpus 1 --steps 5 --warmup 2 2>&1 | tail -1 | cut -c1-400
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 2>&1 | tail -1 > gpurun_out/final_bench.json
python - <<PY
import json
d = json.loads(open("gpurun_out/final_bench.json").read())
print("value %.4g (%.3f ms) e2e %.4g (%.3f ms, h2d %d) roofline %.3f fp64 %.3f intersect %.3f cpu %.4g launches %d clocks %s" % (
    d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["h2d_bytes_per_step"], d["roofline"]["frac"],
    d["roofline_fp64"]["frac"], d["intersect_stage"]["frac_of_hbm_peak"], d["cpu_baseline"]["value"], d["gpu_launches"], d["clocks"]))
PY
