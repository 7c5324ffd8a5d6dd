#!/bin/bash
cd "$(dirname "$0")/.."
for pass in 1 2; do
echo -n "service warps:  "; timeout 100 python tools/profile_trace.py lsc_default 1e7 3 | tail -1
echo -n "tally in place: "; PVT_TALLY_IN_PLACE=1 timeout 100 python tools/profile_trace.py lsc_default 1e7 3 | tail -1
done
for cfg in validation nested_cylinders hello_world lsc_coated; do timeout 100 python tools/profile_trace.py $cfg 1e7 3 | tail -1; done
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -3
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2o_bench.json 2> gpurun_out/r2o_bench.err; python -c "
import json; d=json.loads(open('gpurun_out/r2o_bench.json').read().strip().splitlines()[-1])
print('value %.4g ms %.3f | e2e %.4g ms %.3f | frac %.3f | fp64 %.3f | intersect %.3f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['roofline']['frac'], d['roofline_fp64'].get('frac', -1), d['intersect_stage']['frac_of_hbm_peak']))"
