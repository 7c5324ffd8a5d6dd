#!/bin/bash
cd "$(dirname "$0")/.."
timeout 300 python -m pytest tests/test_gpu_engine.py -x -q -k "intersect" 2>&1 | tail -3
for v in 0 1 2 3 4 5 -1; do
  PVT_INTERSECT_VARIANT=$v timeout 120 python tools/intersect_bench.py lsc_default 1e7 2>&1 | tail -3
done
PVT_INTERSECT_VARIANT=0 timeout 120 python tools/intersect_bench.py lsc_default 2e7 2>&1 | tail -3
PVT_INTERSECT_VARIANT=0 timeout 120 python tools/intersect_bench.py nested_cylinders 1e7 0.5 2>&1 | tail -3
for t in 2 4 8 12 16; do
  echo "scan threads $t:"; PVT_SCAN_THREADS=$t PVT_DEBUG_TIMING=1 timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | python -c "
import sys, json
for line in sys.stdin:
    if line.startswith('{'):
        d = json.loads(line); print('  e2e ms', d['e2e']['ms_per_step'], 'kernel ms', d['roofline']['kernel_ms'])
"
done
