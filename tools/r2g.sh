#!/bin/bash
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_gpu_reference_parity.py tests/test_gpu_coatings.py -q -s 2>&1 | grep -E "max .z.|passed|failed|Error" | tail -25
timeout 900 python -m pytest tests -q -m gpu --deselect tests/test_gpu_reference_parity.py --deselect tests/test_gpu_coatings.py 2>&1 | tail -12
PVT_INTERSECT_VARIANT=0 timeout 120 python tools/intersect_bench.py lsc_default 1e7 2>&1 | tail -3
