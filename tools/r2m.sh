#!/bin/bash
cd "$(dirname "$0")/.."
for cfg in lsc_default hello_world validation nested_cylinders lsc_coated; do bash tools/ab_variants.sh "$cfg 1e7" prev cur | tail -2; done
timeout 600 python -m pytest tests/test_gpu_trace_parity.py tests/test_gpu_engine.py -q -x 2>&1 | tail -2
