#!/bin/bash
cd "$(dirname "$0")/.."
bash tools/ab_variants.sh "lsc_default 1e7" prev cur
bash tools/ab_variants.sh "hello_world 1e7" prev cur | tail -2
timeout 600 python -m pytest tests/test_gpu_host_path.py tests/test_gpu_engine.py tests/test_gpu_trace_parity.py -q -x 2>&1 | tail -3
PVT_DEBUG_TIMING=1 timeout 200 python tools/e2e_upload_timing.py 2>&1 | grep -E "elided=|mask" | tail -4
