#!/bin/bash
cd "$(dirname "$0")/.."
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
PVT_INTERSECT_VARIANT=0 timeout 120 python tools/intersect_bench.py lsc_default 1e7 2>&1 | tail -3
PVT_INTERSECT_VARIANT=5 timeout 120 python tools/intersect_bench.py lsc_default 1e7 2>&1 | tail -3
timeout 100 python tools/profile_trace.py lsc_default 1e7 3 | tail -1
PVT_DEBUG_TIMING=1 timeout 200 python tools/e2e_upload_timing.py 2>&1 | grep -E "elided=1|mask 5" | tail -2
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -4
