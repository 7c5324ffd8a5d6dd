#!/bin/bash
cd "$(dirname "$0")/.."
timeout 100 python tools/profile_trace.py lsc_default 1e7 3 | tail -1
for g in 125 110 150 200; do
  echo "growth $g:"; PVT_UPLOAD_GROWTH_PCT=$g PVT_DEBUG_TIMING=1 timeout 200 python tools/e2e_upload_timing.py 2>&1 | grep -E "elided=1|mask 5" | tail -2
done
echo "growth 125 min chunk 32768:"; PVT_UPLOAD_MIN_CHUNK=32768 PVT_DEBUG_TIMING=1 timeout 200 python tools/e2e_upload_timing.py 2>&1 | grep -E "elided=1|mask 5" | tail -2
