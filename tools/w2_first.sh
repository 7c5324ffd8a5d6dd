#!/bin/bash
# first light of warp_wavefront_kernel: parity (ray by ray vs the oracle) + timing against the CTA wavefront
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_gpu_trace_parity.py -x -q -k "warp or cta" 2>&1 | tail -15
for k in wave1 wave2; do
  echo -n "$k: "; PVT_KERNEL=$k timeout 120 python tools/profile_trace.py lsc_default 1e7 3 | tail -1
done
for v in "16 64" "20 64" "16 96" "12 64"; do
  set -- $v
  echo -n "wave2 W=$1 N=$2: "; PVT_KERNEL=wave2 PVT_WAVE2_WARPS=$1 PVT_WAVE2_SLOTS=$2 timeout 120 python tools/profile_trace.py lsc_default 1e7 3 | tail -1
done
