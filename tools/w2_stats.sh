#!/bin/bash
cd "$(dirname "$0")/.."
for v in "16 64" "16 96"; do
  set -- $v
  echo "wave2 W=$1 N=$2 (stats build):"; PVTRACE_B200_LIB=$PWD/pvtrace_b200/csrc/lib_w2stats.so PVT_KERNEL=wave2 PVT_WAVE2_WARPS=$1 PVT_WAVE2_SLOTS=$2 timeout 120 python tools/profile_trace.py lsc_default 1e7 2 | tail -2
done
