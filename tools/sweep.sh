#!/bin/bash
# time wavefront kernel variants on one config:  tools/sweep.sh lsc_default 1e7 "512 1024 1" "640 1152 1" ...
cfg=${1:-lsc_default}; n=${2:-1e7}; shift; shift
if [ $# -eq 0 ]; then set -- "512 1024 1" "512 1280 1" "640 1280 1" "576 1152 1" "448 1344 1" "640 1152 1"; fi
for v in "$@"; do
  set -- $v
  echo -n "T=$1 P=$2 B=$3: "
  PVT_WAVEFRONT_THREADS=$1 PVT_WAVEFRONT_POOL=$2 PVT_WAVEFRONT_CTAS=$3 python tools/profile_trace.py $cfg $n 2 | tail -1
done
