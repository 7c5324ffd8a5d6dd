#!/bin/bash
# time the wavefront kernel variants on one config:  tools/sweep.sh lsc_default 1e7
cfg=${1:-lsc_default}; n=${2:-1e7}
for v in "1024 1" "768 1" "640 1" "512 2" "512 1" "384 2"; do
  set -- $v
  echo -n "T=$1 B=$2: "
  PVT_WAVEFRONT_THREADS=$1 PVT_WAVEFRONT_CTAS=$2 python tools/profile_trace.py $cfg $n 2 | tail -1
done
