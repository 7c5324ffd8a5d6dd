#!/bin/bash
cd "$(dirname "$0")/.."
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wavefront -c 1 -o gpurun_out/r2n_wavefront -f python tools/profile_trace.py lsc_default 1e7 1 > gpurun_out/r2n_ncu_full.log 2>&1; tail -2 gpurun_out/r2n_ncu_full.log
timeout 100 python tools/profile_trace.py lsc_default 1e7 3 | tail -1
