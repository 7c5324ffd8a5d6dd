#!/bin/bash
cd "$(dirname "$0")/.."
timeout 200 python -m pytest tests/test_gpu_trace_parity.py -x -q -k "default" 2>&1 | tail -3
PVTRACE_B200_LIB=$PWD/pvtrace_b200/csrc/lib_svc64.so timeout 200 python -m pytest tests/test_gpu_trace_parity.py -x -q -k "default and lsc_default" 2>&1 | tail -3
for pass in 1 2; do for name in cur svc64; do
    lib=$PWD/pvtrace_b200/csrc/lib_$name.so; [ "$name" = cur ] && lib=$PWD/pvtrace_b200/csrc/libpvtrace_b200.so
    echo -n "$name: "; PVTRACE_B200_LIB=$lib timeout 60 python tools/profile_trace.py lsc_default 1e7 3 | tail -1; echo
done; done
