#!/bin/bash
cd "$(dirname "$0")/.."
for pass in 1 2; do
for name in cur fixedkey; do
  lib=$PWD/pvtrace_b200/csrc/lib_$name.so; [ "$name" = cur ] && lib=$PWD/pvtrace_b200/csrc/libpvtrace_b200.so
  echo -n "$name: "; PVTRACE_B200_LIB=$lib python tools/profile_trace.py lsc_default 1e7 3 | tail -1
done; done
echo -n "stages: "; PVTRACE_B200_LIB=$PWD/pvtrace_b200/csrc/lib_stages.so python tools/profile_trace.py lsc_default 1e7 2 | tail -1
