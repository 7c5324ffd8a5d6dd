#!/bin/bash
# A/B two builds of the library on the GPU box in ONE call (boxes differ by a few %):
#   git stash; python -m pvtrace_b200.csrc.build --force; cp pvtrace_b200/csrc/libpvtrace_b200.so pvtrace_b200/csrc/lib_before.so; git stash pop; rebuild
#   gpurun -- tools/ab.sh lsc_default 1e7
cfg=${1:-lsc_default}; n=${2:-1e7}
for i in 1 2; do
  echo -n "before: "; PVTRACE_B200_LIB=$PWD/pvtrace_b200/csrc/lib_before.so python tools/profile_trace.py $cfg $n 3 | tail -1
  echo -n "after:  "; python tools/profile_trace.py $cfg $n 3 | tail -1
done
