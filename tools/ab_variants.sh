#!/bin/bash
# tools/ab_variants.sh "cfg n" name1 name2 ...   (interleaved, two passes; env PVT_WAVEFRONT_* passes through)
set -- $@
cfg=$1; n=$2; shift; shift
for pass in 1 2; do
  for name in "$@"; do
    lib=$PWD/pvtrace_b200/csrc/lib_$name.so; [ "$name" = cur ] && lib=$PWD/pvtrace_b200/csrc/libpvtrace_b200.so
    echo -n "$name: "; PVTRACE_B200_LIB=$lib python tools/profile_trace.py $cfg $n 3 | tail -1
  done
done
