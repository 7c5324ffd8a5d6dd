#!/bin/bash
# Build A/B variants of the library side by side:  tools/build_variants.sh name1="-DFLAG1" name2="-DFLAG1 -DFLAG2" ...
# -> pvtrace_b200/csrc/lib_<name>.so (git-ignored; they travel to the GPU box).  Compare with tools/ab_variants.sh.
cd "$(dirname "$0")/../pvtrace_b200/csrc" || exit 1
for spec in "$@"; do
  name=${spec%%=*}; flags=${spec#*=}
  ( nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -shared -Xcompiler -fPIC -Xcompiler -O2 $flags \
      -o lib_$name.so pvt_api.cu 2>&1 | grep -E "error|warning" ; echo "built lib_$name.so ($flags)" ) &
done
wait
