#!/bin/bash
# Build A/B variants of the library side by side:  tools/build_variants.sh name1="-DFLAG1" name2="-DFLAG1 -DFLAG2" ...
# -> pvtrace_b200/csrc/lib_<name>.so (git-ignored; they travel to the GPU box).  Compare with tools/ab_variants.sh
# WITHIN ONE gpurun call: leases differ by up to 6 % on the trace kernel.
cd "$(dirname "$0")/.." || exit 1
for spec in "$@"; do
  name=${spec%%=*}; flags=${spec#*=}
  ( python -m pvtrace_b200.csrc.build --output lib_$name.so -- $flags | tail -1; echo "built lib_$name.so ($flags)" ) &
done
wait
