#!/bin/bash
cd "$(dirname "$0")/.."
bash tools/ab_variants.sh "lsc_default 1e7" cur latec latev
for cfg in hello_world nested_cylinders; do timeout 100 python tools/profile_trace.py $cfg 1e7 3 | tail -1; done
