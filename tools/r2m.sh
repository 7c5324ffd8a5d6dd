#!/bin/bash
cd "$(dirname "$0")/.."
bash tools/ab_variants.sh "lsc_default 1e7" r2h cur lean
PVT_INTERSECT_VARIANT=0 timeout 120 python tools/intersect_bench.py lsc_default 1e7 2>&1 | tail -3
