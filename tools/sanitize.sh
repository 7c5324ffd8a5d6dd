#!/bin/bash
# tools/sanitize.sh [n] -- compute-sanitizer over small runs of every kernel family (race detection).
# memcheck: out-of-bounds / misaligned accesses; racecheck: shared-memory hazards of the wavefront kernel's pool, queues
# and counters (two flag-synchronised pairs are expected, see profiles/r2_sanitizer.txt); synccheck: barrier misuse (named barriers, mbarrier, setmaxnreg regions).  Output: gpurun_out/sanitize_*.log
cd "$(dirname "$0")/.."
n=${1:-20000}
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  extra=""
  [ $tool = racecheck ] && extra="--racecheck-report analysis"
  timeout 900 compute-sanitizer --tool $tool $extra --print-limit 400 python tools/sanitize_run.py $n > gpurun_out/sanitize_$tool.log 2>&1
  echo "== $tool rc=$? : $(grep -c -E '^[a-z_]+ +[a-z]' gpurun_out/sanitize_$tool.log) case lines, $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitize_$tool.log | tail -1)"
done
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_edge_cases.py -q -x > gpurun_out/sanitize_memcheck_edge_cases.log 2>&1
echo "== memcheck over tests/test_gpu_edge_cases.py rc=$? : $(grep -E 'passed|failed' gpurun_out/sanitize_memcheck_edge_cases.log | tail -1) $(grep 'ERROR SUMMARY' gpurun_out/sanitize_memcheck_edge_cases.log | tail -1)"
