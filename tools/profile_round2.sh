#!/bin/bash
# round 2: bench line, ncu captures (full set + fp64 op counts) of the trace and intersect kernels, launch list
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; tail -c 600 gpurun_out/r2f_bench.json; tail -3 gpurun_out/r2f_bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wavefront -c 1 -o gpurun_out/r2f_wavefront -f python tools/profile_trace.py lsc_default 1e7 1 > gpurun_out/r2f_ncu_full.log 2>&1; tail -1 gpurun_out/r2f_ncu_full.log
timeout 600 ncu --clock-control none -k regex:wavefront -c 1 --csv --metrics smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__thread_inst_executed.sum python tools/profile_trace.py lsc_default 1e7 1 > gpurun_out/r2f_fp64_ops.csv 2>&1; tail -8 gpurun_out/r2f_fp64_ops.csv | cut -c1-40,200-
timeout 600 ncu --set full --clock-control none --import-source on -k regex:intersect_ring -c 1 -s 3 -o gpurun_out/r2f_intersect -f python tools/intersect_bench.py lsc_default 1e7 > gpurun_out/r2f_ncu_intersect.log 2>&1; tail -1 gpurun_out/r2f_ncu_intersect.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2f_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r2f_ncu_b.log 2>&1; tail -2 gpurun_out/r2f_launches.csv | cut -c1-200
PVTRACE_B200_LIB=$PWD/pvtrace_b200/csrc/lib_stages.so timeout 100 python tools/profile_trace.py lsc_default 1e7 2 | tail -1
