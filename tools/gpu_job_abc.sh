#!/bin/bash
# Kernels (a) and (c) stand-alone: CUDA-event GB/s at SURVEY 8(d) sizes, then one `ncu --set full` capture of the same
# kernels (only after the plain run exited 0), then the launch list of a cohort search restricted to kernels (a)/(b)/(c).
TAG=${1:-r02}
mkdir -p gpurun_out
python tools/kernel_abc_bench.py --json gpurun_out/abc_bench_$TAG.json > gpurun_out/abc_bench_$TAG.txt 2>&1; rc=$?; echo "abc bench rc=$rc"; cat gpurun_out/abc_bench_$TAG.txt
[ $rc -eq 0 ] && ncu --set full --clock-control none --import-source on -k "regex:mask_logsoftmax|kv_gather" -c 16 -f \
    -o gpurun_out/prof_abc_$TAG python tools/kernel_abc_bench.py --iters 1 --quick > gpurun_out/ncu_abc_$TAG.log 2>&1; echo "ncu abc rc=$?"
python tools/one_user.py --cohort 8 --users 8 > gpurun_out/plain_c_$TAG.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k "regex:mask_logsoftmax|kv_gather|cohort_verify|cohort_select|tree_verify|tree_select" \
    -c 24 -f -o gpurun_out/prof_abc_path_$TAG python tools/one_user.py --cohort 8 --users 8 > gpurun_out/ncu_abc_path_$TAG.log 2>&1; echo "ncu abc-in-path rc=$?"
