#!/bin/bash
# Round GPU job: parity tests, bench lines, ncu launch lists and --set full captures (run under gpurun from the repo root).
TAG=${1:-r01c}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/gpu_tests_$TAG.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/gpu_tests_$TAG.log
python bench.py > gpurun_out/bench_$TAG.log 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
python bench.py --cohort 1 --lanes 1 --no-cpu-baseline --hf-baseline-users 0 > gpurun_out/bench_single_$TAG.log 2>/dev/null; echo "bench single rc=$?"
python bench.py --do-sample --dataset games --K 20 --constraint positional --no-cpu-baseline --hf-baseline-users 0 > gpurun_out/bench_relaxed_$TAG.log 2>/dev/null; echo "bench relaxed rc=$?"
# launch lists: single-user search (1 user after warm-up) and a cohort of 8 users
python tools/one_user.py > gpurun_out/plain_$TAG.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,launch__grid_size --clock-control none -s 1040 -c 1100 --csv --log-file gpurun_out/launches_single_$TAG.csv \
    python tools/one_user.py > gpurun_out/ncu_list_$TAG.log 2>&1; echo "ncu list single rc=$?"
python tools/one_user.py --cohort 8 --users 8 > gpurun_out/plain_c_$TAG.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,launch__grid_size --clock-control none -s 2600 -c 2700 --csv --log-file gpurun_out/launches_cohort_$TAG.csv \
    python tools/one_user.py --cohort 8 --users 8 > gpurun_out/ncu_list_c_$TAG.log 2>&1; echo "ncu list cohort rc=$?"
# full captures on the cohort path (second pass): GEMMs of 3 layers of a large forward, attention, row-wise, kernels (a)/(b)/(c)
ncu --set full --clock-control none --import-source on -k regex:gemm_wx -s 1330 -c 16 -f -o gpurun_out/prof_gemm_$TAG \
    python tools/one_user.py --cohort 8 --users 8 > gpurun_out/ncu_gemm_$TAG.log 2>&1; echo "ncu gemm rc=$?"
ncu --set full --clock-control none --import-source on -k "regex:tree_attention|mask_logsoftmax|kv_gather|cohort_verify|cohort_select|residual_rmsnorm|qkv_rope|silu_mul" \
    -s 2200 -c 40 -f -o gpurun_out/prof_small_$TAG python tools/one_user.py --cohort 8 --users 8 > gpurun_out/ncu_small_$TAG.log 2>&1; echo "ncu small rc=$?"
tail -c 2500 gpurun_out/bench_$TAG.log
