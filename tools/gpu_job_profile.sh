#!/bin/bash
# Round GPU job: parity tests, bench line, ncu launch list and --set full captures (run under gpurun from the repo root).
TAG=${1:-r01b}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests_$TAG.log 2>&1; echo "pytest rc=$?"
python bench.py > gpurun_out/bench_$TAG.log 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
python tools/one_user.py > gpurun_out/plain_$TAG.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 1040 -c 1100 --csv --log-file gpurun_out/launches_$TAG.csv \
    python tools/one_user.py > gpurun_out/ncu_list_$TAG.log 2>&1; echo "ncu list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:gemm_wx -s 468 -c 12 -f -o gpurun_out/prof_gemm_$TAG \
    python tools/one_user.py > gpurun_out/ncu_gemm_$TAG.log 2>&1; echo "ncu gemm rc=$?"
ncu --set full --clock-control none --import-source on -k regex:gemm_wx -s 596 -c 1 -f -o gpurun_out/prof_lmhead_$TAG \
    python tools/one_user.py > gpurun_out/ncu_lmhead_$TAG.log 2>&1; echo "ncu lmhead rc=$?"
ncu --set full --clock-control none --import-source on -k regex:tree_attention -s 114 -c 2 -f -o gpurun_out/prof_attn_$TAG \
    python tools/one_user.py > gpurun_out/ncu_attn_$TAG.log 2>&1; echo "ncu attn rc=$?"
ncu --set full --clock-control none --import-source on -k "regex:mask_logsoftmax|kv_gather|tree_verify|tree_select|residual_rmsnorm|qkv_rope|silu_mul" \
    -s 520 -c 24 -f -o gpurun_out/prof_small_$TAG python tools/one_user.py > gpurun_out/ncu_small_$TAG.log 2>&1; echo "ncu small rc=$?"
tail -c 1500 gpurun_out/bench_$TAG.log
