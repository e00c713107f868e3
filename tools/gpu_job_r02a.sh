#!/bin/bash
# Round 2, GPU session A: parity tests, the driver's exact bench command, pair-kernel soak with diagnostics (opt-in kernel),
# stand-alone roofline of kernels (a)/(c), GEMM micro-bench at cohort sizes, attention variants, ncu captures.
# Every step has its own `timeout`; run under gpurun from the repo root.
TAG=${1:-r02a}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv,noheader
timeout 900 python -m pytest tests -m gpu -q -x > $O/gpu_tests_$TAG.log 2>&1; echo "pytest rc=$?"; tail -4 $O/gpu_tests_$TAG.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench_$TAG.log 2> $O/bench_$TAG.err; echo "bench rc=$?"; tail -c 1500 $O/bench_$TAG.log
# ---- pair kernel under three lanes: does it stall, and if so WHO waits for WHAT (HangDiag) ----
ATSPEED_GEMM_2CTA=1 timeout 150 python tools/soak.py --seconds 40 > $O/soak_pair_$TAG.log 2> $O/soak_pair_$TAG.err; echo "soak pair rc=$?"; cat $O/soak_pair_$TAG.log
ATSPEED_GEMM_2CTA=1 ATSPEED_PDL=0 timeout 150 python tools/soak.py --seconds 30 > $O/soak_pair_nopdl_$TAG.log 2> $O/soak_pair_nopdl_$TAG.err; echo "soak pair no-PDL rc=$?"; cat $O/soak_pair_nopdl_$TAG.log
ATSPEED_GEMM_2CTA=1 timeout 150 python tools/soak.py --seconds 30 --lanes 1 > $O/soak_pair_1lane_$TAG.log 2> $O/soak_pair_1lane_$TAG.err; echo "soak pair 1 lane rc=$?"; cat $O/soak_pair_1lane_$TAG.log
# ---- stand-alone rooflines ----
timeout 300 python tools/kernel_abc_bench.py --json $O/abc_bench_$TAG.json > $O/abc_bench_$TAG.txt 2>&1; echo "abc rc=$?"; cat $O/abc_bench_$TAG.txt
timeout 300 python tools/gemm_bench.py --json $O/gemm_bench_$TAG.json > $O/gemm_bench_$TAG.txt 2>&1; echo "gemm bench rc=$?"; cat $O/gemm_bench_$TAG.txt
ATSPEED_GEMM_2CTA=1 timeout 300 python tools/gemm_bench.py --T 289,400,512 > $O/gemm_bench_pair_$TAG.txt 2>&1; echo "gemm bench pair rc=$?"; cat $O/gemm_bench_pair_$TAG.txt
timeout 120 python tools/att_bench.py > $O/att_bench_$TAG.txt 2>&1; cat $O/att_bench_$TAG.txt
ATSPEED_ATT_BQ=32 timeout 120 python tools/att_bench.py > $O/att_bench_bq32_$TAG.txt 2>&1; cat $O/att_bench_bq32_$TAG.txt
ATSPEED_ATT_PLO=0 timeout 120 python tools/att_bench.py > $O/att_bench_plo0_$TAG.txt 2>&1; cat $O/att_bench_plo0_$TAG.txt
ATSPEED_ATT_BQ=32 timeout 300 python -m pytest tests/test_gpu_kernels.py -q -k attention > $O/att_test_bq32_$TAG.log 2>&1; echo "att bq32 test rc=$?"; tail -2 $O/att_test_bq32_$TAG.log
ATSPEED_ATT_PLO=0 timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_e2e.py -q -k "attention or forward_matches" > $O/att_test_plo0_$TAG.log 2>&1; echo "att plo0 test rc=$?"; tail -2 $O/att_test_plo0_$TAG.log
# ---- ncu: kernels (a)/(c) stand-alone, GEMMs at T = 512 (both only after the plain commands above exited 0) ----
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:mask_logsoftmax|kv_gather" -c 16 -f \
    -o $O/prof_abc_$TAG python tools/kernel_abc_bench.py --iters 1 --quick > $O/ncu_abc_$TAG.log 2>&1; echo "ncu abc rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:gemm_wx" -c 20 -f \
    -o $O/prof_gemm512_$TAG python tools/gemm_bench.py --T 512 --iters 2 > $O/ncu_gemm512_$TAG.log 2>&1; echo "ncu gemm rc=$?"
# launch list of a cohort search (second pass of 8 users): shares per kernel
timeout 300 python tools/one_user.py --cohort 8 --users 8 > $O/plain_c_$TAG.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum,launch__grid_size --clock-control none -c 4000 --csv --log-file $O/launches_cohort_$TAG.csv \
    python tools/one_user.py --cohort 8 --users 8 > $O/ncu_list_c_$TAG.log 2>&1; echo "ncu list cohort rc=$?"
ls -la $O | tail -30
