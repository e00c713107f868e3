#!/bin/bash
# First GPU call of round 2 (1 GPU): everything that was left unverified when round 1's GPU budget ran out.
#   1. GPU parity tests (incl. the new timing-loop test)        2. default bench line
#   3. rank 0-of-8's user slice on ONE GPU (is the 8-GPU stall data-dependent?)   4. every T through the GEMM (pair kernel on)
#   5. kernels (a)/(c) stand-alone roofline
# Run the multi-GPU part separately:  gpurun --gpus 2 -- 'bash tools/gpu_job_multigpu.sh 2'   then with
# ATSPEED_GEMM_2CTA=1 exported, then --gpus 8.  bench.py's breadcrumbs (stderr) name the phase of any stall.
TAG=${1:-r02a}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/tests_$TAG.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/tests_$TAG.log
timeout 600 python bench.py > gpurun_out/bench_$TAG.log 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; tail -c 400 gpurun_out/bench_$TAG.log
for r in 0 3; do
  timeout 300 python bench.py --emulate-shard $r/8 --steps 6 --warmup 3 --no-cpu-baseline --hf-baseline-users 0 \
      > gpurun_out/bench_shard${r}of8_$TAG.log 2> gpurun_out/bench_shard${r}of8_$TAG.err; echo "emulate shard $r/8 rc=$?"
  tail -2 gpurun_out/bench_shard${r}of8_$TAG.err
done
timeout 900 python tools/gemm_T_sweep_check.py --lo 250 --hi 512 > gpurun_out/gemm_T_sweep_$TAG.txt 2>&1; echo "gemm T sweep rc=$?"; tail -12 gpurun_out/gemm_T_sweep_$TAG.txt
timeout 600 python tools/kernel_abc_bench.py --json gpurun_out/abc_bench_$TAG.json > gpurun_out/abc_bench_$TAG.txt 2>&1; echo "abc rc=$?"; cat gpurun_out/abc_bench_$TAG.txt
# experimental kernels written blind at the end of round 1: correctness first, then timing
ATSPEED_ATT_BQ=32 timeout 300 python -m pytest tests/test_gpu_kernels.py -q -k attention > gpurun_out/att_bq32_$TAG.log 2>&1; echo "attention BQ=32 rc=$?"; tail -3 gpurun_out/att_bq32_$TAG.log
ATSPEED_GEMM_4CTA=1 timeout 600 python tools/gemm_T_sweep_check.py --lo 257 --hi 512 --shapes 7b --limit-s 15 > gpurun_out/gemm_4cta_check_$TAG.txt 2>&1; echo "4-CTA GEMM check rc=$?"; tail -8 gpurun_out/gemm_4cta_check_$TAG.txt
ATSPEED_ATT_PLO=0 timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_e2e.py -q -k "attention or forward_matches" > gpurun_out/att_plo0_$TAG.log 2>&1; echo "attention PLO=0 rc=$?"; tail -3 gpurun_out/att_plo0_$TAG.log
ATSPEED_TOPK_UNROLL=8 timeout 600 python -m pytest tests/test_gpu_kernels.py -q -k topk > gpurun_out/topk_unr8_$TAG.log 2>&1; echo "kernel (a) unroll 8 rc=$?"; tail -2 gpurun_out/topk_unr8_$TAG.log
ATSPEED_TOPK_UNROLL=8 timeout 600 python tools/kernel_abc_bench.py --quick > gpurun_out/abc_bench_unr8_$TAG.txt 2>&1; cat gpurun_out/abc_bench_unr8_$TAG.txt
