#!/bin/bash
# Round 2, GPU session Q: cluster-of-4 pair GEMM (two pairs share the activation stream by multicast), ATSPEED_GEMM_CLUSTER=4.
TAG=${1:-r02q}
O=gpurun_out
mkdir -p $O
export ATSPEED_SPIN_LIMIT_MS=2000
echo "== cluster 2 (default) kernel tests"
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -x -k gemm > $O/tests_c2_$TAG.log 2>&1; echo "rc=$?"; tail -2 $O/tests_c2_$TAG.log
echo "== cluster 4 kernel tests"
ATSPEED_GEMM_CLUSTER=4 timeout 300 python -m pytest tests/test_gpu_kernels.py -q -x -k gemm > $O/tests_c4_$TAG.log 2>&1; echo "rc=$?"; tail -5 $O/tests_c4_$TAG.log
echo "== cluster 4 T sweep (257..512 step 17), 7b + 68m shapes"
ATSPEED_GEMM_CLUSTER=4 timeout 600 python tools/gemm_T_sweep_check.py --lo 257 --hi 512 --step 17 > $O/sweep_c4_$TAG.log 2>&1; echo "rc=$?"; tail -5 $O/sweep_c4_$TAG.log
for c in 2 4 2 4; do
  echo "== gemm_bench cluster $c"
  ATSPEED_GEMM_CLUSTER=$c timeout 300 python tools/gemm_bench.py --T 289,400,512 2>&1 | tee $O/gemm_bench_c${c}_$TAG.txt
done
