#!/bin/bash
# Round 2, GPU session X: pair-kernel token padding 32 instead of 64 -- every T in 257..512 against torch, bench.
TAG=${1:-r02x}
O=gpurun_out
export ATSPEED_SPIN_LIMIT_MS=2000
timeout 900 python tools/gemm_T_sweep_check.py --lo 257 --hi 512 --shapes 7b > $O/sweep_pad32_$TAG.log 2>&1; echo "sweep rc=$?"; tail -7 $O/sweep_pad32_$TAG.log
ATSPEED_GEMM_CLUSTER=4 timeout 600 python tools/gemm_T_sweep_check.py --lo 257 --hi 512 --step 5 --shapes 7b > $O/sweep_pad32_c4_$TAG.log 2>&1; echo "sweep c4 rc=$?"; tail -3 $O/sweep_pad32_c4_$TAG.log
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_cohort.py -q -x 2>&1 | tail -2
timeout 300 python tools/gemm_bench.py --T 289,300,400,420,480 2>&1 | tee $O/gemm_bench_pad32_$TAG.txt
for i in 1 2; do
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline --hf-baseline-users 0 --check-users 0 > $O/bench_pad32_${i}_$TAG.log 2> $O/bench_pad32_${i}_$TAG.err
python - <<PY
import json
j = json.loads(open('$O/bench_pad32_${i}_$TAG.log').read().strip().splitlines()[-1])
print('value', round(j['value'],1), 'mhz', j['clocks']['sm_mhz'], 'per-GHz', round(j['value']/j['clocks']['sm_mhz']*1000,1), {k: round(v['ms_per_user'],3) for k, v in j['kernel_groups'].items()}, 'frac', round(j['roofline']['frac'],3))
PY
done
