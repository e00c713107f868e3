#!/bin/bash
# GPU session 3d: register-store epilogue as the only epilogue -- full GPU tier, T sweep, bench.
TAG=${1:-r03d}
O=gpurun_out
export ATSPEED_SPIN_LIMIT_MS=2000
timeout 900 python tools/gemm_T_sweep_check.py --lo 1 --hi 512 --step 3 > $O/sweep_$TAG.log 2>&1; echo "sweep rc=$?"; tail -3 $O/sweep_$TAG.log
timeout 1500 python -m pytest tests/ -m gpu -q -x > $O/tests_all_$TAG.log 2>&1; echo "tests rc=$?"; tail -3 $O/tests_all_$TAG.log
timeout 300 python tools/gemm_bench.py 2>&1 | tee $O/gemm_bench_$TAG.txt
for i in 1 2; do
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline --hf-baseline-users 0 --check-users 0 > $O/bench_${i}_$TAG.log 2> $O/bench_${i}_$TAG.err
python - <<PY
import json
j = json.loads(open('$O/bench_${i}_$TAG.log').read().strip().splitlines()[-1])
print('value', round(j['value'],1), 'e2e', round(j['e2e']['value'],1), 'mhz', j['clocks']['sm_mhz'], 'per-GHz', round(j['value']/j['clocks']['sm_mhz']*1000,1), {k: round(v['ms_per_user'],3) for k, v in j['kernel_groups'].items()}, 'frac', round(j['roofline']['frac'],3))
PY
done
