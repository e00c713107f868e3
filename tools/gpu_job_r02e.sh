#!/bin/bash
# Round 2, GPU session E: fused epilogues with prefetched partials / RoPE tables -- parity + bench A/B.
TAG=${1:-r02e}
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_fused_epilogue.py tests/test_gpu_e2e.py tests/test_gpu_cohort.py -q -x > $O/fused_tests_$TAG.log 2>&1; echo "fused+e2e tests rc=$?"; tail -3 $O/fused_tests_$TAG.log
for cfg in 1 0; do
  ATSPEED_FUSED_EPI=$cfg timeout 600 python bench.py --gpus 1 --steps 12 --warmup 4 --no-cpu-baseline --hf-baseline-users 0 > $O/bench_fused${cfg}_$TAG.log 2> $O/bench_fused${cfg}_$TAG.err; echo "bench fused=$cfg rc=$?"
  python - <<PY
import json
try:
    j = json.loads(open('$O/bench_fused${cfg}_$TAG.log').read().strip().splitlines()[-1])
    print('fused=$cfg value', round(j['value'], 1), 'e2e', round(j['e2e']['value'], 1), 'p50', round(j['latency_ms_p50'], 2), 'roofline', round(j['roofline']['frac'], 3), 'avg gemm us', round(j['roofline']['avg_launch_us'], 1),
          {k: (round(v['ms_per_user'], 3), round(v['launches_per_user'], 1)) for k, v in j['kernel_groups'].items()})
except Exception as e:
    print('ERR', e); print(open('$O/bench_fused${cfg}_$TAG.err').read()[-1500:])
PY
done
ATSPEED_FUSED_EPI=1 timeout 300 python bench.py --gpus 1 --steps 6 --warmup 3 --cohort 1 --lanes 1 --no-cpu-baseline --hf-baseline-users 0 > $O/bench_single_fused_$TAG.log 2> /dev/null; python -c "
import json; j=json.loads(open('$O/bench_single_fused_$TAG.log').read().strip().splitlines()[-1]); print('single search fused: value', round(j['value'],1), 'p50', round(j['latency_ms_p50'],2))"
ATSPEED_FUSED_EPI=0 timeout 300 python bench.py --gpus 1 --steps 6 --warmup 3 --cohort 1 --lanes 1 --no-cpu-baseline --hf-baseline-users 0 > $O/bench_single_plain_$TAG.log 2> /dev/null; python -c "
import json; j=json.loads(open('$O/bench_single_plain_$TAG.log').read().strip().splitlines()[-1]); print('single search rowwise: value', round(j['value'],1), 'p50', round(j['latency_ms_p50'],2))"
