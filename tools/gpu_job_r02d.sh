#!/bin/bash
# Round 2, GPU session D: fused GEMM epilogues (RoPE + KV append, SiLU * up) -- parity of every tile path, the whole GPU suite,
# bench with and without them.
TAG=${1:-r02d}
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_fused_epilogue.py -q -x -s > $O/fused_tests_$TAG.log 2>&1; echo "fused tests rc=$?"; grep -E "max \||passed|failed|Error|error" $O/fused_tests_$TAG.log | tail -30
timeout 900 python -m pytest tests -m gpu -q -x > $O/gpu_tests_$TAG.log 2>&1; echo "pytest rc=$?"; tail -5 $O/gpu_tests_$TAG.log
for cfg in 1 0; do
  ATSPEED_FUSED_EPI=$cfg timeout 600 python bench.py --gpus 1 --steps 12 --warmup 4 --no-cpu-baseline --hf-baseline-users 0 > $O/bench_fused${cfg}_$TAG.log 2> $O/bench_fused${cfg}_$TAG.err; echo "bench fused=$cfg rc=$?"
  python - <<PY
import json
try:
    j = json.loads(open('$O/bench_fused${cfg}_$TAG.log').read().strip().splitlines()[-1])
    print('fused=$cfg value', round(j['value'], 1), 'e2e', round(j['e2e']['value'], 1), 'p50', round(j['latency_ms_p50'], 2), 'roofline', round(j['roofline']['frac'], 3),
          'consistency', j['pass_consistency'], {k: (round(v['ms_per_user'], 3), round(v['launches_per_user'], 1)) for k, v in j['kernel_groups'].items()})
except Exception as e:
    print('ERR', e); print(open('$O/bench_fused${cfg}_$TAG.err').read()[-1500:])
PY
done
timeout 300 python bench.py --gpus 1 --steps 6 --warmup 3 --cohort 1 --lanes 1 --no-cpu-baseline --hf-baseline-users 0 > $O/bench_single_$TAG.log 2> $O/bench_single_$TAG.err; echo "bench single-search rc=$?"; python -c "
import json; j=json.loads(open('$O/bench_single_$TAG.log').read().strip().splitlines()[-1]); print('single search: value', round(j['value'],1), 'p50', round(j['latency_ms_p50'],2), 'roofline', j['roofline']['bound'], round(j['roofline']['frac'],3))"
