#!/bin/bash
# GPU session 3h: shared prompt prefix (K/V of the prompts' common opening tokens computed once per session) -- parity + A/B.
TAG=${1:-r03h}
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_cohort.py tests/test_gpu_cohort_fp32.py -q -x -s > $O/tests_$TAG.log 2>&1; echo "tests rc=$?"; grep -E "passed|failed|error|shared prefix" $O/tests_$TAG.log | tail -8
run() { # name, args...
  n=$1; shift
  timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline --hf-baseline-users 0 "$@" > $O/bench_${n}_$TAG.log 2> $O/bench_${n}_$TAG.err
  python - $O/bench_${n}_$TAG.log "$n" <<'PY'
import json, sys
try:
    j = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], 'value', round(j['value'], 1), 'e2e', round(j['e2e']['value'], 1), 'p50', round(j['latency_ms_p50'], 2), 'mhz', j['clocks']['sm_mhz'], 'per-GHz', round(j['value'] / j['clocks']['sm_mhz'] * 1000, 1),
          'prefix', j['config'].get('shared_prompt_prefix_tokens'), 'consistency', j.get('pass_consistency'), 'parity', j.get('parity_vs_oracle'), 'frac', round(j['roofline']['frac'], 3))
except Exception as e:
    print(sys.argv[2], 'ERR', e)
PY
}
run prefix_on
run prefix_off --no-shared-prefix
run prefix_on2
run prefix_off2 --no-shared-prefix
