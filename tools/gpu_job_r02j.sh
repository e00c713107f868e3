#!/bin/bash
# Round 2, GPU session J: L2 eviction hints on the GEMM's TMA loads (A/B), lanes x cohort sweep, the bench line with the fp32
# yardstick of the HF comparison.
TAG=${1:-r02j}
O=gpurun_out
mkdir -p $O
summ() { python - "$1" "$2" <<'PY'
import json, sys
try:
    j = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r = j.get('roofline') or {}
    print(sys.argv[2], 'value', round(j['value'], 1), 'e2e', round(j['e2e']['value'], 1), 'p50', round(j['latency_ms_p50'], 2), 'gemm', round(r.get('frac', 0), 3),
          {k: round(v['ms_per_user'], 3) for k, v in (j.get('kernel_groups') or {}).items()})
except Exception as e:
    print(sys.argv[2], 'ERR', e)
PY
}
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_e2e.py tests/test_gpu_cohort.py tests/test_gpu_fused_epilogue.py -q -x > $O/tests_$TAG.log 2>&1; echo "tests rc=$?"; tail -2 $O/tests_$TAG.log
for h in 1 0 1 0; do
  ATSPEED_GEMM_L2HINT=$h timeout 600 python bench.py --gpus 1 --steps 12 --warmup 4 --no-cpu-baseline --hf-baseline-users 0 --check-users 0 > $O/bench_l2hint${h}_$TAG.log 2> $O/bench_l2hint${h}_$TAG.err
  summ $O/bench_l2hint${h}_$TAG.log "l2hint=$h"
done
for lanes in 2 4 6; do for cohort in 8 16; do
  timeout 600 python bench.py --gpus 1 --steps 8 --warmup 3 --lanes $lanes --cohort $cohort --no-cpu-baseline --hf-baseline-users 0 --check-users 0 > $O/bench_l${lanes}_c${cohort}_$TAG.log 2> $O/bench_l${lanes}_c${cohort}_$TAG.err
  summ $O/bench_l${lanes}_c${cohort}_$TAG.log "lanes=$lanes cohort=$cohort"
done; done
timeout 600 python bench.py --gpus 1 --steps 8 --warmup 3 --lanes 3 --cohort 16 --no-cpu-baseline --hf-baseline-users 0 --check-users 0 > $O/bench_l3_c16_$TAG.log 2> /dev/null; summ $O/bench_l3_c16_$TAG.log "lanes=3 cohort=16"
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench_default_$TAG.log 2> $O/bench_default_$TAG.err; echo "bench default rc=$?"; summ $O/bench_default_$TAG.log default
python - <<PY
import json
j = json.loads(open('$O/bench_default_$TAG.log').read().strip().splitlines()[-1])
print(j.get('parity_vs_oracle')); print({k: v for k, v in (j.get('hf_gpu_baseline') or {}).items() if k != 'what'})
PY
