#!/bin/bash
# GPU session 3a: best-fit-decreasing packing of the target forwards with deferral of under-filled packs -- parity + A/B.
TAG=${1:-r03a}
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_cohort.py tests/test_gpu_cohort_fp32.py -q -x > $O/tests_$TAG.log 2>&1; echo "tests rc=$?"; tail -2 $O/tests_$TAG.log
run() { # name, env, args...
  n=$1; shift; e=$1; shift
  env $e timeout 900 python bench.py --gpus 1 --no-cpu-baseline --hf-baseline-users 0 --check-users 0 "$@" > $O/bench_${n}_$TAG.log 2> $O/bench_${n}_$TAG.err
  python - $O/bench_${n}_$TAG.log "$n" $O/bench_${n}_$TAG.err <<'PY'
import json, sys, statistics
try:
    j = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    T = [int(l.split()[2].split('=')[1]) for l in open(sys.argv[3]) if l.startswith('atspeed-pack')]
    print(sys.argv[2], 'value', round(j['value'], 1), 'e2e', round(j['e2e']['value'], 1), 'p50', round(j['latency_ms_p50'], 2), 'loaded p50', round(j['latency_ms_p50_loaded'], 1), 'mhz', j['clocks']['sm_mhz'], 'per-GHz', round(j['value'] / j['clocks']['sm_mhz'] * 1000, 1),
          'packs', len(T), 'mean T', round(statistics.mean(T), 1) if T else None, 'consistency', j.get('pass_consistency'))
except Exception as e:
    print(sys.argv[2], 'ERR', e)
PY
}
run defer0_c8 "ATSPEED_COHORT_DEFER=0 ATSPEED_COHORT_LOG=1" --steps 8 --warmup 3
run defer1_c8 "ATSPEED_COHORT_DEFER=1 ATSPEED_COHORT_LOG=1" --steps 8 --warmup 3
run defer1_c16 "ATSPEED_COHORT_DEFER=1 ATSPEED_COHORT_LOG=1" --steps 8 --warmup 3 --cohort 16
run defer1_c16_l2 "ATSPEED_COHORT_DEFER=1 ATSPEED_COHORT_LOG=1" --steps 8 --warmup 3 --cohort 16 --lanes 2
run defer1_c16_u96 "ATSPEED_COHORT_DEFER=1 ATSPEED_COHORT_LOG=1" --steps 4 --warmup 3 --cohort 16 --users-per-step 96
run defer0_c16 "ATSPEED_COHORT_DEFER=0 ATSPEED_COHORT_LOG=1" --steps 8 --warmup 3 --cohort 16
