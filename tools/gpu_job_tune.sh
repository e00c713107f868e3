#!/bin/bash
# Scheduler-level tuning of the default bench configuration (no kernel changes): cohorts in flight (lanes), users per cohort,
# tokens packed per forward.  One line per point: users/s (device-resident), e2e users/s, single-search p50.
TAG=${1:-r02}
mkdir -p gpurun_out
: > gpurun_out/tune_$TAG.txt
run() {
  python bench.py --no-cpu-baseline --hf-baseline-users 0 --steps 3 --warmup 3 "$@" 2>/dev/null | python -c "
import json,sys
try:
    j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$*', round(j['value'],1), round(j['e2e']['value'],1), round(j['latency_ms_p50'],2), (j.get('pass_consistency') or {}).get('identical_ranked_lists'))
except Exception as e: print('$*', 'ERR', e)" | tee -a gpurun_out/tune_$TAG.txt
}
for lanes in 2 3 4; do run --lanes $lanes --cohort 8; done
for cohort in 6 12 16; do run --lanes 3 --cohort $cohort; done
for tok in 384 448; do run --lanes 3 --cohort 8 --cohort-tokens $tok; done
run --lanes 2 --cohort 16
run --lanes 4 --cohort 6
