#!/bin/bash
for cfg in "1 4" "4 1" "8 1" "16 1" "8 2" "16 2" "8 3"; do set -- $cfg
python bench.py --no-cpu-baseline --hf-baseline-users 0 --cohort $1 --lanes $2 --users-per-step 48 --steps 3 2>gpurun_out/bench_cohort.err | python -c "
import json,sys
try:
    j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cohort', $1, 'lanes', $2, round(j['value'],2), round(j['e2e']['value'],2), round(j['latency_ms_p50'],2), round(j['latency_ms_p50_loaded'],2), {k:round(v['ms_per_user'],2) for k,v in j['kernel_groups'].items()}, j['gpu_launches'])
except Exception as e: print('cohort', $1, 'lanes', $2, 'ERR', e)"; tail -3 gpurun_out/bench_cohort.err; done
