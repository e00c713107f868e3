#!/bin/bash
# Round 2, GPU session Z: batch size of a bench step (drain bubble at the end of every step) and lanes x cohort after the kernel work.
TAG=${1:-r02z}
O=gpurun_out
run() { # name, args...
  n=$1; shift
  timeout 900 python bench.py --gpus 1 --no-cpu-baseline --hf-baseline-users 0 --check-users 0 "$@" > $O/bench_${n}_$TAG.log 2> $O/bench_${n}_$TAG.err
  python - $O/bench_${n}_$TAG.log "$n" <<'PY'
import json, sys
try:
    j = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], 'value', round(j['value'], 1), 'e2e', round(j['e2e']['value'], 1), 'ms/step', round(j['ms_per_step'], 1), 'mhz', j['clocks']['sm_mhz'], 'per-GHz', round(j['value'] / j['clocks']['sm_mhz'] * 1000, 1))
except Exception as e:
    print(sys.argv[2], 'ERR', e)
PY
}
run u48 --users-per-step 48 --steps 8 --warmup 3
run u96 --users-per-step 96 --steps 4 --warmup 3
run u192 --users-per-step 192 --steps 3 --warmup 3
run u96_l2 --users-per-step 96 --steps 4 --warmup 3 --lanes 2
run u96_l4 --users-per-step 96 --steps 4 --warmup 3 --lanes 4
run u96_c16 --users-per-step 96 --steps 4 --warmup 3 --cohort 16
run u96_l2_c16 --users-per-step 96 --steps 4 --warmup 3 --lanes 2 --cohort 16
run u48b --users-per-step 48 --steps 8 --warmup 3
