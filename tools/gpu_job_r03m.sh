#!/bin/bash
# GPU session 3m: last configuration A/B with the shared prefix on: deferral of under-filled packs, lanes.
TAG=${1:-r03m}
O=gpurun_out
run() { n=$1; shift; e=$1; shift
  env $e ATSPEED_COHORT_LOG=1 timeout 600 python bench.py --gpus 1 --steps 8 --warmup 3 --no-cpu-baseline --hf-baseline-users 0 --check-users 0 "$@" > $O/bench_${n}_$TAG.log 2> $O/bench_${n}_$TAG.err
  python - $O/bench_${n}_$TAG.log "$n" $O/bench_${n}_$TAG.err <<'PY'
import json, sys, statistics
try:
    j = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    T = [int(l.split()[2].split('=')[1]) for l in open(sys.argv[3]) if l.startswith('atspeed-pack')]
    print(sys.argv[2], 'value', round(j['value'], 1), 'e2e', round(j['e2e']['value'], 1), 'mhz', j['clocks']['sm_mhz'], 'per-GHz', round(j['value'] / j['clocks']['sm_mhz'] * 1000, 1), 'packs', len(T), 'mean T', round(statistics.mean(T), 1), 'frac', round(j['roofline']['frac'], 3))
except Exception as e:
    print(sys.argv[2], 'ERR', e)
PY
}
run base "A=1"
run nodefer "ATSPEED_COHORT_DEFER=0"
run lanes3 "A=1" --lanes 3
run base2 "A=1"
run nodefer2 "ATSPEED_COHORT_DEFER=0"
