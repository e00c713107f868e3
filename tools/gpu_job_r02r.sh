#!/bin/bash
# Round 2, GPU session R: is row-wise consumer time on the critical path of the 3-lane bench?  (consumers run twice: idempotent)
TAG=${1:-r02r}
O=gpurun_out
summ() { python - "$1" "$2" <<'PY'
import json, sys
try:
    j = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r = j.get('roofline') or {}
    print(sys.argv[2], 'value', round(j['value'], 1), 'e2e', round(j['e2e']['value'], 1), 'p50', round(j['latency_ms_p50'], 2), 'mhz', j['clocks']['sm_mhz'], 'gemm', round(r.get('frac', 0), 3),
          {k: round(v['ms_per_user'], 3) for k, v in (j.get('kernel_groups') or {}).items()})
except Exception as e:
    print(sys.argv[2], 'ERR', e)
PY
}
for d in 0 1 0 1; do
  ATSPEED_DEBUG_DUP_ROWWISE=$d timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline --hf-baseline-users 0 --check-users 0 > $O/bench_dup${d}_$TAG.log 2> $O/bench_dup${d}_$TAG.err
  summ $O/bench_dup${d}_$TAG.log "dup_rowwise=$d"
done
for c in 4 2; do
  ATSPEED_GEMM_CLUSTER=$c timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline --hf-baseline-users 0 --check-users 0 > $O/bench_cl${c}_$TAG.log 2> $O/bench_cl${c}_$TAG.err
  summ $O/bench_cl${c}_$TAG.log "cluster=$c"
done
