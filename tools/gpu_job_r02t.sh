#!/bin/bash
# Round 2, GPU session T: row-wise consumers with every slice load issued up front (sum_slices), silu rows per thread.
TAG=${1:-r02t}
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
for nr in 1 2 4; do echo "== silu NR=$nr"; ATSPEED_SILU_NR=$nr timeout 120 python tools/rowwise_bench.py --T 289,480 2>&1 | tee $O/rowwise_nr${nr}_$TAG.txt; done
timeout 900 python -m pytest tests/test_gpu_e2e.py tests/test_gpu_cohort.py tests/test_gpu_fused_epilogue.py -q -x > $O/tests_$TAG.log 2>&1; echo "tests rc=$?"; tail -2 $O/tests_$TAG.log
for nr in 2 4 2; do
  ATSPEED_SILU_NR=$nr timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline --hf-baseline-users 0 --check-users 0 > $O/bench_nr${nr}_$TAG.log 2> $O/bench_nr${nr}_$TAG.err
  summ $O/bench_nr${nr}_$TAG.log "silu_nr=$nr"
done
