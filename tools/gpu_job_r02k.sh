#!/bin/bash
# Round 2, GPU session K: pipelined GEMM epilogue (tcgen05.ld ping-pong, 3-tile staging ring, one barrier per chunk, early
# TMEM release) and the L2 priority of the partial-sum stores.
TAG=${1:-r02k}
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
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_fused_epilogue.py tests/test_gpu_cohort.py -q -x > $O/tests_$TAG.log 2>&1; echo "tests rc=$?"; tail -2 $O/tests_$TAG.log
timeout 300 python tools/gemm_bench.py --T 130,220,289,400,512 > $O/gemm_bench_pair_$TAG.txt 2>&1; cat $O/gemm_bench_pair_$TAG.txt
for h in 0 1 2 0 1; do
  ATSPEED_GEMM_STORE_HINT=$h timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline --hf-baseline-users 0 --check-users 0 > $O/bench_sthint${h}_$TAG.log 2> $O/bench_sthint${h}_$TAG.err
  summ $O/bench_sthint${h}_$TAG.log "store_hint=$h"
done
