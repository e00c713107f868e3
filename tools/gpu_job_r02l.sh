#!/bin/bash
# Round 2, GPU session L: attention sparse phase with key-list compaction + next-key prefetch, early first tile.
TAG=${1:-r02l}
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
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_cohort.py tests/test_gpu_e2e.py -q -x > $O/tests_$TAG.log 2>&1; echo "tests rc=$?"; tail -2 $O/tests_$TAG.log
timeout 300 python tools/att_bench.py > $O/att_bench_$TAG.txt 2>&1; cat $O/att_bench_$TAG.txt
timeout 300 python tools/att_bench.py --P 200 --tree 300 >> $O/att_bench_$TAG.txt 2>&1; tail -4 $O/att_bench_$TAG.txt
for i in 1 2; do
  timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline --hf-baseline-users 0 --check-users 0 > $O/bench_att${i}_$TAG.log 2> $O/bench_att${i}_$TAG.err
  summ $O/bench_att${i}_$TAG.log "att run $i"
done
