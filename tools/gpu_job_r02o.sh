#!/bin/bash
# Round 2, GPU session O: L2 look-ahead of the weight tiles (ATSPEED_GEMM_PREFETCH = distance in k-blocks), same-box A/B.
TAG=${1:-r02o}
O=gpurun_out
mkdir -p $O
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
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -x > $O/tests_$TAG.log 2>&1; echo "tests rc=$?"; tail -2 $O/tests_$TAG.log
for pf in 0 4 8 16 32; do
  echo "== prefetch distance $pf"
  ATSPEED_GEMM_PREFETCH=$pf timeout 300 python tools/gemm_bench.py --T 10,130,289,400,512 2>&1 | tee $O/gemm_bench_pf${pf}_$TAG.txt
done
for pf in 0 8 0 8; do
  ATSPEED_GEMM_PREFETCH=$pf timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline --hf-baseline-users 0 --check-users 0 > $O/bench_pf${pf}_$TAG.log 2> $O/bench_pf${pf}_$TAG.err
  summ $O/bench_pf${pf}_$TAG.log "prefetch=$pf"
done
