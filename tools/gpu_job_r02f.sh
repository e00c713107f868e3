#!/bin/bash
# Round 2, GPU session F: tree attention with the sparse tree phase (parity + timing), fused epilogue A/B after the RoPE fix.
TAG=${1:-r02f}
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x > $O/gpu_tests_$TAG.log 2>&1; echo "pytest rc=$?"; tail -5 $O/gpu_tests_$TAG.log
timeout 120 python tools/att_bench.py > $O/att_bench_$TAG.txt 2>&1; cat $O/att_bench_$TAG.txt
timeout 120 python tools/att_bench.py --P 400 --tree 100 --T 130,512 >> $O/att_bench_$TAG.txt 2>&1; tail -3 $O/att_bench_$TAG.txt
for cfg in 0 1; do
  ATSPEED_FUSED_EPI=$cfg timeout 600 python bench.py --gpus 1 --steps 12 --warmup 4 --no-cpu-baseline --hf-baseline-users 0 > $O/bench_fused${cfg}_$TAG.log 2> $O/bench_fused${cfg}_$TAG.err; echo "bench fused=$cfg rc=$?"
  python - <<PY
import json
try:
    j = json.loads(open('$O/bench_fused${cfg}_$TAG.log').read().strip().splitlines()[-1])
    print('fused=$cfg value', round(j['value'], 1), 'e2e', round(j['e2e']['value'], 1), 'p50', round(j['latency_ms_p50'], 2), 'roofline', round(j['roofline']['frac'], 3), 'avg gemm us', round(j['roofline']['avg_launch_us'], 1),
          {k: (round(v['ms_per_user'], 3), round(v['launches_per_user'], 1)) for k, v in j['kernel_groups'].items()})
except Exception as e:
    print('ERR', e); print(open('$O/bench_fused${cfg}_$TAG.err').read()[-1500:])
PY
done
