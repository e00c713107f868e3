#!/bin/bash
# Round 2, GPU session W: full GPU test tier + default bench line after the row-wise / attention / epilogue work.
TAG=${1:-r02w}
O=gpurun_out
timeout 1500 python -m pytest tests/ -m gpu -q -x > $O/tests_all_$TAG.log 2>&1; echo "tests rc=$?"; tail -3 $O/tests_all_$TAG.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench_default_$TAG.log 2> $O/bench_default_$TAG.err; echo "bench rc=$?"
python - <<PY
import json
j = json.loads(open('$O/bench_default_$TAG.log').read().strip().splitlines()[-1])
print('value', round(j['value'],1), 'e2e', round(j['e2e']['value'],1), 'p50', round(j['latency_ms_p50'],2), 'clocks', j['clocks'])
print({k: (round(v['ms_per_user'],3), round(v['launches_per_user'],1)) for k, v in j['kernel_groups'].items()})
print('roofline', {k: j['roofline'][k] for k in ('frac','achieved','avg_launch_us','flops_per_launch')})
print(j.get('parity_vs_oracle')); print({k: v for k, v in (j.get('hf_gpu_baseline') or {}).items() if k not in ('what','prompt_logits_vs_ours','parity_vs_ours')})
print(j.get('cpu_baseline'))
PY
