#!/bin/bash
# quick GPU check: all GPU parity tests + the default bench line
TAG=${1:-q}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/tests_$TAG.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/tests_$TAG.log
timeout 300 python bench.py --no-cpu-baseline --hf-baseline-users 0 > gpurun_out/bench_$TAG.log 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
python - <<PY
import json
try:
    j=json.loads(open('gpurun_out/bench_$TAG.log').read().strip().splitlines()[-1])
    print({k:j[k] for k in ('value','ms_per_step','latency_ms_p50','gpu_launches')}, j['e2e']['value'])
    print({k:(round(v['ms_per_user'],3)) for k,v in j['kernel_groups'].items()}, j['roofline']['bound'], j['roofline']['frac'])
except Exception as e: print('no bench', e)
PY
