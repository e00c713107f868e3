#!/bin/bash
# quick GPU check: kernel + e2e parity tests, GEMM micro-benchmark, bench line
TAG=${1:-q}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q > gpurun_out/tests_kernels_$TAG.log 2>&1; echo "kernels rc=$?"; tail -3 gpurun_out/tests_kernels_$TAG.log
timeout 600 python -m pytest tests -m gpu -q --deselect tests/test_gpu_kernels.py > gpurun_out/tests_e2e_$TAG.log 2>&1; echo "e2e rc=$?"; tail -3 gpurun_out/tests_e2e_$TAG.log
timeout 300 python tools/gemm_bench.py > gpurun_out/gemm_bench_$TAG.txt 2>&1; echo "gemm bench rc=$?"
timeout 600 python bench.py --no-cpu-baseline --hf-baseline-users 0 > gpurun_out/bench_$TAG.log 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
python - <<'PY'
import json,sys,glob
try:
    j=json.loads(open(sorted(glob.glob('gpurun_out/bench_*.log'))[-1]).read().strip().splitlines()[-1])
    print({k:j[k] for k in ('value','ms_per_step','latency_ms_p50','gpu_launches')}, j['e2e']['value'])
    print({k:(round(v['ms_per_user'],3)) for k,v in j['kernel_groups'].items()}, j['roofline']['frac'])
except Exception as e: print('no bench', e)
PY
