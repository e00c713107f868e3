#!/bin/bash
timeout 200 python -m pytest tests/test_gpu_kernels.py -q -k gemm 2>&1 | tail -4
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
timeout 300 python tools/gemm_sweep.py bigT 2>&1 | head -3
for cfg in "8 3 512" "8 3 768" "12 3 1024" "16 2 1024" "12 1 1024"; do set -- $cfg
python bench.py --no-cpu-baseline --hf-baseline-users 0 --cohort $1 --lanes $2 --cohort-tokens $3 --users-per-step 48 --steps 3 2>gpurun_out/bench_cohort.err | python -c "
import json,sys
try:
    j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cohort', $1, 'lanes', $2, 'tokens', $3, round(j['value'],2), round(j['e2e']['value'],2), round(j['latency_ms_p50'],2), round(j['latency_ms_p50_loaded'],2), {k:round(v['ms_per_user'],2) for k,v in j['kernel_groups'].items()}, j['gpu_launches'], j['roofline']['bound'], round(j['roofline']['frac'],3), round(j['roofline']['roofline_time_frac'],3))
except Exception as e: print('cohort', $1, 'lanes', $2, 'ERR', e)"; tail -3 gpurun_out/bench_cohort.err; done
