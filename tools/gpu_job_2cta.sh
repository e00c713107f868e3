#!/bin/bash
timeout 300 python tools/gemm_sweep.py 2cta 2>&1 | head -4
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
for cfg in "1" "0"; do
ATSPEED_GEMM_2CTA=$cfg python bench.py --no-cpu-baseline --hf-baseline-users 0 --steps 3 2>gpurun_out/bench_2cta.err | python -c "
import json,sys
try:
    j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('2cta=$cfg', round(j['value'],2), round(j['e2e']['value'],2), round(j['latency_ms_p50'],2), {k:round(v['ms_per_user'],2) for k,v in j['kernel_groups'].items()}, j['roofline']['bound'], round(j['roofline']['frac'],3))
except Exception as e: print('ERR', e)"; tail -2 gpurun_out/bench_2cta.err; done
