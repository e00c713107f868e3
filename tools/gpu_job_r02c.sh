#!/bin/bash
# Round 2, GPU session C: the CTA-pair GEMM after the fix (cluster barrier before tcgen05.alloc.cta_group::2): soaks WITHOUT the
# progress trace (the trace's extra stores hid the stall in session B), then the driver's bench command with the pair kernel.
TAG=${1:-r02c}
O=gpurun_out
mkdir -p $O
ATSPEED_GEMM_2CTA=1 timeout 120 python tools/soak.py --lanes 1 --seconds 30 --stall-s 10 > $O/soak_pair_1lane_$TAG.log 2> $O/soak_pair_1lane_$TAG.err; echo "soak pair 1 lane rc=$?"; head -c 3000 $O/soak_pair_1lane_$TAG.log
ATSPEED_GEMM_2CTA=1 timeout 150 python tools/soak.py --lanes 3 --seconds 60 --stall-s 10 > $O/soak_pair_3lanes_$TAG.log 2> $O/soak_pair_3lanes_$TAG.err; echo "soak pair 3 lanes rc=$?"; head -c 3000 $O/soak_pair_3lanes_$TAG.log
ATSPEED_GEMM_2CTA=1 ATSPEED_PDL=0 timeout 150 python tools/soak.py --lanes 3 --seconds 30 --stall-s 10 > $O/soak_pair_nopdl_$TAG.log 2> $O/soak_pair_nopdl_$TAG.err; echo "soak pair 3 lanes no PDL rc=$?"; head -c 3000 $O/soak_pair_nopdl_$TAG.log
for i in 1 2; do
  ATSPEED_GEMM_2CTA=1 timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --hf-baseline-users 0 > $O/bench_pair${i}_$TAG.log 2> $O/bench_pair${i}_$TAG.err; echo "bench pair $i rc=$?"
  python - <<PY
import json
try:
    j = json.loads(open('$O/bench_pair${i}_$TAG.log').read().strip().splitlines()[-1])
    print('value', round(j['value'], 1), 'e2e', round(j['e2e']['value'], 1), 'p50', round(j['latency_ms_p50'], 2), 'roofline', round(j['roofline']['frac'], 3), j['roofline']['kernel'], 'consistency', j['pass_consistency'], 'incomplete' in j)
except Exception as e:
    print('ERR', e)
PY
done
ATSPEED_GEMM_2CTA=1 timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_cohort.py -q > $O/tests_pair_$TAG.log 2>&1; echo "tests with pair kernel rc=$?"; tail -2 $O/tests_pair_$TAG.log
