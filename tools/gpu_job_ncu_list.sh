#!/bin/bash
TAG=${1:-l}
python tools/one_user.py > gpurun_out/plain_$TAG.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,launch__grid_size --clock-control none -s 1040 -c 1100 --csv --log-file gpurun_out/launches_$TAG.csv \
    python tools/one_user.py > gpurun_out/ncu_list_$TAG.log 2>&1; echo "ncu list rc=$?"
for cfg in "" "ATSPEED_PDL=0" "ATSPEED_GEMM_BM=128" "ATSPEED_GEMM_BUFS=1"; do
  echo "== $cfg"; env $cfg python bench.py --no-cpu-baseline --hf-baseline-users 0 --steps 2 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(round(j['value'],2), round(j['latency_ms_p50'],2), {k:round(v['ms_per_user'],2) for k,v in j['kernel_groups'].items()})"
done
