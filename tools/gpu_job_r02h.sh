#!/bin/bash
# Round 2, multi-GPU session: N ranks exactly as the driver launches them (torchrun, NCCL), plus at N >= 2 the end-to-end metrics
# test (sharded users -> NCCL all-gather -> Recall/NDCG identical on every rank and equal to the reference's).
# usage: gpurun --gpus N -- 'bash tools/gpu_job_r02h.sh N'
N=${1:-2}
TAG=${2:-r02h}_n$N
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 600 python -m pytest tests/test_zz_gpu_sharded_metrics.py -q -s > $O/sharded_metrics_$TAG.log 2>&1; echo "sharded metrics rc=$?"; grep -E "passed|failed|skipped|Recall|identical" $O/sharded_metrics_$TAG.log | tail -5
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 12 --warmup 4 \
    > $O/bench_$TAG.log 2> $O/bench_$TAG.err; echo "bench N=$N rc=$?"
python - <<PY
import json
try:
    j = [json.loads(l) for l in open('$O/bench_$TAG.log') if l.startswith('{')][-1]
    print('N', j['n_gpus'], 'value', round(j['value'], 1), 'per GPU', round(j['value'] / j['n_gpus'], 1), 'e2e', round(j['e2e']['value'], 1), 'pair', j['config']['gemm_pair_kernel'],
          'consistency', j['pass_consistency'], 'incomplete' in j, 'roofline', round(j['roofline']['frac'], 3))
except Exception as e:
    print('ERR', e); print(open('$O/bench_$TAG.err').read()[-2500:])
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus $N --steps 2 --warmup 1 \
    > $O/bench_ref_$TAG.log 2> $O/bench_ref_$TAG.err; echo "reference arm N=$N rc=$?"; tail -c 300 $O/bench_ref_$TAG.log
