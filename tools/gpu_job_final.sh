#!/bin/bash
# lanes sweep, relaxed-mode bench (configs[2]), K=20 games, and the default bench line with cpu/HF baselines
TAG=${1:-f}
mkdir -p gpurun_out
for l in 4 6 8; do python bench.py --no-cpu-baseline --hf-baseline-users 0 --lanes $l --users-per-step 16 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('lanes', $l, round(j['value'],2), round(j['e2e']['value'],2), round(j['latency_ms_p50'],2), round(j['latency_ms_p50_loaded'],2))"; done
python bench.py --no-cpu-baseline --hf-baseline-users 0 --do-sample --dataset games --K 20 --constraint positional > gpurun_out/bench_relaxed_$TAG.log 2>gpurun_out/bench_relaxed_$TAG.err; echo "relaxed rc=$?"
python bench.py --no-cpu-baseline --hf-baseline-users 0 --dataset games --K 20 > gpurun_out/bench_games20_$TAG.log 2>gpurun_out/bench_games20_$TAG.err; echo "games20 rc=$?"
python bench.py > gpurun_out/bench_$TAG.log 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$TAG.log 2> gpurun_out/bench_ref_$TAG.err; echo "ref rc=$?"
for f in bench_relaxed bench_games20 bench; do python - <<PY
import json
try:
    j=json.loads(open('gpurun_out/${f}_$TAG.log').read().strip().splitlines()[-1])
    print('$f', round(j['value'],2), round(j['e2e']['value'],2), j['latency_ms_p50'], j.get('accepted_tokens_per_verify'), j.get('hf_gpu_baseline'), j.get('cpu_baseline'))
except Exception as e: print('$f', 'ERR', e)
PY
done
tail -c 600 gpurun_out/bench_ref_$TAG.log
