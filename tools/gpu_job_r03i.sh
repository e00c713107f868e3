#!/bin/bash
# GPU session 3i (closing run of round 2, code as committed): full GPU test tier, the driver's bench command three times in a
# row (complete lines: roofline, cpu_baseline, e2e, HF-on-GPU, parity record), the reference arm, smoke().
TAG=${1:-r03i}
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests/ -m gpu -q > $O/tests_all_$TAG.log 2>&1; echo "tests rc=$?"; tail -3 $O/tests_all_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke()" > $O/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -1 $O/smoke_$TAG.log
for i in 1 2 3; do
  timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench_default${i}_$TAG.log 2> $O/bench_default${i}_$TAG.err; echo "bench default $i rc=$?"
  python - $O/bench_default${i}_$TAG.log <<'PY'
import json, sys
j = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
r = j['roofline']
print('value', round(j['value'], 1), 'e2e', round(j['e2e']['value'], 1), 'p50', round(j['latency_ms_p50'], 2), 'mhz', j['clocks']['sm_mhz'], j['clocks']['reasons'], 'frac', round(r['frac'], 3),
      'prefix', j['config'].get('shared_prompt_prefix_tokens'), {k: round(v['ms_per_user'], 3) for k, v in j['kernel_groups'].items()})
print('  parity', j.get('parity_vs_oracle')); print('  cpu', j.get('cpu_baseline'))
print('  hf', {k: v for k, v in (j.get('hf_gpu_baseline') or {}).items() if k in ('users_per_s', 'latency_ms_p50', 'latency_ms_min', 'speedup_e2e_throughput', 'speedup_single_search')})
PY
done
timeout 300 python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 > $O/bench_ref_$TAG.log 2> $O/bench_ref_$TAG.err; echo "reference arm rc=$?"; tail -c 300 $O/bench_ref_$TAG.log
timeout 600 python bench.py --steps 10 --warmup 3 --do-sample --dataset games --K 20 --constraint positional --no-cpu-baseline --hf-baseline-users 0 \
    > $O/bench_relaxed_$TAG.log 2> $O/bench_relaxed_$TAG.err; echo "relaxed rc=$?"
timeout 600 python bench.py --steps 10 --warmup 3 --dataset games --K 20 --no-cpu-baseline --hf-baseline-users 0 --check-users 0 \
    > $O/bench_games20_$TAG.log 2> $O/bench_games20_$TAG.err; echo "games20 rc=$?"
for f in relaxed games20; do python - $O/bench_${f}_$TAG.log $f <<'PY'
import json, sys
j = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[2], 'value', round(j['value'], 1), 'e2e', round(j['e2e']['value'], 1), 'p50', round(j['latency_ms_p50'], 2), 'mhz', j['clocks']['sm_mhz'], 'frac', round(j['roofline']['frac'], 3), 'acc/verify', round(j['accepted_tokens_per_verify'], 2))
PY
done
