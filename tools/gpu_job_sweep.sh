#!/bin/bash
# BASELINE.json configs[4]: beam-width / draft-length sweep K in {1,5,10,20} x gamma in {2,3,4} (gamma=4 == gamma=3 for 4 new
# tokens, code/beamSD.py:504, kept to show it).  One bench line per point (users/s, p50, accepted tokens/verify, GEMM roofline,
# per-kernel-group share and algorithmic GB/s) appended to gpurun_out/sweep_$TAG.jsonl; N = max(K, 40).
TAG=${1:-r02}
mkdir -p gpurun_out
: > gpurun_out/sweep_$TAG.jsonl
for K in 1 5 10 20; do
  for G in 2 3 4; do
    timeout 300 python bench.py --K $K --N 40 --gamma $G --steps 3 --warmup 3 --no-cpu-baseline --hf-baseline-users 0 \
        2> gpurun_out/sweep_${TAG}_K${K}_g${G}.err | tail -1 >> gpurun_out/sweep_$TAG.jsonl
    echo "K=$K gamma=$G rc=${PIPESTATUS[0]}"
  done
done
python - <<PY
import json
for l in open('gpurun_out/sweep_$TAG.jsonl'):
    try:
        j = json.loads(l)
    except Exception:
        continue
    w = j['config']['workload']
    k = w.split(' K=')[1].split(' ')[0]; g = w.split('gamma=')[1].split(' ')[0]
    r = j.get('roofline') or {}
    print(f"K={k:>2} gamma={g} users/s={j['value']:7.1f} e2e={j['e2e']['value']:7.1f} p50={j['latency_ms_p50']:6.2f} ms "
          f"acc/verify={j['accepted_tokens_per_verify']:5.2f} gemm {r.get('bound')} frac={r.get('frac', 0):.3f}")
PY
