#!/bin/bash
# Round 2, GPU session G: the measurement matrix through the driver-compatible bench (BASELINE.json configs[1], [2], [4]),
# accepted tokens per verify on a correlated draft, an `ncu --set full` capture of the cohort GEMMs for roofline.traffic,
# and the driver's exact command three times in a row.
TAG=${1:-r02g}
O=gpurun_out
mkdir -p $O
summ() { python - "$1" <<'PY'
import json, sys
try:
    j = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r = j.get('roofline') or {}
    print(sys.argv[1].split('/')[-1], 'value', round(j['value'], 1), 'e2e', round(j['e2e']['value'], 1), 'p50', round(j['latency_ms_p50'], 2), 'acc/verify',
          round(j['accepted_tokens_per_verify'], 2), 'gemm', r.get('bound'), round(r.get('frac', 0), 3), 'incomplete' in j,
          {k: round(v['share'], 3) for k, v in (j.get('kernel_groups') or {}).items()})
except Exception as e:
    print(sys.argv[1], 'ERR', e)
PY
}
timeout 300 python -m pytest tests/test_gpu_prompts.py -q > $O/prompt_tests_$TAG.log 2>&1; echo "prompt tests rc=$?"; tail -2 $O/prompt_tests_$TAG.log
# ---- configs[4]: K x gamma sweep (N = 40; gamma = 4 == gamma = 3 for 4 new tokens, code/beamSD.py:504, kept to show it) ----
: > $O/sweep_$TAG.jsonl
for K in 1 5 10 20; do for G in 2 3 4; do
  timeout 300 python bench.py --K $K --N 40 --gamma $G --steps 4 --warmup 3 --no-cpu-baseline --hf-baseline-users 0 --check-users 0 \
      2> $O/sweep_${TAG}_K${K}_g${G}.err | tail -1 >> $O/sweep_$TAG.jsonl; echo "K=$K gamma=$G rc=${PIPESTATUS[0]}"
done; done
python - <<PY
import json
for l in open('$O/sweep_$TAG.jsonl'):
    try: j = json.loads(l)
    except Exception: continue
    w = j['config']['workload']; k = w.split(' K=')[1].split(' ')[0]; g = w.split('gamma=')[1].split(' ')[0]
    r = j.get('roofline') or {}; kg = j.get('kernel_groups') or {}
    print(f"K={k:>2} gamma={g} users/s={j['value']:7.1f} e2e={j['e2e']['value']:7.1f} p50={j['latency_ms_p50']:6.2f} ms acc/verify={j['accepted_tokens_per_verify']:5.2f} "
          f"gemm {r.get('bound')} frac={r.get('frac', 0):.3f} topk_GB/s={(kg.get('topk') or {}).get('algorithmic_gbs') or 0:.0f} launches/user={sum(v['launches_per_user'] for v in kg.values()):.0f}")
PY
# ---- configs[2]: AtSpeed-R, Games, K = 20 ----
timeout 600 python bench.py --steps 10 --warmup 3 --do-sample --dataset games --K 20 --constraint positional --no-cpu-baseline --hf-baseline-users 0 \
    > $O/bench_relaxed_$TAG.log 2> $O/bench_relaxed_$TAG.err; echo "relaxed rc=$?"; summ $O/bench_relaxed_$TAG.log
# ---- Games K = 20 strict (configs[3]'s per-GPU workload) ----
timeout 600 python bench.py --steps 10 --warmup 3 --dataset games --K 20 --no-cpu-baseline --hf-baseline-users 0 --check-users 0 \
    > $O/bench_games20_$TAG.log 2> $O/bench_games20_$TAG.err; echo "games20 rc=$?"; summ $O/bench_games20_$TAG.log
# ---- accepted tokens per verify with a correlated draft at the benchmark target shape ----
timeout 600 python bench.py --steps 6 --warmup 3 --draft corr24 --no-cpu-baseline --hf-baseline-users 0 --check-users 0 \
    > $O/bench_corr24_$TAG.log 2> $O/bench_corr24_$TAG.err; echo "corr24 rc=$?"; summ $O/bench_corr24_$TAG.log
# ---- ncu --set full of the cohort forward's GEMMs (second pass of 8 users; plain run first) ----
timeout 300 python tools/one_user.py --cohort 8 --users 8 > $O/plain_c_$TAG.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:gemm_wx" -s 900 -c 40 -f -o $O/prof_gemm_cohort_$TAG \
    python tools/one_user.py --cohort 8 --users 8 > $O/ncu_gemm_cohort_$TAG.log 2>&1; echo "ncu cohort gemm rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:tree_attention|residual_rmsnorm|qkv_rope|silu_mul" -s 1500 -c 24 -f -o $O/prof_rowwise_cohort_$TAG \
    python tools/one_user.py --cohort 8 --users 8 > $O/ncu_rowwise_cohort_$TAG.log 2>&1; echo "ncu cohort rowwise rc=$?"
# ---- the driver's command, three times in a row ----
for i in 1 2 3; do
  timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench_final${i}_$TAG.log 2> $O/bench_final${i}_$TAG.err; echo "bench final $i rc=$?"; summ $O/bench_final${i}_$TAG.log
done
timeout 600 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > $O/bench_ref_$TAG.log 2> $O/bench_ref_$TAG.err; echo "reference arm rc=$?"; tail -c 400 $O/bench_ref_$TAG.log
