#!/bin/bash
# GPU session 3g (final evidence of round 2, kernels as committed): full GPU test tier, ncu launch list + --set full tables of
# the cohort forward, stand-alone kernel benches, the bench lines of BASELINE's configurations (default three times in a row),
# the K x gamma sweep.
TAG=${1:-r03g}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv,noheader
timeout 1500 python -m pytest tests/ -m gpu -q > $O/tests_all_$TAG.log 2>&1; echo "tests rc=$?"; tail -3 $O/tests_all_$TAG.log
summ() { python - "$1" "$2" <<'PY'
import json, sys
try:
    j = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r = j.get('roofline') or {}
    print(sys.argv[2], 'value', round(j['value'], 1), 'e2e', round(j['e2e']['value'], 1), 'p50', round(j['latency_ms_p50'], 2), 'mhz', j['clocks']['sm_mhz'], 'acc/verify', round(j.get('accepted_tokens_per_verify', 0), 2), 'frac', round(r.get('frac', 0), 3), r.get('bound'),
          {k: round(v['ms_per_user'], 3) for k, v in (j.get('kernel_groups') or {}).items()})
except Exception as e:
    print(sys.argv[2], 'ERR', e)
PY
}
# ---- the driver's command, three times in a row ----
for i in 1 2 3; do
  timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench_default${i}_$TAG.log 2> $O/bench_default${i}_$TAG.err; echo "bench default $i rc=$?"; summ $O/bench_default${i}_$TAG.log default$i
done
timeout 300 python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 > $O/bench_ref_$TAG.log 2> $O/bench_ref_$TAG.err; echo "reference arm rc=$?"; tail -c 400 $O/bench_ref_$TAG.log
# ---- ncu (after the plain command exited 0) ----
timeout 300 python tools/one_user.py --cohort 8 --users 8 > $O/plain_c_$TAG.log 2>&1; echo "plain rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum,launch__grid_size --clock-control none -c 4000 --csv --log-file $O/launches_cohort_$TAG.csv \
    python tools/one_user.py --cohort 8 --users 8 > $O/ncu_list_c_$TAG.log 2>&1; echo "ncu list rc=$?"
python tools/ncu_summary.py $O/launches_cohort_$TAG.csv > $O/launches_cohort_$TAG.txt 2>&1; head -24 $O/launches_cohort_$TAG.txt
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:gemm_wx" -s 900 -c 24 -f -o /tmp/prof_gemm_cohort \
    python tools/one_user.py --cohort 8 --users 8 > $O/ncu_gemm_cohort_$TAG.log 2>&1; echo "ncu cohort gemm rc=$?"
python tools/ncu_launch_table.py /tmp/prof_gemm_cohort.ncu-rep "cohort forward (tools/one_user.py --cohort 8 --users 8), GEMM launches 900..923 of the second pass" > $O/ncu_gemm_cohort_table_$TAG.txt 2>&1
python tools/ncu_traffic.py $O/ncu_traffic_$TAG.json cohort /tmp/prof_gemm_cohort.ncu-rep > $O/ncu_traffic_$TAG.txt 2>&1; cat $O/ncu_traffic_$TAG.txt
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:tree_attention|residual_rmsnorm|qkv_rope|silu_mul" -s 1500 -c 16 -f -o /tmp/prof_rowwise_cohort \
    python tools/one_user.py --cohort 8 --users 8 > $O/ncu_rowwise_cohort_$TAG.log 2>&1; echo "ncu cohort rowwise rc=$?"
python tools/ncu_launch_table.py /tmp/prof_rowwise_cohort.ncu-rep "cohort forward: tree attention and row-wise kernels, launches 1500..1515" > $O/ncu_rowwise_cohort_table_$TAG.txt 2>&1; cut -c1-220 $O/ncu_rowwise_cohort_table_$TAG.txt | head -8
# ---- stand-alone benches ----
timeout 300 python tools/gemm_bench.py > $O/gemm_bench_$TAG.txt 2>&1; tail -36 $O/gemm_bench_$TAG.txt
timeout 120 python tools/att_bench.py > $O/att_bench_$TAG.txt 2>&1; cat $O/att_bench_$TAG.txt
timeout 120 python tools/rowwise_bench.py > $O/rowwise_bench_$TAG.txt 2>&1; tail -9 $O/rowwise_bench_$TAG.txt
# ---- the other configurations ----
timeout 600 python bench.py --steps 10 --warmup 3 --do-sample --dataset games --K 20 --constraint positional --no-cpu-baseline --hf-baseline-users 0 \
    > $O/bench_relaxed_$TAG.log 2> $O/bench_relaxed_$TAG.err; echo "relaxed rc=$?"; summ $O/bench_relaxed_$TAG.log relaxed
timeout 600 python bench.py --steps 10 --warmup 3 --dataset games --K 20 --no-cpu-baseline --hf-baseline-users 0 --check-users 0 \
    > $O/bench_games20_$TAG.log 2> $O/bench_games20_$TAG.err; echo "games20 rc=$?"; summ $O/bench_games20_$TAG.log games20
timeout 600 python bench.py --steps 6 --warmup 3 --draft corr24 --no-cpu-baseline --hf-baseline-users 0 --check-users 0 \
    > $O/bench_corr24_$TAG.log 2> $O/bench_corr24_$TAG.err; echo "corr24 rc=$?"; summ $O/bench_corr24_$TAG.log corr24
timeout 300 python bench.py --gpus 1 --steps 6 --warmup 3 --cohort 1 --lanes 1 --no-cpu-baseline --hf-baseline-users 0 --check-users 0 > $O/bench_single_$TAG.log 2> $O/bench_single_$TAG.err; echo "single rc=$?"; summ $O/bench_single_$TAG.log single
# ---- configs[4]: K x gamma sweep ----
: > $O/sweep_$TAG.jsonl
for K in 1 5 10 20; do for G in 2 3 4; do
  timeout 300 python bench.py --K $K --N 40 --gamma $G --steps 4 --warmup 3 --no-cpu-baseline --hf-baseline-users 0 --check-users 0 \
      2> $O/sweep_${TAG}_K${K}_g${G}.err | tail -1 >> $O/sweep_$TAG.jsonl; echo "K=$K gamma=$G rc=${PIPESTATUS[0]}"
done; done
python - <<PY > $O/sweep_$TAG.txt
import json
for l in open('$O/sweep_$TAG.jsonl'):
    try: j = json.loads(l)
    except Exception: continue
    w = j['config']['workload']; k = w.split(' K=')[1].split(' ')[0]; g = w.split('gamma=')[1].split(' ')[0]
    r = j.get('roofline') or {}; kg = j.get('kernel_groups') or {}
    print(f"K={k:>2} gamma={g} users/s={j['value']:7.1f} e2e={j['e2e']['value']:7.1f} p50={j['latency_ms_p50']:6.2f} ms acc/verify={j['accepted_tokens_per_verify']:5.2f} "
          f"gemm {r.get('bound')} frac={r.get('frac', 0):.3f} topk_GB/s={(kg.get('topk') or {}).get('algorithmic_gbs') or 0:.0f} launches/user={sum(v['launches_per_user'] for v in kg.values()):.0f} sm_mhz={j['clocks']['sm_mhz']}")
PY
cat $O/sweep_$TAG.txt
rm -f $O/sweep_${TAG}_K*.err
du -sh $O
