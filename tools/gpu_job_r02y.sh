#!/bin/bash
# Round 2, GPU session Y: evidence for profiles/ with the kernels as committed: ncu launch list of a cohort search, ncu --set full
# of the cohort GEMMs and of the row-wise / attention kernels (reduced to tables on the box), stand-alone kernel benches, and
# the bench lines of BASELINE's configurations.
TAG=${1:-r02y}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv,noheader
timeout 300 python tools/one_user.py --cohort 8 --users 8 > $O/plain_c_$TAG.log 2>&1; echo "plain rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum,launch__grid_size --clock-control none -c 4000 --csv --log-file $O/launches_cohort_$TAG.csv \
    python tools/one_user.py --cohort 8 --users 8 > $O/ncu_list_c_$TAG.log 2>&1; echo "ncu list rc=$?"
python tools/ncu_summary.py $O/launches_cohort_$TAG.csv > $O/launches_cohort_$TAG.txt 2>&1; head -30 $O/launches_cohort_$TAG.txt
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:gemm_wx" -s 900 -c 24 -f -o /tmp/prof_gemm_cohort \
    python tools/one_user.py --cohort 8 --users 8 > $O/ncu_gemm_cohort_$TAG.log 2>&1; echo "ncu cohort gemm rc=$?"
python tools/ncu_launch_table.py /tmp/prof_gemm_cohort.ncu-rep "cohort forward (tools/one_user.py --cohort 8 --users 8), GEMM launches 900..923 of the second pass" > $O/ncu_gemm_cohort_table_$TAG.txt 2>&1
python tools/ncu_traffic.py $O/ncu_traffic_$TAG.json cohort /tmp/prof_gemm_cohort.ncu-rep > $O/ncu_traffic_$TAG.txt 2>&1; cat $O/ncu_traffic_$TAG.txt
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:tree_attention|residual_rmsnorm|qkv_rope|silu_mul" -s 1500 -c 16 -f -o /tmp/prof_rowwise_cohort \
    python tools/one_user.py --cohort 8 --users 8 > $O/ncu_rowwise_cohort_$TAG.log 2>&1; echo "ncu cohort rowwise rc=$?"
python tools/ncu_launch_table.py /tmp/prof_rowwise_cohort.ncu-rep "cohort forward: tree attention and row-wise kernels, launches 1500..1515" > $O/ncu_rowwise_cohort_table_$TAG.txt 2>&1; cut -c1-220 $O/ncu_rowwise_cohort_table_$TAG.txt
# ---- stand-alone benches ----
timeout 300 python tools/gemm_bench.py > $O/gemm_bench_$TAG.txt 2>&1; cat $O/gemm_bench_$TAG.txt
timeout 120 python tools/att_bench.py > $O/att_bench_$TAG.txt 2>&1; cat $O/att_bench_$TAG.txt
timeout 120 python tools/rowwise_bench.py > $O/rowwise_bench_$TAG.txt 2>&1; cat $O/rowwise_bench_$TAG.txt
timeout 300 python tools/kernel_abc_bench.py > $O/abc_bench_$TAG.txt 2>&1; tail -30 $O/abc_bench_$TAG.txt
# ---- bench lines for profiles/ ----
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench_default_$TAG.log 2> $O/bench_default_$TAG.err; echo "bench default rc=$?"
timeout 600 python bench.py --steps 10 --warmup 3 --do-sample --dataset games --K 20 --constraint positional --no-cpu-baseline --hf-baseline-users 0 \
    > $O/bench_relaxed_$TAG.log 2> $O/bench_relaxed_$TAG.err; echo "relaxed rc=$?"
timeout 600 python bench.py --steps 10 --warmup 3 --dataset games --K 20 --no-cpu-baseline --hf-baseline-users 0 --check-users 0 \
    > $O/bench_games20_$TAG.log 2> $O/bench_games20_$TAG.err; echo "games20 rc=$?"
timeout 600 python bench.py --steps 6 --warmup 3 --draft corr24 --no-cpu-baseline --hf-baseline-users 0 --check-users 0 \
    > $O/bench_corr24_$TAG.log 2> $O/bench_corr24_$TAG.err; echo "corr24 rc=$?"
timeout 300 python bench.py --gpus 1 --steps 6 --warmup 3 --cohort 1 --lanes 1 --no-cpu-baseline --hf-baseline-users 0 --check-users 0 > $O/bench_single_$TAG.log 2> $O/bench_single_$TAG.err; echo "single rc=$?"
for f in default relaxed games20 corr24 single; do python - $O/bench_${f}_$TAG.log $f <<'PY'
import json, sys
try:
    j = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r = j.get('roofline') or {}
    print(sys.argv[2], 'value', round(j['value'], 1), 'e2e', round(j['e2e']['value'], 1), 'p50', round(j['latency_ms_p50'], 2), 'mhz', j['clocks']['sm_mhz'], 'acc/verify', round(j.get('accepted_tokens_per_verify', 0), 2), 'frac', round(r.get('frac', 0), 3), r.get('bound'),
          {k: round(v['ms_per_user'], 3) for k, v in (j.get('kernel_groups') or {}).items()})
except Exception as e:
    print(sys.argv[2], 'ERR', e)
PY
done
du -sh $O
