#!/bin/bash
# Round 2, GPU session I: ncu --set full captures of the cohort forward (reduced to CSV / tables ON THE BOX: the .ncu-rep files
# exceed what gpurun copies back), the attention kernel's source-level stall profile, and the bench lines kept in profiles/.
TAG=${1:-r02i}
O=gpurun_out
mkdir -p $O
timeout 300 python tools/one_user.py --cohort 8 --users 8 > $O/plain_c_$TAG.log 2>&1; echo "plain rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:gemm_wx" -s 900 -c 24 -f -o /tmp/prof_gemm_cohort \
    python tools/one_user.py --cohort 8 --users 8 > $O/ncu_gemm_cohort_$TAG.log 2>&1; echo "ncu cohort gemm rc=$?"
python tools/ncu_launch_table.py /tmp/prof_gemm_cohort.ncu-rep "cohort forward (tools/one_user.py --cohort 8 --users 8), GEMM launches 900..923 of the second pass" > $O/ncu_gemm_cohort_table_$TAG.txt 2>&1
python tools/ncu_traffic.py $O/ncu_traffic_$TAG.json cohort /tmp/prof_gemm_cohort.ncu-rep > $O/ncu_traffic_$TAG.txt 2>&1; cat $O/ncu_traffic_$TAG.txt
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:tree_attention|residual_rmsnorm|qkv_rope|silu_mul" -s 1500 -c 16 -f -o /tmp/prof_rowwise_cohort \
    python tools/one_user.py --cohort 8 --users 8 > $O/ncu_rowwise_cohort_$TAG.log 2>&1; echo "ncu cohort rowwise rc=$?"
python tools/ncu_launch_table.py /tmp/prof_rowwise_cohort.ncu-rep "cohort forward: tree attention and row-wise kernels, launches 1500..1515" > $O/ncu_rowwise_cohort_table_$TAG.txt 2>&1; cat $O/ncu_rowwise_cohort_table_$TAG.txt | cut -c1-200
ncu -i /tmp/prof_rowwise_cohort.ncu-rep --page source --csv -k regex:tree_attention -c 1 > $O/att_source_$TAG.csv 2>/dev/null; wc -c $O/att_source_$TAG.csv
ncu -i /tmp/prof_rowwise_cohort.ncu-rep --page details --csv -k regex:tree_attention -c 1 > $O/att_details_$TAG.csv 2>/dev/null; wc -c $O/att_details_$TAG.csv
# ---- bench lines for profiles/ ----
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench_default_$TAG.log 2> $O/bench_default_$TAG.err; echo "bench default rc=$?"
timeout 600 python bench.py --steps 10 --warmup 3 --do-sample --dataset games --K 20 --constraint positional --no-cpu-baseline --hf-baseline-users 0 \
    > $O/bench_relaxed_$TAG.log 2> $O/bench_relaxed_$TAG.err; echo "relaxed rc=$?"
timeout 600 python bench.py --steps 10 --warmup 3 --dataset games --K 20 --no-cpu-baseline --hf-baseline-users 0 --check-users 0 \
    > $O/bench_games20_$TAG.log 2> $O/bench_games20_$TAG.err; echo "games20 rc=$?"
timeout 600 python bench.py --steps 6 --warmup 3 --draft corr24 --no-cpu-baseline --hf-baseline-users 0 --check-users 0 \
    > $O/bench_corr24_$TAG.log 2> $O/bench_corr24_$TAG.err; echo "corr24 rc=$?"
timeout 300 python bench.py --gpus 1 --steps 6 --warmup 3 --cohort 1 --lanes 1 --no-cpu-baseline --hf-baseline-users 0 --check-users 0 > $O/bench_single_$TAG.log 2> $O/bench_single_$TAG.err; echo "single rc=$?"
ls -la $O | tail -25; du -sh $O
