#!/bin/bash
# Round 2, GPU session B: where does the opt-in CTA-pair GEMM stall (progress trace), do the two CUTLASS-style orderings
# cure it; the tightened / new parity tests; kernel (a) with 8 loads in flight; parity + comparison records of the bench.
TAG=${1:-r02b}
O=gpurun_out
mkdir -p $O
# ---- 1. pair kernel, ONE lane, progress trace: the waits the mbarrier guard cannot bound ----
run_soak() {   # name, env...
  local name=$1; shift
  env "$@" ATSPEED_GEMM_2CTA=1 ATSPEED_GEMM_TRACE=1 timeout 120 python tools/soak.py --lanes 1 --seconds 45 --stall-s 10 \
      > $O/soak_${name}_$TAG.log 2> $O/soak_${name}_$TAG.err
  echo "soak $name rc=$?"; head -c 6000 $O/soak_${name}_$TAG.log; echo; tail -2 $O/soak_${name}_$TAG.err
}
run_soak pair_trace X=1
run_soak pair_pdl_late ATSPEED_PDL_LATE=1
run_soak pair_relinq_late ATSPEED_RELINQ_LATE=1
run_soak pair_both_late ATSPEED_PDL_LATE=1 ATSPEED_RELINQ_LATE=1
run_soak pair_nopdl ATSPEED_PDL=0
# ---- 2. tests written / tightened after session A ----
timeout 900 python -m pytest tests/test_gpu_e2e.py tests/test_gpu_cohort.py tests/test_gpu_cohort_fp32.py tests/test_zz_gpu_shape7b.py \
    tests/test_zz_gpu_sharded_metrics.py -q -x -s > $O/gpu_tests_new_$TAG.log 2>&1; echo "pytest new rc=$?"; grep -E "passed|failed|identical|parity|explained" $O/gpu_tests_new_$TAG.log | tail -20
# ---- 3. kernel (a): 8 independent 128-bit loads in flight per thread ----
ATSPEED_TOPK_UNROLL=8 timeout 300 python tools/kernel_abc_bench.py > $O/abc_bench_unroll8_$TAG.txt 2>&1; echo "abc unroll8 rc=$?"; head -12 $O/abc_bench_unroll8_$TAG.txt
ATSPEED_TOPK_UNROLL=8 timeout 300 python -m pytest tests/test_gpu_kernels.py -q -k "topk or mask" > $O/topk_test_unroll8_$TAG.log 2>&1; echo "topk unroll8 test rc=$?"; tail -2 $O/topk_test_unroll8_$TAG.log
# ---- 4. bench records: oracle parity at the full benchmark shape, AtSpeed-R Games K=20 (configs[2]), correlated draft ----
timeout 900 python bench.py --steps 6 --warmup 3 --check-users 3 > $O/bench_check_$TAG.log 2> $O/bench_check_$TAG.err; echo "bench check rc=$?"
python - <<PY
import json
try:
    j = json.loads(open('$O/bench_check_$TAG.log').read().strip().splitlines()[-1])
    print('value', j['value'], 'e2e', j['e2e']['value'], 'p50', j['latency_ms_p50'])
    print('parity_vs_oracle', j.get('parity_vs_oracle'))
    print('hf', {k: v for k, v in (j.get('hf_gpu_baseline') or {}).items() if k != 'what'})
except Exception as e:
    print('ERR', e)
PY
timeout 600 python bench.py --steps 6 --warmup 3 --do-sample --dataset games --K 20 --constraint positional --no-cpu-baseline --hf-baseline-users 0 \
    > $O/bench_relaxed_$TAG.log 2> $O/bench_relaxed_$TAG.err; echo "bench relaxed rc=$?"; tail -c 600 $O/bench_relaxed_$TAG.log
timeout 600 python bench.py --steps 6 --warmup 3 --draft corr2 --no-cpu-baseline --hf-baseline-users 0 \
    > $O/bench_corr2_$TAG.log 2> $O/bench_corr2_$TAG.err; echo "bench corr2 rc=$?"; tail -c 600 $O/bench_corr2_$TAG.log
