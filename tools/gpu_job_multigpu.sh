#!/bin/bash
# N-GPU bench exactly as the driver launches it (torchrun, NCCL).  bench.py's own watchdog (240 s without progress, 600 s
# overall) ends a stalled run with stack dumps and the partial result; the outer timeout is only a backstop.
# usage: gpurun --gpus N -- 'bash tools/gpu_job_multigpu.sh N [steps]'      (export ATSPEED_GEMM_2CTA=1 to test the pair kernel)
N=${1:-2}
mkdir -p gpurun_out
timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps ${2:-3} --warmup 3 > gpurun_out/bench_n$N.log 2> gpurun_out/bench_n$N.err; echo "bench N=$N rc=$?"
tail -c 600 gpurun_out/bench_n$N.log; grep "^\[bench r\|watchdog\|failed in phase" gpurun_out/bench_n$N.err | tail -40 | cut -c1-200
