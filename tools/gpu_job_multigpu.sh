#!/bin/bash
# N-GPU bench exactly as the driver launches it (torchrun, NCCL); a watchdog kills a hung run after 200 s
N=${1:-2}
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps ${2:-3} --warmup 3 > gpurun_out/bench_n$N.log 2> gpurun_out/bench_n$N.err; echo "bench N=$N rc=$?"
tail -c 300 gpurun_out/bench_n$N.log; tail -3 gpurun_out/bench_n$N.err | cut -c1-300
