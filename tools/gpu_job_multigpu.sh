#!/bin/bash
# N-GPU bench exactly as the driver launches it (torchrun, NCCL), plus the reference arm under torchrun
N=${1:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_n$N.log 2> gpurun_out/bench_n$N.err; echo "bench N=$N rc=$?"
tail -c 1200 gpurun_out/bench_n$N.log; tail -5 gpurun_out/bench_n$N.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus $N --steps 1 --warmup 1 > gpurun_out/bench_ref_n$N.log 2> gpurun_out/bench_ref_n$N.err; echo "ref N=$N rc=$?"
tail -c 600 gpurun_out/bench_ref_n$N.log
