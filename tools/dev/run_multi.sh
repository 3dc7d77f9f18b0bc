#!/bin/bash
# on the GPU box (gpurun --gpus N): the whole bench line under torchrun, as the driver launches it
N=$1
mkdir -p gpurun_out
t0=$(date +%s)
timeout 1700 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n$N.jsonl 2> gpurun_out/bench_n$N.err
echo "bench N=$N rc=$? in $(( $(date +%s) - t0 )) s"
tail -c 400 gpurun_out/bench_n$N.jsonl
