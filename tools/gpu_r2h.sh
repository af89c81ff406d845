#!/bin/bash
# round 2, GPU call H (8 GPUs): the bench exactly as the driver launches it at N=8, final build (graph-replayed V-cycles)
set -u
mkdir -p gpurun_out
N=8
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2h_bench_n$N.json 2> gpurun_out/r2h_bench_n$N.err
echo "rc=$?" >> gpurun_out/r2h_bench_n$N.err
head -c 400 gpurun_out/r2h_bench_n$N.json; echo; grep "bench r0\|rc=" gpurun_out/r2h_bench_n$N.err | tail -8
