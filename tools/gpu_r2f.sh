#!/bin/bash
# round 2, GPU call F (8 GPUs): the bench as the driver runs it at N=8 (parity case, strong-scaling headline, weak run), then BASELINE
# config 5 (2048 x 1024 x 1024, ~1 M markers, coupled mode with restart 1).  Tight limits: a stuck run must not eat the budget.
set -u
mkdir -p gpurun_out
N=8
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2f_bench_n$N.json 2> gpurun_out/r2f_bench_n$N.err
echo "rc=$?" >> gpurun_out/r2f_bench_n$N.err
timeout 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $N --workload channel --steps 3 --warmup 3 --no-e2e --no-parity --no-cpu-baseline > gpurun_out/r2f_bench_config5_n$N.json 2> gpurun_out/r2f_bench_config5_n$N.err
echo "rc=$?" >> gpurun_out/r2f_bench_config5_n$N.err
head -c 500 gpurun_out/r2f_bench_n$N.json; echo; grep "bench r0" gpurun_out/r2f_bench_n$N.err | tail -12; head -c 500 gpurun_out/r2f_bench_config5_n$N.json; echo; grep "bench r0\|rror\|rc=" gpurun_out/r2f_bench_config5_n$N.err | tail -12
