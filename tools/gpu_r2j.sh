#!/bin/bash
# round 2, GPU call J (4 GPUs): strong scaling point N=4 of config 4 (same steps as the N=1 / N=8 lines)
set -u
mkdir -p gpurun_out
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 4 --steps 10 --warmup 3 --scaling strong --no-e2e --no-cpu-baseline > gpurun_out/r2j_bench_n4.json 2> gpurun_out/r2j_bench_n4.err
echo "rc=$?" >> gpurun_out/r2j_bench_n4.err
head -c 300 gpurun_out/r2j_bench_n4.json; echo; grep "ms per step\|rc=" gpurun_out/r2j_bench_n4.err | tail -3
