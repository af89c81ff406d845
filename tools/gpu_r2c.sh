#!/bin/bash
# round 2, GPU call C (8 GPUs): NCCL parity tests, strong-scaling breakdown of config 4 (communication class timed)
set -u
mkdir -p gpurun_out
N=${1:-8}
timeout 300 python -m pytest tests/test_multirank_gloo.py -m gpu -q --timeout 200 > gpurun_out/r2c_nccl_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r2c_nccl_tests.log
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 --scaling strong --no-e2e --no-cpu-baseline > gpurun_out/r2c_bench_strong_n$N.json 2> gpurun_out/r2c_bench_strong_n$N.err
echo "rc=$?" >> gpurun_out/r2c_bench_strong_n$N.err
tail -3 gpurun_out/r2c_nccl_tests.log; head -c 400 gpurun_out/r2c_bench_strong_n$N.json; tail -2 gpurun_out/r2c_bench_strong_n$N.err
