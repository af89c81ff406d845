#!/bin/bash
# round 2, GPU call K (1 GPU): final build -- full GPU suite, FD bench, N=1 bench as the driver runs it
set -u
mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/r2k_gpu_suite.log 2>&1
echo "suite rc=$?" >> gpurun_out/r2k_gpu_suite.log
timeout 100 python tools/fd_bench.py --n 512 --reps 20 > gpurun_out/r2k_fd_bench.json 2> gpurun_out/r2k_fd_bench.err
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/r2k_bench_n1.json 2> gpurun_out/r2k_bench_n1.err
echo "rc=$?" >> gpurun_out/r2k_bench_n1.err
timeout 120 python bench.py --workload cavity --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-parity > gpurun_out/r2k_bench_cavity256.json 2> gpurun_out/r2k_bench_cavity256.err
tail -3 gpurun_out/r2k_gpu_suite.log; cat gpurun_out/r2k_fd_bench.json; head -c 330 gpurun_out/r2k_bench_n1.json; echo; tail -2 gpurun_out/r2k_bench_n1.err; head -c 330 gpurun_out/r2k_bench_cavity256.json
