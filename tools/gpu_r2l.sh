#!/bin/bash
# round 2, GPU call L (1 GPU): the committed final build -- full GPU suite (incl. the config-5 code-path cases) and the N=1 line
set -u
mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/r2l_gpu_suite.log 2>&1
echo "suite rc=$?" >> gpurun_out/r2l_gpu_suite.log
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r2l_bench_n1.json 2> gpurun_out/r2l_bench_n1.err
echo "rc=$?" >> gpurun_out/r2l_bench_n1.err
tail -3 gpurun_out/r2l_gpu_suite.log; head -c 330 gpurun_out/r2l_bench_n1.json; echo; tail -2 gpurun_out/r2l_bench_n1.err
