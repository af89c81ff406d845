#!/bin/bash
# round 2, GPU call I (1 GPU): N=1 lines of the final build (same step ranges as the N=8 run, and the driver's 20/5), IBM sphere check
set -u
mkdir -p gpurun_out
timeout 200 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-parity > gpurun_out/r2i_bench_n1_10steps.json 2> gpurun_out/r2i_bench_n1_10steps.err
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/r2i_bench_n1.json 2> gpurun_out/r2i_bench_n1.err
echo "rc=$?" >> gpurun_out/r2i_bench_n1.err
timeout 200 python tools/ibm_sphere_validation.py --n 256 --time 30 --budget 110 --out gpurun_out/r2i_ibm_sphere_re300_256.json > gpurun_out/r2i_ibm_sphere.log 2>&1
head -c 330 gpurun_out/r2i_bench_n1_10steps.json; echo; head -c 330 gpurun_out/r2i_bench_n1.json; echo; tail -2 gpurun_out/r2i_bench_n1.err; tail -1 gpurun_out/r2i_ibm_sphere.log | cut -c1-600
