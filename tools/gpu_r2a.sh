#!/bin/bash
# round 2, GPU call A: whole GPU suite (incl. the FlucaFD kernel, never run before), FD bench, bench N=1, launch list
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/r2a_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -x --durations=15 > gpurun_out/r2a_gpu_suite.log 2>&1
echo "suite rc=$?" >> gpurun_out/r2a_gpu_suite.log
timeout 200 python tools/fd_bench.py --n 512 --reps 20 > gpurun_out/r2a_fd_bench.json 2> gpurun_out/r2a_fd_bench.err
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2a_bench_sphere512.json 2> gpurun_out/r2a_bench_sphere512.err
echo "bench rc=$?" >> gpurun_out/r2a_bench_sphere512.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2a_launches_sphere256.csv \
  python bench.py --n 256 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-parity > gpurun_out/r2a_ncu_list.log 2>&1
tail -5 gpurun_out/r2a_gpu_suite.log; head -c 1500 gpurun_out/r2a_bench_sphere512.json; tail -3 gpurun_out/r2a_bench_sphere512.err; cat gpurun_out/r2a_fd_bench.json
