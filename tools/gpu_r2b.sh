#!/bin/bash
# round 2, GPU call B: full GPU suite, FD bench (fast path), bench N=1 after the fusions / side CTAs, launch list, ncu full captures
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 --durations=8 > gpurun_out/r2b_gpu_suite.log 2>&1
echo "suite rc=$?" >> gpurun_out/r2b_gpu_suite.log
timeout 200 python tools/fd_bench.py --n 512 --reps 20 > gpurun_out/r2b_fd_bench.json 2> gpurun_out/r2b_fd_bench.err
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2b_bench_sphere512.json 2> gpurun_out/r2b_bench_sphere512.err
echo "bench rc=$?" >> gpurun_out/r2b_bench_sphere512.err
FLUCA_B200_NO_LAZY_BASIS=1 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-parity > gpurun_out/r2b_bench_sphere512_nolazy.json 2> gpurun_out/r2b_bench_sphere512_nolazy.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2b_launches_sphere256.csv \
  python bench.py --n 256 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-parity > gpurun_out/r2b_ncu_list.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:'AApplyTile|FaceStarRhs|ProjectAll|DivCell|GradCells|PoissonTile|MGSmoothTile|MGFirstTwoTile|MGResidTile|MGProlong' -c 40 -o gpurun_out/r2b_prof_256 \
  python bench.py --n 256 --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --no-parity > gpurun_out/r2b_ncu_full.log 2>&1
timeout 400 ncu --set full --clock-control none -k regex:'AApplyTile' -c 2 -o gpurun_out/r2b_prof_512_aapply \
  python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --no-parity > gpurun_out/r2b_ncu_full512.log 2>&1
tail -4 gpurun_out/r2b_gpu_suite.log; head -c 600 gpurun_out/r2b_bench_sphere512.json; echo; tail -3 gpurun_out/r2b_bench_sphere512.err; cat gpurun_out/r2b_fd_bench.json; ls -la gpurun_out/*.ncu-rep
