#!/bin/bash
# round 2, GPU call D (1 GPU): full GPU suite, kernel A/B (occupancy of the streaming kernels, planes per barrier), bench N=1, FD, ncu
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 --durations=6 > gpurun_out/r2d_gpu_suite.log 2>&1
echo "suite rc=$?" >> gpurun_out/r2d_gpu_suite.log
K=momentum_apply,poisson_apply,mg_vcycle,face_star_rhs,project_all,div_cell,coupled_abf_output,momentum_rhs
timeout 200 python tools/kernel_bench.py --n 512 --reps 10 --kernels $K --tag default_minb3_planes2 > gpurun_out/r2d_kernels.jsonl 2> gpurun_out/r2d_kernels.err
FLUCA_B200_BOX_MINB=2 timeout 200 python tools/kernel_bench.py --n 512 --reps 10 --kernels face_star_rhs,project_all,div_cell,coupled_abf_output --tag minb2 >> gpurun_out/r2d_kernels.jsonl 2>> gpurun_out/r2d_kernels.err
FLUCA_B200_BOX_MINB=4 timeout 200 python tools/kernel_bench.py --n 512 --reps 10 --kernels face_star_rhs,project_all,div_cell,coupled_abf_output --tag minb4 >> gpurun_out/r2d_kernels.jsonl 2>> gpurun_out/r2d_kernels.err
timeout 200 python tools/kernel_bench.py --n 512 --reps 10 --kernels poisson_apply,mg_vcycle --tag planes1 --lib fluca_b200/csrc/variants/libfluca_b200_p1.so >> gpurun_out/r2d_kernels.jsonl 2>> gpurun_out/r2d_kernels.err
timeout 100 python tools/fd_bench.py --n 512 --reps 20 > gpurun_out/r2d_fd_bench.json 2> gpurun_out/r2d_fd_bench.err
timeout 100 python tools/fd_bench.py --n 256 --reps 20 --assembled > gpurun_out/r2d_fd_bench_assembled.json 2>> gpurun_out/r2d_fd_bench.err
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2d_bench_sphere512.json 2> gpurun_out/r2d_bench_sphere512.err
echo "bench rc=$?" >> gpurun_out/r2d_bench_sphere512.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2d_launches_sphere256.csv \
  python bench.py --n 256 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-parity > gpurun_out/r2d_ncu_list.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'AApplyTile<2>|FaceStarRhs|ProjectAll|PoissonTile|MGSmoothTile<0>|MGFirstTwoTile|MGResidTile|MGProlong' -c 24 -o gpurun_out/r2d_prof_256 \
  python bench.py --n 256 --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --no-parity > gpurun_out/r2d_ncu_full.log 2>&1
timeout 300 ncu --set full --clock-control none --kernel-name-base demangled -k regex:'AApplyTile<2>' -c 1 -o gpurun_out/r2d_prof_512_aapply \
  python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --no-parity > gpurun_out/r2d_ncu_full512.log 2>&1
tail -4 gpurun_out/r2d_gpu_suite.log; cat gpurun_out/r2d_kernels.jsonl | cut -c1-200; head -c 300 gpurun_out/r2d_bench_sphere512.json; echo; tail -2 gpurun_out/r2d_bench_sphere512.err; cat gpurun_out/r2d_fd_bench.json gpurun_out/r2d_fd_bench_assembled.json; ls -la gpurun_out/*.ncu-rep
