#!/bin/bash
# round 2, GPU call E (1 GPU): full GPU suite, kernel timings, bench N=1 as the driver runs it, ncu captures for profiles/
set -u
mkdir -p gpurun_out
timeout 700 python -m pytest tests -m gpu -q --timeout 400 --durations=5 > gpurun_out/r2e_gpu_suite.log 2>&1
echo "suite rc=$?" >> gpurun_out/r2e_gpu_suite.log
K=momentum_apply,poisson_apply,mg_vcycle,face_star_rhs,project_all,div_cell,coupled_abf_output,momentum_rhs
timeout 200 python tools/kernel_bench.py --n 512 --reps 10 --kernels $K --tag final > gpurun_out/r2e_kernels.jsonl 2> gpurun_out/r2e_kernels.err
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2e_bench_sphere512.json 2> gpurun_out/r2e_bench_sphere512.err
echo "bench rc=$?" >> gpurun_out/r2e_bench_sphere512.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2e_bench_reference.json 2> gpurun_out/r2e_bench_reference.err
timeout 200 python bench.py --workload cavity --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-parity > gpurun_out/r2e_bench_cavity256.json 2> gpurun_out/r2e_bench_cavity256.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2e_launches_sphere256.csv \
  python bench.py --n 256 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-parity > gpurun_out/r2e_ncu_list.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'AApplyTile<\(int\)4>|ProjectAll|FaceStarRhs|MGSmoothTile<\(int\)0>|PoissonTile|momentum_solve' -c 14 -o gpurun_out/r2e_prof_256 \
  python bench.py --n 256 --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --no-parity > gpurun_out/r2e_ncu_full.log 2>&1
timeout 300 ncu --set full --clock-control none --kernel-name-base demangled -k regex:'AApplyTile<\(int\)4>' -c 1 -o gpurun_out/r2e_prof_512_aapply \
  python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --no-parity > gpurun_out/r2e_ncu_full512.log 2>&1
tail -3 gpurun_out/r2e_gpu_suite.log; cut -c1-190 gpurun_out/r2e_kernels.jsonl; head -c 300 gpurun_out/r2e_bench_sphere512.json; echo; tail -2 gpurun_out/r2e_bench_sphere512.err; head -c 300 gpurun_out/r2e_bench_cavity256.json; echo; head -c 400 gpurun_out/r2e_bench_reference.json; echo; ls -la gpurun_out/r2e*.ncu-rep
