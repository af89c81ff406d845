#!/bin/bash
# First GPU call of the next round: everything this round could not verify or measure on a B200 after its GPU budget ran out.
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/next_round_gpu.sh'
# Outputs land in gpurun_out/ (scratch); copy what should be judged into profiles/.
set -u
mkdir -p gpurun_out
# 1. the FlucaFD apply kernel has only run in host emulation
FLUCA_B200_RUN_UNVERIFIED=1 timeout 120 python -m pytest tests/test_zz_fd_apply.py -m gpu -q --timeout 60 > gpurun_out/next_fd_apply.log 2>&1
timeout 200 python tools/fd_bench.py --n 512 --reps 20 > gpurun_out/next_fd_bench.json 2> gpurun_out/next_fd_bench.err
# 2. whole GPU suite (the new files of this round included)
timeout 400 python -m pytest tests -m gpu -x -q --timeout 120 > gpurun_out/next_gpu_suite.log 2>&1
# 3. bench with the resident / overlapped end-to-end loop and the Poisson-solve rate; DIAG and ROWSUM variants at 256^3
timeout 300 python bench.py --steps 3 --warmup 3 > gpurun_out/next_bench_sphere512.json 2> gpurun_out/next_bench_sphere512.err
for v in DIAG ROWSUM; do
  timeout 120 python bench.py --workload cavity --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --schur-ainv $v --upper-ainv $v > gpurun_out/next_bench_cavity256_$v.json 2> gpurun_out/next_bench_cavity256_$v.err
done
# 4. launch list + full capture of the variant Schur kernels (only after the plain run above exited)
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/next_launches_cavity128_diag.csv \
  python bench.py --workload cavity --n 128 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --schur-ainv DIAG --upper-ainv DIAG > gpurun_out/next_ncu_list.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'GradScaleCells|SchurVariantApplyDot' -c 4 -o gpurun_out/next_prof_schur_variant \
  python bench.py --workload cavity --n 256 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --schur-ainv DIAG --upper-ainv DIAG > gpurun_out/next_ncu_full.log 2>&1
tail -3 gpurun_out/next_fd_apply.log gpurun_out/next_gpu_suite.log
