#!/bin/bash
# tools/profile_gpu.sh TAG -- run on the GPU box (under gpurun): plain bench, ncu launch list, ncu --set full
# captures of the TMA stencil kernels.  Outputs land in gpurun_out/.
TAG=${1:-r1}
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
mkdir -p gpurun_out
$CMD > gpurun_out/plain_$TAG.log 2> gpurun_out/plain_$TAG.err || { echo "plain run failed"; tail -5 gpurun_out/plain_$TAG.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_list_$TAG.log 2>&1
for K in AApplyTile PoissonTile MGSmoothTile; do
  ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:$K -s 6 -c 2 -f -o gpurun_out/prof_${K}_$TAG $CMD > gpurun_out/ncu_${K}_$TAG.log 2>&1
done
ls -la gpurun_out | tail -12
