#!/bin/bash
# tools/profile_final.sh TAG -- launch list + ncu --set full captures on the sphere workload at 256^3 (same code path as the
# 512^3 bench line, 1/8 of the cells so that ncu's replays stay short)
TAG=${1:-x}
CMD="python bench.py --n 256 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
mkdir -p gpurun_out
$CMD > gpurun_out/plain_$TAG.log 2> gpurun_out/plain_$TAG.err || { echo "plain run failed"; tail -5 gpurun_out/plain_$TAG.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -s 1400 -c 2600 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_list_$TAG.log 2>&1
for K in AApplyTile PoissonTile MGSmoothTile MGFirstTwoTile CoupledCells MGResidRestrict; do
  ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:$K -s 2 -c 1 -f -o gpurun_out/prof_${K}_$TAG $CMD > gpurun_out/ncu_${K}_$TAG.log 2>&1
done
ls gpurun_out | grep $TAG
