#!/bin/bash
# tools/profile_kernels.sh TAG "regex1 regex2 ..." -- ncu --set full of one launch of each named kernel (256^3 cavity bench)
TAG=${1:-x}; shift
CMD="python bench.py --workload cavity --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
mkdir -p gpurun_out
$CMD > gpurun_out/plain_$TAG.log 2> gpurun_out/plain_$TAG.err || { echo "plain run failed"; tail -5 gpurun_out/plain_$TAG.err; exit 1; }
for K in $1; do
  ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:$K -s ${SKIP:-3} -c 1 -f -o gpurun_out/prof_${K}_$TAG $CMD > gpurun_out/ncu_${K}_$TAG.log 2>&1
done
ls gpurun_out | grep $TAG
