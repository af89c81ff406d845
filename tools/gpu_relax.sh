#!/bin/bash
# A/B of the inner-tolerance relaxation constant on both bench workloads
for c in ${RELAX_LIST:-0.1 1}; do
  for w in cavity sphere; do
    FLUCA_B200_RELAX=$c timeout 300 python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/relax_${c}_$w.json 2>/dev/null
    python -c "
import json; d=json.load(open('gpurun_out/relax_${c}_$w.json')); print('relax $c $w', round(d['value'],1), round(d['ms_per_step'],1), d['config']['iterations_per_step'])"
  done
done
