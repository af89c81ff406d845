#!/bin/bash
# round 2, GPU call Q (1 GPU): the CUDA library against the fixtures computed by the REFERENCE's own compiled sources
# (tests/golden/ns_reference.npz) and against the oracle on the 3-D outlet cases (operator T's upper-outlet quirk reached the product's
# tables), the operator-level parity on the TMA path, and the bench's own parity case at a small size
set -u
mkdir -p gpurun_out
timeout 70 python -m pytest tests/test_oracle_vs_reference.py tests/test_golden_ns.py -m gpu -q -x > gpurun_out/r2q_reference_fixtures.log 2>&1
echo "rc=$?" >> gpurun_out/r2q_reference_fixtures.log; tail -2 gpurun_out/r2q_reference_fixtures.log
timeout 50 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "channel or operator_level or restarted" > gpurun_out/r2q_parity.log 2>&1
echo "rc=$?" >> gpurun_out/r2q_parity.log; tail -2 gpurun_out/r2q_parity.log
timeout 45 python bench.py --n 128 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2q_bench128.json 2> gpurun_out/r2q_bench128.err
echo "bench rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/r2q_bench128.json').read().strip().splitlines()[-1]); print('parity', d['parity']['ok'], {k:(v['v'],v['p'],v['outer_its_gpu'],v['outer_its_oracle']) for k,v in d['parity']['runs'].items()}, 'value', d['value'])"
