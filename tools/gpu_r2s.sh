#!/bin/bash
# round 2, GPU call S (1 GPU): the pressure PCG with the flexible (Polak-Ribiere) beta from iteration 4 of a solve on -- closed-domain
# cases (cavities, Taylor-Green) through the CUDA library against the oracle and the reference fixtures
set -u
mkdir -p gpurun_out
timeout 80 python -m pytest tests/test_gpu_parity.py tests/test_oracle_vs_reference.py tests/test_golden_ns.py -m gpu -q -x -k "cavity or tgv or config1" > gpurun_out/r2s_pcg.log 2>&1
echo "rc=$?" >> gpurun_out/r2s_pcg.log; tail -3 gpurun_out/r2s_pcg.log
