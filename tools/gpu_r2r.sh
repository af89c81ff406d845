#!/bin/bash
# round 2, GPU call R (1 GPU): the reference's own programs (oracle/_ref, CPU binaries on the PETSc model) with glue/nsb200.c loaded
# through -dll_append and linked to the CUDA library: -ns_type b200 on the B200 against -ns_type cnlinear of the same program
set -u
mkdir -p gpurun_out
timeout 100 python -m pytest tests/test_reference_app.py -m gpu -q -x > gpurun_out/r2r_reference_app.log 2>&1
echo "rc=$?" >> gpurun_out/r2r_reference_app.log; tail -5 gpurun_out/r2r_reference_app.log
