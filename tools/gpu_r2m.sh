#!/bin/bash
# round 2, GPU call M (1 GPU): compute-sanitizer memcheck over smoke() (small cases that reach the tile kernels, the IBM kernels and
# the periodic-x wrap path)
set -u
mkdir -p gpurun_out
timeout 170 compute-sanitizer --tool memcheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2m_sanitizer_memcheck.log 2>&1
echo "rc=$?" >> gpurun_out/r2m_sanitizer_memcheck.log
tail -12 gpurun_out/r2m_sanitizer_memcheck.log
