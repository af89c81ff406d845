#!/bin/bash
# Pre-flight without a GPU: the bodies of the -m gpu parity tests on the host-emulation build (tools/hostemu_pytest_plugin.py says
# what that does and does not prove).  Serial CPU code on the GPU tests' sizes: about ten minutes.
#   tools/gpu_suite_on_hostemu.sh [pytest args]      default: the solver-numerics files
set -eu
cd "$(dirname "$0")/.."
make -C tests/hostemu > /dev/null
FILES="tests/test_bench_workload_parity.py tests/test_gpu_parity.py tests/test_ibm.py tests/test_abf_ainv.py tests/test_golden_ns.py tests/test_oracle_vs_reference.py tests/test_state_view.py tests/test_zz_fd_apply.py"
# the full-size property tests (1024 x 512, 512^3) are for the device: serial host code would need hours and 60 GB
PYTHONPATH=tools python -m pytest -p hostemu_pytest_plugin ${@:-$FILES} -m gpu -q \
  --deselect tests/test_gpu_parity.py::test_config4_sphere512_full_size_properties --deselect tests/test_gpu_parity.py::test_config2_cylinder2d_full_size_properties
