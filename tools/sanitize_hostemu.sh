#!/bin/bash
# Memory / undefined-behaviour check that can run without a GPU (compute-sanitizer is closed on the pool): the solver sources built
# for the host (-DFLUCA_HOSTEMU: every kernel functor runs in a serial loop, same indexing, same host logic) with
# -fsanitize=address,undefined, under (1) the reference-style glue driver on the PETSc model, leak check on, and (2) the
# host-emulation tests of the CPU suite.  Any "runtime error" / "AddressSanitizer" line is a finding.  ~8 minutes.
set -u
cd "$(dirname "$0")/.."
make -C tests/hostemu asan > /dev/null || exit 1
B=$PWD/tests/hostemu/_build
OUT=${1:-/tmp/fluca_b200_sanitize}
mkdir -p "$OUT"
CC=$([ -x /usr/bin/gcc ] && echo /usr/bin/gcc || echo gcc)
$CC -std=gnu11 -O1 -g -fsanitize=address,undefined -fno-omit-frame-pointer -I tests/petsc_stub -I include tests/c/ns_b200_glue_driver.c \
  tests/petsc_stub/petsc_fluca_mock.c glue/nsb200.c -o "$OUT/drv" -L "$B" -l:libfluca_b200_hostemu_asan.so -Wl,-rpath,"$B" -lm -lstdc++ || exit 1
: > "$OUT/log.txt"
for a in "case=cavity2d steps=2" "case=cavity3d steps=2 scenario=restart" "case=tgv_periodic steps=2 -ns_b200_mode=1" \
  "case=channel2d_t pout=0.3 scenario=formfunction" "case=channel3d_pz init=smooth scenario=edit" \
  "case=cavity3d_full n=6,5,4 stretch=0.15 scenario=stage -ns_b200_sync_interval=0" \
  "case=channel3d n=16,8,8 -ns_pc_abf_schur_ainv_type=DIAG -ns_pc_abf_upper_ainv_type=ROWSUM"; do
  echo "== glue driver: $a" >> "$OUT/log.txt"
  ASAN_OPTIONS=detect_leaks=1 UBSAN_OPTIONS=print_stacktrace=1 "$OUT/drv" $a out="$OUT/a.bin" >> "$OUT/log.txt" 2>&1
done
FLUCA_B200_HOSTEMU_LIB=$B/libfluca_b200_hostemu_asan.so LD_PRELOAD=$($CC -print-file-name=libasan.so):$($CC -print-file-name=libubsan.so) \
  ASAN_OPTIONS=detect_leaks=0 UBSAN_OPTIONS=print_stacktrace=1 python -m pytest tests/test_hostlogic.py tests/test_ibm.py tests/test_abf_ainv.py \
  tests/test_state_view.py tests/test_golden_ns.py tests/test_multirank_gloo.py tests/test_bench_workload_hostlogic.py tests/test_zz_fd_apply.py \
  tests/test_fd_stencils.py -q -m "not gpu" -p no:cacheprovider -s >> "$OUT/log.txt" 2>&1
echo "findings: $(grep -cE 'runtime error|AddressSanitizer|LeakSanitizer' "$OUT/log.txt")"
grep -E "runtime error|AddressSanitizer|LeakSanitizer|passed|failed" "$OUT/log.txt" | sort | uniq -c | head -20
