#!/bin/bash
# round 2, GPU call P (1 GPU): the glue run against the PETSc model on the CUDA library, the state-view and C-driver tests, and an
# A/B of the resident time loop (bench line key e2e_resident): reduction sums published through mapped pinned memory + one
# cudaMemcpy3DAsync per field, against the former D2H memcpy per reduction + one cudaMemcpy2DAsync per plane
set -u
mkdir -p gpurun_out
timeout 150 python -m pytest tests/test_glue_mock.py tests/test_state_view.py tests/test_c_driver.py -m gpu -q -x > gpurun_out/r2p_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r2p_tests.log
tail -3 gpurun_out/r2p_tests.log
B="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-parity"
timeout 100 $B > gpurun_out/r2p_bench_new.json 2> gpurun_out/r2p_bench_new.err; echo "new rc=$?"
FLUCA_B200_RESULT_MEMCPY=1 FLUCA_B200_COPY_PLANES=1 timeout 100 $B > gpurun_out/r2p_bench_old.json 2> gpurun_out/r2p_bench_old.err; echo "old rc=$?"
FLUCA_B200_RESULT_MEMCPY=1 timeout 100 $B > gpurun_out/r2p_bench_3dcopy_only.json 2> gpurun_out/r2p_bench_3dcopy_only.err; echo "3d-only rc=$?"
python - <<'PY'
import json
for t in ("new", "old", "3dcopy_only"):
    try:
        d = json.loads(open(f"gpurun_out/r2p_bench_{t}.json").read().strip().splitlines()[-1])
        print(t, "value", round(d["value"], 1), "ms", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"], 1), "resident", d.get("e2e_resident"))
    except Exception as e:
        print(t, "failed", e)
PY
