#!/bin/bash
# tools/gpu_check.sh TAG [pytest -k expr] -- GPU parity tests, then the cavity (256^3) and sphere (512^3) bench lines
TAG=${1:-x}; K=${2:-}
mkdir -p gpurun_out
if [ -n "$K" ]; then timeout 600 python -m pytest tests -m gpu -x -q --timeout 180 -k "$K" 2>&1 | tail -6; else timeout 900 python -m pytest tests -m gpu -x -q --timeout 180 2>&1 | tail -6; fi
timeout 300 python bench.py --workload cavity --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_${TAG}_cavity.json 2> gpurun_out/bench_${TAG}_cavity.err || tail -5 gpurun_out/bench_${TAG}_cavity.err
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_${TAG}_sphere.json 2> gpurun_out/bench_${TAG}_sphere.err || tail -5 gpurun_out/bench_${TAG}_sphere.err
python - <<PY
import json
for w in ("cavity","sphere"):
    try:
        d=json.load(open("gpurun_out/bench_${TAG}_%s.json"%w))
        r=d["roofline"]
        print(w, "value %.1f ms/step %.1f | mom_apply avg_ms %.3f frac %.3f | step frac %.3f"%(d["value"],d["ms_per_step"],r["avg_ms"],r["frac"],d["step_roofline"]["frac"]), d["kernel_shares"], d["config"]["iterations_per_step"])
    except Exception as e:
        print(w, "failed", e)
PY
