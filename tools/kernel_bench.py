"""Times single kernels / kernel groups of the step on device-resident scratch fields (fluca_b200_time_kernel: CUDA events on the
solver stream, mean of `reps` launches after 2 warm-up launches) and prints one JSON line per kernel with its algorithmic
rate against the measured HBM peak.  For A/B runs of the run-time switches (FLUCA_B200_BOX_MINB, ...).

    python tools/kernel_bench.py --n 512 --reps 10 [--kernels momentum_apply,project_all,...]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

ALL = ["momentum_apply", "poisson_apply", "mg_vcycle", "face_star_rhs", "project_all", "div_cell", "coupled_abf_output", "momentum_rhs"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=512)
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--kernels", default=",".join(ALL))
    ap.add_argument("--tag", default="")
    ap.add_argument("--lib", default=None, help="a variant of the product library (tools/build_variant.sh) for A/B runs")
    a = ap.parse_args()
    import fluca_b200 as fb
    from fluca_b200 import workloads as W

    lib = fb._lib.load(a.lib)
    case = W.sphere_bench_case(a.n, a.n)
    ns = W.make_ns(case, lib, "fractional")
    s = fb.NSB200GetSolver(ns)
    v, U, p = W.uniform_inflow_state(case)
    s.set_state(v=v, U=U, p=p, phalf=p)
    del v, U, p
    fb.NSStep(ns)  # a developed-enough state; boundary planes uploaded
    s.prepare_step(ns.t, ns.step)
    peak = 6650.0
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", peak))
    except Exception:
        pass
    for k in a.kernels.split(","):
        ms, by = s.time_kernel(k, a.reps)
        print(json.dumps({"kernel": k, "n": a.n, "tag": a.tag, "avg_ms": ms, "algorithmic_GB": by / 1e9, "achieved_GBs": by / (ms * 1e-3) / 1e9, "frac": by / (ms * 1e-3) / 1e9 / peak, "env": {e: os.environ[e] for e in os.environ if e.startswith("FLUCA_B200_")}}), flush=True)
    fb.NSDestroy(ns)


if __name__ == "__main__":
    main()
