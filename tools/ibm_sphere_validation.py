"""Physical check of the immersed-boundary coupling on the GPU at BASELINE config 4's own geometry:
flow past a sphere at Re = 300 on [-4,12] x [-8,8]^2 with the config's boundary set (inflow, pressure outlet, symmetry), D = 1 at
the origin, markers spaced ~ h on a Fibonacci lattice, direct forcing, product library on cuda:0.  Literature (not
from the reference): C_D ~ 0.65-0.66 at Re = 300 (mean), St ~ 0.134-0.137 (the wake needs t ~ 100+ D/U to settle into shedding;
a short run gives the drag level, not the Strouhal number).

    python tools/ibm_sphere_validation.py --n 256 --time 30 --out gpurun_out/ibm_sphere.json [--library PATH] [--passes 2]
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=256)
    ap.add_argument("--time", type=float, default=30.0)
    ap.add_argument("--out", default="gpurun_out/ibm_sphere.json")
    ap.add_argument("--mode", default="fractional")
    ap.add_argument("--passes", type=int, default=2)
    ap.add_argument("--library", default=None)
    ap.add_argument("--budget", type=float, default=150.0, help="stop after this many seconds of wall time")
    a = ap.parse_args()
    import fluca_b200 as fb
    from fluca_b200 import workloads as W

    lib = fb._lib.load(a.library)
    n, h = a.n, 16.0 / a.n
    case = W.sphere_bench_case(n, n)
    ns = W.make_ns(case, lib, a.mode, ns_ksp_gmres_restart=5)
    s = fb.NSB200GetSolver(ns)
    v, U, p = W.uniform_inflow_state(case)
    s.set_state(v=v, U=U, p=p, phalf=p)
    del v, U, p
    nm = int(np.ceil(np.pi / h**2))  # surface area / h^2: marker spacing ~ h
    mk = W.sphere_markers((0.0, 0.0137, 0.0071), 1.0, nm, h)  # slightly off the grid's symmetry planes
    fb.NSB200SetMarkers(ns, mk["X"], mk["Ud"], mk["dV"], 4, a.passes)
    nsteps = int(round(a.time / case.dt))
    hist, t0 = [], time.time()
    area = np.pi / 4.0
    for k in range(nsteps):
        fb.NSStep(ns)
        if k % 5 == 4 or k == nsteps - 1:
            F, Um = fb.NSB200GetMarkerForces(ns)
            cd, cy, cz = (-F[c].sum() / (0.5 * area) for c in range(3))  # force on the body = - force on the fluid
            hist.append((float((k + 1) * case.dt), float(cd), float(cy), float(cz)))
        if time.time() - t0 > a.budget:
            break
    el = time.time() - t0
    t = np.array([x[0] for x in hist])
    cd = np.array([x[1] for x in hist])
    late = t >= 0.6 * t[-1]
    res = {"n": n, "D_over_h": 1.0 / h, "markers": nm, "mode": a.mode, "forcing_passes": a.passes, "dt": case.dt, "steps": k + 1, "t_end": float(t[-1]), "seconds": el, "ms_per_step": 1e3 * el / (k + 1),
           "CD_mean_last_40pct": float(cd[late].mean()), "CD_min_max_last_40pct": [float(cd[late].min()), float(cd[late].max())], "literature_CD": [0.65, 0.66],
           "side_force_rms_last_40pct": float(np.sqrt(np.mean(np.array([x[2] for x in hist])[late] ** 2 + np.array([x[3] for x in hist])[late] ** 2))), "history_t_CD_Cy_Cz": hist[:: max(1, len(hist) // 200)]}
    os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
    json.dump(res, open(a.out, "w"))
    print(json.dumps({k2: v2 for k2, v2 in res.items() if k2 != "history_t_CD_Cy_Cz"}))
    fb.NSDestroy(ns)


if __name__ == "__main__":
    main()
