"""Generates tests/golden/ns_reference.npz: Navier-Stokes fixtures computed by the REFERENCE's own sources.

oracle/_ref/libfluca_ref_ns.so is thecasterian/fluca's cartdiscret.c, cnlinear.c, cnlinearcart2d.c, cnlinearcart3d.c and abfpc.c
compiled from /root/reference (`make -C oracle ref`) on a single-rank model of the PETSc API subset they use (oracle/ref_model/):
every stencil weight, operator, boundary-condition vector, right-hand side, ABF factor and the step's update are the reference's
code; the linear solves, which the reference delegates to PETSc, are exact (dense LU), i.e. the limit its Krylov solvers converge
to.  The reference stores no NS golden output of its own; these are the closest thing to one that can be produced
without PETSc.  The reference tree is not on the GPU box, so the outputs are committed:

  per case   in_v, in_U*, in_p            seeded inputs (fluca_b200/workloads.py)
             rhs                          NSFormFunction at step 0: b = (r_mom, r_int, r_con)
             coupled/{v,U*,p,phalf}       state after two steps, SNESSolve = exact solution of J x = b
             fractional/{v,U*,p,phalf}    state after two steps, SNESSolve = one application of the reference's PCABF
             gmres_hist                   residual history of right-preconditioned GMRES + PCABF on step 0 (exact inner solves)
             abf_in, abf_out              b = J x for a seeded x, and PCABF applied to b (operators of step 0)

    python tests/golden/make_ns_reference_golden.py      # needs /root/reference; rewrites tests/golden/ns_reference.npz
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import oracle as O  # noqa: E402  (only its BC container; nothing of the oracle's arithmetic is used here)
from oracle import ref as R  # noqa: E402
from tests import cases  # noqa: E402

ID, DIAG, ROWSUM = 0, 1, 2  # PCABFAinvType, flucans.h:99-103


def fixtures():
    """name -> (case, seed, (Schur, upper) PCABFAinvType).  Small: the coupled solves are dense."""
    st = cases.cavity2d(n=10)
    st.stretch = 0.2
    c3 = cases.channel3d(n=(8, 6, 5), pout=0.2, dt=0.05)
    c3s = cases.channel3d(n=(8, 6, 5), pout=0.2, dt=0.05)
    c3s.stretch = 0.15
    return {
        "cavity2d_8": (cases.cavity2d(n=8), 3, (ID, ID)),
        "cavity2d_10_stretched": (st, 4, (ID, ID)),
        "tgv_8": (cases.tgv(n=8, dt=0.05), None, (ID, ID)),
        "tgv_periodic_8": (cases.tgv(n=8, periodic=True, dt=0.05), None, (ID, ID)),
        "channel2d_outlet_timedep_12x8": (cases.channel2d(n=(12, 8), pout=0.3, time_dependent=True), 6, (ID, ID)),
        "cavity3d_sym_6x5x4": (cases.cavity3d(n=(6, 5, 4)), 5, (ID, ID)),
        "channel3d_outlet_8x6x5": (c3, 11, (ID, ID)),
        "channel3d_outlet_8x6x5_stretched": (c3s, 12, (ID, ID)),
        "channel3d_periodic_z_8x5x4": (cases.channel3d(n=(8, 5, 4), periodic_z=True, dt=0.05), 13, (ID, ID)),
        "cavity2d_8_abf_diag_rowsum": (cases.cavity2d(n=8), 3, (DIAG, ROWSUM)),
        "channel3d_outlet_8x6x5_abf_rowsum_diag": (c3, 11, (ROWSUM, DIAG)),
    }


def make_reference(case):
    bcs = [O.BC(b["type"], velocity=b["velocity"], pressure=b["pressure"]) for b in case.bcs]
    return R.Reference(case.n, case.faces(), case.rho, case.mu, case.dt, bcs)


def run(case, seed, ainv, nsteps=2):
    state = case.initial_state(seed=seed)
    out = {}
    for mode, tag in ((R.EXACT, "coupled"), (R.ABF_ONCE, "fractional")):
        ref = make_reference(case)
        ref.set_state(*state)
        if tag == "coupled":
            out["rhs"] = ref.form_function()
        for _ in range(nsteps):
            ref.step(mode=mode, schur_ainv=ainv[0], upper_ainv=ainv[1])
        out[tag] = ref.get_state()
    ref = make_reference(case)
    ref.set_state(*state)
    its, hist = ref.step(mode=R.GMRES_ABF, schur_ainv=ainv[0], upper_ainv=ainv[1], rtol=1e-12, maxit=60)
    out["gmres_hist"] = hist
    # a vector in the range of J (b = J x, x seeded): the Schur system of the ABF application is then consistent on closed domains too
    rng = np.random.default_rng(77)
    b = ref.apply_jacobian(rng.standard_normal(ref.nsol))
    out["abf_in"], out["abf_out"] = b, ref.abf_apply(b, ainv[0], ainv[1])
    return state, out


def main():
    assert R.available(), "the reference tree (/root/reference) is needed to build oracle/_ref"
    data = {}
    for name, (case, seed, ainv) in fixtures().items():
        state, out = run(case, seed, ainv)
        data[f"{name}/in_v"], data[f"{name}/in_p"] = state[0], state[2]
        for d, u in enumerate(state[1]):
            data[f"{name}/in_U{d}"] = u
        for k in ("rhs", "gmres_hist", "abf_in", "abf_out"):
            data[f"{name}/{k}"] = out[k]
        for tag in ("coupled", "fractional"):
            st = out[tag]
            data[f"{name}/{tag}/v"], data[f"{name}/{tag}/p"], data[f"{name}/{tag}/phalf"] = st["v"], st["p"], st["phalf"]
            for d, u in enumerate(st["U"]):
                data[f"{name}/{tag}/U{d}"] = u
        print(name, "gmres iterations", len(out["gmres_hist"]) - 1)
    path = os.path.join(ROOT, "tests", "golden", "ns_reference.npz")
    np.savez_compressed(path, **data)
    print(path, os.path.getsize(path), "bytes,", len(data), "arrays")


if __name__ == "__main__":
    main()
