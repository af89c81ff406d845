"""Shared parity harness: drive the C ABI (through the reference-style NS API of fluca_b200.ns)
and the CPU oracle on the same seeded case and compare.  Used by the -m gpu tests with the CUDA
product library and by the CPU tests of the host logic with the host-emulation test double."""
from __future__ import annotations

import os
import subprocess

import numpy as np

import fluca_b200 as fb
from oracle import oracle as O
from tests import cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# FLUCA_B200_HOSTEMU_LIB: another build of the test double, e.g. the AddressSanitizer one of `make -C tests/hostemu asan`
HOSTEMU = os.environ.get("FLUCA_B200_HOSTEMU_LIB") or os.path.join(ROOT, "tests", "hostemu", "_build", "libfluca_b200_hostemu.so")


def hostemu_library():
    subprocess.run(["make", "-C", os.path.join(ROOT, "tests", "hostemu")], check=True, stdout=subprocess.DEVNULL)
    return fb._lib.load(HOSTEMU)


def rel(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-300))


def relU(Ua, Ub):
    num = np.sqrt(sum(np.sum((a - b) ** 2) for a, b in zip(Ua, Ub)))
    den = np.sqrt(sum(np.sum(b**2) for b in Ub))
    return float(num / max(den, 1e-300))


from fluca_b200.workloads import make_ns, set_initial  # noqa: E402,F401  (product-side: bench.py uses them without the oracle)


TIGHT = {"ns_ksp_rtol": 1e-13, "ns_abf_momentum_ksp_rtol": 1e-13, "ns_abf_schur_ksp_rtol": 1e-13, "ns_ksp_max_it": 60}
ORC_TIGHT = dict(outer_rtol=1e-13, mom_rtol=1e-13, schur_rtol=1e-13)


AINV_OPTION = {0: "ID", 1: "DIAG", 2: "ROWSUM"}  # PCABFAinvTypes[], abfpc.c:4


def compare_steps(case, library=None, mode="coupled", nsteps=2, seed=None, tol=1e-10, ptol=None, markers=None, ainv=(0, 0), state=None, fast_oracle=False, ilu_blocks=None):
    """K steps on both sides at tight tolerances; returns the per-step relative L2 differences.
    ainv = (schur, upper) PCABFAinvType of both sides (-ns_pc_abf_{schur,upper}_ainv_type)."""
    orc = cases.make_oracle_fast(case) if fast_oracle else cases.make_oracle(case)
    if state is None:
        state = case.initial_state(seed=seed)
    orc.set_state(*state)
    ns = make_ns(case, library, mode, ns_pc_abf_schur_ainv_type=AINV_OPTION[ainv[0]], ns_pc_abf_upper_ainv_type=AINV_OPTION[ainv[1]], **TIGHT)
    set_initial(ns, state)
    if markers is not None:  # immersed boundary: the same marker list on both sides
        orc.set_markers(markers["X"], markers["Ud"], markers["dV"], markers.get("npts", 4), markers.get("iterations", 1))
        fb.NSB200SetMarkers(ns, markers["X"], markers["Ud"], markers["dV"], markers.get("npts", 4), markers.get("iterations", 1))
    out = []
    oopt = O.default_options(mode=0 if mode == "coupled" else 1, schur_ainv=ainv[0], upper_ainv=ainv[1], **ORC_TIGHT)
    if ilu_blocks:  # block-Jacobi ILU(0) of the inner solves on that many threads: tight inner tolerances make the result independent of it
        oopt.ilu_blocks = int(ilu_blocks)
    for _ in range(nsteps):
        oi = orc.step(oopt)
        fb.NSStep(ns)
        st = fb.NSB200GetStats(ns)
        a, b = orc.get_state(), fb.NSB200GetSolver(ns).get_state()
        ev, eU, ep, eh = rel(b["v"], a["v"]), relU(b["U"], a["U"]), rel(b["p"], a["p"]), rel(b["phalf"], a["phalf"])
        out.append(dict(v=ev, U=eU, p=ep, phalf=eh, outer=(st.outer_its, oi.outer_its), hist_gpu=[st.hist[i] for i in range(st.nhist)], hist_orc=[oi.hist[i] for i in range(oi.nhist)], mom=st.mom_its, schur=st.schur_its))
        if markers is not None:
            (Fo, Uo), (Fg, Ug) = orc.marker_forces(), fb.NSB200GetMarkerForces(ns)
            out[-1]["F"], out[-1]["Um"] = rel(Fg, Fo), rel(Ug, Uo)
            assert out[-1]["F"] <= 100 * tol and out[-1]["Um"] <= tol, (case.name, mode, out[-1])
        assert ev <= tol and eU <= tol, (case.name, mode, out[-1])
        assert ep <= (ptol or 10 * tol) and eh <= (ptol or 10 * tol), (case.name, mode, out[-1])
    fb.NSDestroy(ns)
    return out


def assert_histories_track(out, n=6):
    """Outer (KSP) residual histories of both sides agree: with exact inner solves they depend only on (M, M~, b)."""
    import pytest

    for o in out:
        assert abs(o["outer"][0] - o["outer"][1]) <= 1, o["outer"]
        m = min(len(o["hist_gpu"]), len(o["hist_orc"]), n)
        for a, b in zip(o["hist_gpu"][:m], o["hist_orc"][:m]):
            assert a == pytest.approx(b, rel=1e-6, abs=1e-12 * o["hist_orc"][0])


def coupled_matrix(orc):
    """The 3 x 3 block operator of the step from the oracle's assembled blocks (MatNest J, nsbasic.c:203-207)."""
    import scipy.sparse as sp

    A, G, negT, negR, D = (orc.matrix(k) for k in ("A", "G", "negT", "negR", "D"))
    return sp.bmat([[A, None, G], [negT, sp.identity(negT.shape[0]), negR], [None, D, None]], format="csr"), A.shape[0], negT.shape[0]


def default_tolerance_check(case, library, state, markers=None, nsteps=3, field_tol=1e-4, orc_steps=0):
    """The product at the reference's DEFAULT tolerances (1e-5, inner solves relaxed by the inexact-Krylov rule):
      * step 0: the residual it reports is the TRUE residual of the coupled system, formed here with the oracle's assembled
        operators (no code shared with the solver) and the right-hand side of the step: |b - M x| <= 1e-5 |b|;
      * after nsteps the fields are within field_tol of the product's own answer at tight tolerances (which the tight
        parity tests pin to the oracle's), and -- orc_steps > 0 -- of the oracle run at the reference defaults."""
    import pytest

    has_outlet = any(bc["type"] == cases.BC_PRESSURE_OUTLET for bc in case.bcs)
    orc = cases.make_oracle_fast(case)
    orc.set_state(*state)
    orc.prepare_step()  # assembles A of step 0
    M, nv, nU = coupled_matrix(orc)
    runs = {}
    for label, opts in (("default", {}), ("tight", TIGHT)):
        ns = make_ns(case, library, "coupled", **opts)
        set_initial(ns, state)
        s = fb.NSB200GetSolver(ns)
        if markers is not None:
            fb.NSB200SetMarkers(ns, markers["X"], markers["Ud"], markers["dV"], markers.get("npts", 4))
        info = []
        for k in range(nsteps):
            fb.NSStep(ns)
            st = fb.NSB200GetStats(ns)
            info.append((st.outer_its, st.mom_its, st.schur_its))
            if k == 0 and label == "default":
                got = s.get_state()
                rm, ri, rc = s.get_rhs()  # the b of the step, immersed-boundary forcing included
                b = np.concatenate([rm.ravel()] + [u.ravel() for u in ri] + [rc.ravel()])
                dp = got["phalf"] - state[2]  # step 0: phalf = p0 + p'
                x = np.concatenate([got["v"].ravel()] + [u.ravel() for u in got["U"]] + [dp.ravel()])
                r = b - M @ x
                if not has_outlet:  # the null space of J is projected out of the residual (nsbasic.c:133-144)
                    r[nv + nU :] -= r[nv + nU :].mean()
                true_rel = float(np.linalg.norm(r) / np.linalg.norm(b))
                assert st.converged and true_rel <= 1e-5 * 1.001, (case.name, true_rel)
                assert true_rel == pytest.approx(st.outer_rnorm / st.outer_rnorm0, rel=5e-2), (true_rel, st.outer_rnorm / st.outer_rnorm0)
        runs[label] = (s.get_state(), info)
        fb.NSDestroy(ns)
    a, b = runs["default"][0], runs["tight"][0]
    err = dict(v=rel(a["v"], b["v"]), U=relU(a["U"], b["U"]), p=rel(a["p"], b["p"]))
    assert err["v"] <= field_tol and err["U"] <= field_tol and err["p"] <= 50 * field_tol, (case.name, err, runs["default"][1], runs["tight"][1])
    out = dict(true_rel=true_rel, err=err, its_default=runs["default"][1], its_tight=runs["tight"][1])
    if orc_steps:
        if markers is not None:
            orc.set_markers(markers["X"], markers["Ud"], markers["dV"], markers.get("npts", 4))
        orc.set_state(*state)
        for _ in range(orc_steps):
            orc.step(O.default_options(mode=0))
        oa = orc.get_state()
        # both are 1e-5-class answers of the same step: they differ by the solve tolerances, not more
        if orc_steps == nsteps:
            out["err_oracle_default"] = dict(v=rel(a["v"], oa["v"]), U=relU(a["U"], oa["U"]))
            assert out["err_oracle_default"]["v"] <= 2 * field_tol and out["err_oracle_default"]["U"] <= 2 * field_tol, out
    return out
