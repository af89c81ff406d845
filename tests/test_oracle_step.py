"""Self-consistency of the oracle's NS step (the reference stores no NS golden; the pin against the reference's compiled sources is
tests/test_oracle_vs_reference.py).

What is checked instead:
  * the oracle's Krylov result equals a sparse-direct (SciPy SuperLU) solve of the very same
    assembled coupled system M x = b  (THEORY_GUIDE.md:190-198) for every BC family;
  * S = D((-T)G - (-R)) built by sparse products as abfpc.c:151-170 equals -D*Gst, is symmetric
    with zero row sums on wall-bounded uniform grids;
  * the discrete continuity equation D U = 0 holds after a step;
  * Taylor-Green vortex (taylor_green_vortex.c:13-22): second-order convergence in space/time;
  * one ABF application (Mode B) differs from the converged coupled solve by O(1e-2), as SURVEY F9.
"""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spl

from oracle import oracle as O
from tests import cases

TIGHT = dict(outer_rtol=1e-13, mom_rtol=1e-13, schur_rtol=1e-13)


def direct_solve(o, rhs):
    A, G, negT, negR, D = (o.matrix(k) for k in ("A", "G", "negT", "negR", "D"))
    nv, nF, N = o.dim * o.ncell, sum(o.nface), o.ncell
    M = sp.bmat([[A, None, G], [negT, sp.identity(nF), negR], [None, D, None]], format="csr")
    b = rhs.copy()
    singular = abs(o.matrix("S").sum(axis=1)).max() < 1e-9 * abs(o.matrix("S")).max()
    if singular:
        b[nv + nF :] -= b[nv + nF :].mean()
        e = np.zeros(M.shape[0])
        e[nv + nF :] = 1.0
        Maug = sp.bmat([[M, e[:, None]], [e[None, :], None]], format="csc")
        return spl.spsolve(Maug, np.r_[b, 0.0])[:-1], singular
    return spl.spsolve(M.tocsc(), b), singular


CASES = [
    ("cavity2d", lambda: cases.cavity2d(n=16), None),
    ("cavity2d_nonuniform", lambda: _stretched(cases.cavity2d(n=12)), 3),
    ("cavity3d_sym", lambda: cases.cavity3d(n=(8, 8, 4)), 5),
    ("tgv_dirichlet", lambda: cases.tgv(n=8, dt=0.05), None),
    ("tgv_periodic", lambda: cases.tgv(n=8, periodic=True, dt=0.05), None),
    ("channel2d_outlet", lambda: cases.channel2d(n=(16, 8), pout=0.3, time_dependent=True), 7),
    ("channel3d_outlet", lambda: cases.channel3d(n=(8, 6, 6), pout=0.2, dt=0.05), 11),
    ("channel3d_periodic_z", lambda: cases.channel3d(n=(8, 6, 6), periodic_z=True, dt=0.05), 13),
]


def _stretched(c):
    c.stretch = 0.6
    return c


@pytest.mark.parametrize("name,mk,seed", CASES, ids=[c[0] for c in CASES])
def test_mode_a_equals_sparse_direct(name, mk, seed):
    case = mk()
    o = cases.make_oracle(case)
    v, U, p = case.initial_state(seed=seed)
    o.set_state(v, U, p, step=0, t=0.0)
    for step in range(2):  # step 0 (q = p0) and step 1 (q = phalf) take different branches
        rhs = o.prepare_step()
        st0 = o.get_state()
        xex, singular = direct_solve(o, rhs)
        info = o.step(O.default_options(**TIGHT))
        assert info.converged
        st = o.get_state()
        vex, Uex, dpex = o.split(xex)
        assert np.linalg.norm(st["v"] - vex) <= 1e-10 * max(np.linalg.norm(vex), 1e-30)
        for d in range(o.dim):
            assert np.linalg.norm(st["U"][d] - Uex[d]) <= 1e-10 * max(np.linalg.norm(np.concatenate([u.ravel() for u in Uex])), 1e-30)
        pref = (st0["p"] + 2.0 * dpex) if step == 0 else (st0["phalf"] + 1.5 * dpex)
        assert np.linalg.norm(st["p"] - pref) <= 1e-9 * max(np.linalg.norm(pref), 1e-30)
        # continuity
        D = o.matrix("D")
        Ucat = np.concatenate([u.ravel() for u in st["U"]])
        assert abs(D @ Ucat).max() <= 1e-9 * max(abs(Ucat).max(), 1.0) * abs(D).max()


def test_schur_complement_structure():
    case = cases.cavity2d(n=16)
    o = cases.make_oracle(case)
    o.set_state(*case.initial_state(seed=1))
    o.prepare_step()
    S, D, Gst = o.matrix("S"), o.matrix("D"), o.matrix("Gst")
    scale = abs(S).max()
    assert abs(S + D @ Gst).max() <= 1e-14 * scale  # wide-stencil terms cancel
    assert abs(S - S.T).max() <= 1e-13 * scale
    assert abs(S.sum(axis=1)).max() <= 1e-12 * scale
    # the product form carries the cancelled +-2 entries as stored zeros: 13-point pattern in 2-D interior is 9
    assert S.nnz > (-(D @ Gst)).nnz


def test_mode_b_is_the_fractional_step_and_differs_from_coupled_solve():
    case = cases.cavity2d(n=32, dt=0.01)
    oa, ob = cases.make_oracle(case), cases.make_oracle(case)
    z = case.initial_state()
    oa.set_state(*z)
    ob.set_state(*z)
    ia = oa.step(O.default_options(mode=0, **TIGHT))
    ib = ob.step(O.default_options(mode=1, **TIGHT))
    assert ib.abf_applies == 1 and ia.abf_applies > 5
    va, vb = oa.get_state()["v"], ob.get_state()["v"]
    rel = np.linalg.norm(va - vb) / np.linalg.norm(va)
    assert 1e-3 < rel < 6e-2


def test_default_tolerances_give_1e5_class_answer():
    case = cases.cavity2d(n=32, dt=0.01)
    oa, ob = cases.make_oracle(case), cases.make_oracle(case)
    z = case.initial_state()
    oa.set_state(*z)
    ob.set_state(*z)
    oa.step(O.default_options(**TIGHT))
    info = ob.step(O.default_options())  # reference defaults: 1e-5 everywhere
    assert info.converged and info.outer_its <= 8
    va, vb = oa.get_state()["v"], ob.get_state()["v"]
    assert np.linalg.norm(va - vb) / np.linalg.norm(va) < 1e-4


def _tgv_error(n, nsteps, periodic):
    t_final = 0.2
    case = cases.tgv(n=n, periodic=periodic, mu=1.0, dt=t_final / nsteps)
    o = cases.make_oracle(case)
    o.set_state(*case.initial_state())
    for _ in range(nsteps):
        assert o.step(O.default_options(**TIGHT)).converged
    st = o.get_state()
    xc = case.centres()
    Y, X = np.meshgrid(xc[1], xc[0], indexing="ij")
    e = np.exp(-2.0 * t_final)
    uex = np.sin(X) * np.cos(Y) * e
    vex = -np.cos(X) * np.sin(Y) * e
    err = np.sqrt(np.mean((st["v"][0][0] - uex) ** 2 + (st["v"][1][0] - vex) ** 2))
    return err


@pytest.mark.parametrize("periodic", [False, True], ids=["dirichlet", "periodic"])
def test_taylor_green_second_order(periodic):
    e1 = _tgv_error(8, 4, periodic)
    e2 = _tgv_error(16, 8, periodic)
    e3 = _tgv_error(32, 16, periodic)
    assert e2 < e1 and e3 < e2
    order = np.log2(e2 / e3)
    assert order > 1.7, (e1, e2, e3)


def test_block_jacobi_ilu_matches_serial_answer():
    case = cases.cavity3d(n=(8, 8, 4))
    o1, o4 = cases.make_oracle(case), cases.make_oracle(case)
    z = case.initial_state(seed=3)
    o1.set_state(*z)
    o4.set_state(*z)
    o1.step(O.default_options(ilu_blocks=1, **TIGHT))
    o4.step(O.default_options(ilu_blocks=4, **TIGHT))
    a, b = o1.get_state(), o4.get_state()
    assert np.linalg.norm(a["v"] - b["v"]) <= 1e-10 * np.linalg.norm(a["v"])
