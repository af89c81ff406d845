"""PCABF variants: the DIAG and ROWSUM approximations of A^-1 in the Schur complement and in the upper
triangular factor (PCABFAinvType, flucans.h:99-107; abfpc.c:80-99, 151-170).

Three layers:
  * the oracle's restatement against an independent scipy composition of the reference's recipe from the oracle's own
    blocks, and the properties the variants must have (CPU);
  * the solver library's host logic through the host-emulation test double against the oracle (CPU);
  * the CUDA product library against the oracle (-m gpu), including a mesh that puts the TMA-staged momentum operator
    under the variant ABF factors."""
import numpy as np
import pytest
import scipy.sparse as sp

import fluca_b200 as fb
from oracle import oracle as O
from tests import cases, parity

ID, DIAG, ROWSUM = 0, 1, 2


# ------------------------------------------------------------------ oracle (CPU)
@pytest.mark.parametrize("mk,seed", [(lambda: cases.cavity2d(n=12), 3), (lambda: cases.channel3d(n=(8, 6, 6), pout=0.2, dt=0.05), 11)], ids=["cavity2d", "channel3d_outlet"])
@pytest.mark.parametrize("kind", [DIAG, ROWSUM])
def test_oracle_schur_is_the_reference_recipe(mk, seed, kind):
    """S = D ((-T) diag(a)^-1 G~ - (-R)), a = MatGetDiagonal(A) or MatGetRowSum(A) (abfpc.c:155-170), composed here with scipy."""
    case = mk()
    orc = cases.make_oracle(case)
    orc.set_state(*case.initial_state(seed=seed))
    orc.prepare_step(O.default_options(mode=1, schur_ainv=kind, upper_ainv=kind))
    A, G, negT, negR, D, S = (orc.matrix(k) for k in ("A", "G", "negT", "negR", "D", "S"))
    a = A.diagonal() if kind == DIAG else np.asarray(A.sum(axis=1)).ravel()
    ref = D @ (negT @ (sp.diags(1.0 / a) @ G) - negR)
    assert abs(ref - S).max() <= 1e-12 * abs(S).max()
    # the ID variant cancels the wide stencil; the variants keep a genuine 5-wide star
    orc.prepare_step(O.default_options(mode=1))
    S0 = orc.matrix("S")
    assert abs(S - S0).max() > 1e-6 * abs(S0).max()


@pytest.mark.parametrize("ainv", [(DIAG, DIAG), (ROWSUM, ROWSUM), (DIAG, ID), (ID, ROWSUM)])
def test_oracle_coupled_solution_does_not_depend_on_the_preconditioner_variant(ainv):
    case = cases.cavity2d(n=16, dt=0.02)

    def run(sa, ua):
        orc = cases.make_oracle(case)
        orc.set_state(*case.initial_state())
        opt = O.default_options(mode=0, schur_ainv=sa, upper_ainv=ua, **parity.ORC_TIGHT)
        its = [orc.step(opt).outer_its for _ in range(2)]
        return orc.get_state(), its

    base, _ = run(ID, ID)
    st, its = run(*ainv)
    assert parity.rel(st["v"], base["v"]) < 1e-10 and parity.rel(st["p"], base["p"]) < 1e-9
    assert all(0 < i < 40 for i in its)


def test_oracle_diag_fractional_step_is_closer_to_the_coupled_solution():
    """What the variants are for (THEORY_GUIDE.md:299-308): on the 32^2 cavity at dt = 0.02 one DIAG application
    leaves less splitting error than one ID application."""
    case = cases.cavity2d(n=32, Re=100.0, dt=0.02)

    def run(mode, kind):
        orc = cases.make_oracle(case)
        orc.set_state(*case.initial_state())
        opt = O.default_options(mode=mode, schur_ainv=kind, upper_ainv=kind, **parity.ORC_TIGHT)
        for _ in range(2):
            orc.step(opt)
        return orc.get_state()["v"]

    exact = run(0, ID)
    e_id, e_diag = parity.rel(run(1, ID), exact), parity.rel(run(1, DIAG), exact)
    assert e_diag < 0.6 * e_id


# ------------------------------------------------------------------ shared comparisons (host emulation on CPU, CUDA on GPU)
def _operator_parity(lib, case, seed, kind):
    orc = cases.make_oracle(case)
    state = case.initial_state(seed=seed)
    orc.set_state(*state)
    name = parity.AINV_OPTION[kind]
    ns = parity.make_ns(case, lib, "fractional", ns_pc_abf_schur_ainv_type=name, ns_pc_abf_upper_ainv_type=name, **parity.TIGHT)
    parity.set_initial(ns, state)
    oopt = O.default_options(mode=1, schur_ainv=kind, upper_ainv=kind, **parity.ORC_TIGHT)
    rhs = orc.prepare_step(oopt)
    ns.ops["formfunction"](ns)
    s = fb.NSB200GetSolver(ns)
    rng = np.random.default_rng(0)
    p = rng.standard_normal(s.cell_shape)
    S = orc.matrix("S")
    assert parity.rel(s.apply_schur(p).ravel(), S @ p.ravel()) < 1e-11
    # one PCApply_ABF on the step's own right-hand side
    ov, oU, op = orc.split(rhs)
    xo, _ = orc.abf_apply(rhs, oopt)
    xv, xU, xp = orc.split(xo)
    gv, gU, gp, st = s.apply_abf(ov, oU, op)  # r_con = 0: compatible with the constant null space as it stands
    assert parity.rel(gv, xv) < 1e-10 and parity.relU(gU, xU) < 1e-10 and parity.rel(gp, xp) < 1e-9
    assert st.abf_applies == 1 and st.mom_its > 0 and st.schur_its > 0
    fb.NSDestroy(ns)


def _st(c, s=0.3):
    c.stretch = s
    return c


EMU_CASES = [
    ("cavity2d", lambda: cases.cavity2d(n=16), None, (DIAG, DIAG)),
    ("cavity2d_rowsum", lambda: cases.cavity2d(n=16), None, (ROWSUM, ROWSUM)),
    ("cavity2d_nonuniform", lambda: _st(cases.cavity2d(n=12)), None, (ROWSUM, ROWSUM)),
    ("cavity2d_nonuniform_mixed", lambda: _st(cases.cavity2d(n=12)), None, (DIAG, ROWSUM)),
    ("cavity3d_sym", lambda: cases.cavity3d(n=(8, 8, 4)), None, (ROWSUM, DIAG)),
    ("tgv_periodic", lambda: cases.tgv(n=8, periodic=True, dt=0.05), None, (DIAG, DIAG)),
    ("channel3d_outlet_seeded", lambda: cases.channel3d(n=(8, 6, 6), pout=0.2, dt=0.05), 5, (DIAG, DIAG)),
    ("channel3d_periodic_z_schur_only", lambda: cases.channel3d(n=(8, 6, 6), periodic_z=True), None, (ROWSUM, ID)),
    ("channel2d_upper_only", lambda: cases.channel2d(n=(16, 8), pout=0.3, time_dependent=True), None, (ID, DIAG)),
]


def _steps(lib, mk, seed, ainv, mode, nsteps=2):
    case = mk()
    out = parity.compare_steps(case, lib, mode=mode, nsteps=nsteps, seed=seed, tol=1e-10, ainv=ainv)
    # Different Schur / upper types on a NON-UNIFORM mesh: the continuity block of M ABF(v) is then (S_upper - S_schur) p,
    # whose volume-weighted sum vanishes but whose plain mean -- the one PETSc's null-space removal subtracts
    # (KSP_RemoveNullSpace with the constant vector of nsbasic.c:229-243) -- does not, so later Schur right-hand sides are
    # slightly inconsistent and each inner solver answers in its own way (GMRES + ILU(0) there, projected right-hand
    # side + BiCGStab here).  The converged step is unaffected (asserted above); only the outer history is not comparable.
    comparable = ainv[0] == ainv[1] or not case.stretch
    if mode == "coupled" and comparable:
        for o in out:  # same (M, M~, b) on both sides: the outer Krylov history tracks the oracle's
            assert abs(o["outer"][0] - o["outer"][1]) <= 1
            n = min(len(o["hist_gpu"]), len(o["hist_orc"]), 6)
            for a, b in zip(o["hist_gpu"][:n], o["hist_orc"][:n]):
                assert a == pytest.approx(b, rel=1e-6, abs=1e-12 * o["hist_orc"][0])
    return out


@pytest.fixture(scope="module")
def emu():
    return parity.hostemu_library()


@pytest.mark.parametrize("mode", ["fractional", "coupled"])
@pytest.mark.parametrize("name,mk,seed,ainv", EMU_CASES, ids=[c[0] for c in EMU_CASES])
def test_hostlogic_step_matches_oracle(emu, name, mk, seed, ainv, mode):
    _steps(emu, mk, seed, ainv, mode)


@pytest.mark.parametrize("kind", [DIAG, ROWSUM])
def test_hostlogic_operator_level(emu, kind):
    _operator_parity(emu, cases.channel3d(n=(8, 6, 6), pout=0.2, dt=0.05), 21, kind)
    _operator_parity(emu, _st(cases.cavity3d(n=(8, 6, 4))), 4, kind)


def test_hostlogic_types_can_change_between_steps_and_bad_types_are_rejected(emu):
    """PCABFSet*AinvType may be called at any time (abfpc.c:300-318); PCSetUp_ABF of the next step picks it up."""
    case = cases.cavity2d(n=12)
    orc = cases.make_oracle(case)
    orc.set_state(*case.initial_state())
    ns = parity.make_ns(case, emu, "fractional", **parity.TIGHT)
    parity.set_initial(ns, case.initial_state())
    for kind in (ID, DIAG, ROWSUM, ID):
        fb.PCABFSetSchurComplementAinvType(ns, kind)
        fb.PCABFSetUpperTriangularAinvType(ns, kind)
        orc.step(O.default_options(mode=1, schur_ainv=kind, upper_ainv=kind, **parity.ORC_TIGHT))
        fb.NSStep(ns)
        a, b = orc.get_state(), fb.NSB200GetSolver(ns).get_state()
        assert parity.rel(b["v"], a["v"]) < 1e-10 and parity.rel(b["p"], a["p"]) < 1e-9
    with pytest.raises(fb._lib.FlucaB200Error):
        fb.NSB200GetSolver(ns).set_abf_ainv_types(3, 0)
    fb.NSDestroy(ns)
    with pytest.raises(fb.FlucaError):
        parity.make_ns(case, emu, "fractional", ns_pc_abf_schur_ainv_type="LUMPED")


@pytest.mark.parametrize("mode", ["fractional", "coupled"])
def test_hostlogic_immersed_boundary_under_variant_factors(emu, mode):
    """The direct-forcing term is a right-hand-side term: it composes with any ABF variant."""
    case = cases.channel3d(n=(12, 8, 8), pout=0.1, dt=0.05)
    mk = cases.sphere_markers((0.1, 0.0, 0.05), 1.2, 60, 0.5)
    parity.compare_steps(case, emu, mode, nsteps=2, seed=31, markers=mk, ainv=(DIAG, DIAG))


# ------------------------------------------------------------------ CUDA product library
@pytest.fixture(scope="module")
def lib():
    L = fb._lib.load()  # the CUDA library; raises if it is missing (no fallback)
    assert L.fluca_b200_is_host_emulation() == 0
    return L


GPU_CASES = [
    ("cavity2d_32", lambda: cases.cavity2d(n=32), None, (DIAG, DIAG), 2),
    ("cavity2d_nonuniform_mixed", lambda: _st(cases.cavity2d(n=24)), None, (DIAG, ROWSUM), 2),
    ("cavity3d_sym", lambda: cases.cavity3d(n=(16, 16, 8)), None, (ROWSUM, ROWSUM), 2),
    ("tgv_periodic", lambda: cases.tgv(n=16, periodic=True, dt=0.05), None, (DIAG, DIAG), 2),
    ("channel3d_outlet_seeded", lambda: cases.channel3d(n=(16, 12, 12), pout=0.2, dt=0.02), 5, (DIAG, DIAG), 2),
    ("channel3d_periodic_z_schur_only", lambda: cases.channel3d(n=(16, 12, 12), periodic_z=True, dt=0.05), None, (ROWSUM, ID), 2),
    # >= 32 x 8 cells per plane: the momentum solves under the variant factors run from the TMA-staged tiles
    ("tma_cavity_ragged_nonuniform", lambda: _st(cases.cavity3d_full(n=(37, 13, 9))), None, (DIAG, DIAG), 1),
    ("tma_channel_outlet_upper_only", lambda: cases.channel3d(n=(40, 16, 10), pout=0.2, dt=0.05), None, (ID, ROWSUM), 1),
]


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["fractional", "coupled"])
@pytest.mark.parametrize("name,mk,seed,ainv,nsteps", GPU_CASES, ids=[c[0] for c in GPU_CASES])
def test_gpu_step_matches_oracle(lib, name, mk, seed, ainv, nsteps, mode):
    _steps(lib, mk, seed, ainv, mode, nsteps)


@pytest.mark.gpu
@pytest.mark.parametrize("kind", [DIAG, ROWSUM])
def test_gpu_operator_level(lib, kind):
    _operator_parity(lib, cases.channel3d(n=(16, 12, 12), pout=0.2, dt=0.05), 21, kind)
    _operator_parity(lib, _st(cases.cavity3d_full(n=(37, 13, 9))), 4, kind)
