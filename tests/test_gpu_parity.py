"""GPU parity tests proper: the CUDA product library (through the C ABI and the reference-style NS API)
against the CPU oracle on the same seeded inputs.  fp64 tolerance: 1e-10 relative L2 on velocity and
face velocity, 1e-9 on pressure, after K steps with both sides at tight solver tolerances
(BASELINE.json north_star: "within 1e-10 relative L2 (fp64) after a fixed number of steps")."""
import numpy as np
import pytest

import fluca_b200 as fb
from oracle import oracle as O
from tests import cases, parity

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib():
    L = fb._lib.load()  # the CUDA library; raises if it is missing (no fallback)
    assert L.fluca_b200_is_host_emulation() == 0
    return L


def _st(c):
    c.stretch = 0.6
    return c


CASES = [
    ("cavity2d_32", lambda: cases.cavity2d(n=32), None, 3),
    ("cavity2d_nonuniform", lambda: _st(cases.cavity2d(n=24)), 3, 2),
    ("cavity3d_sym_reference_geometry", lambda: cases.cavity3d(n=(16, 16, 8)), 5, 2),
    ("cavity3d_full", lambda: cases.cavity3d_full(n=(16, 16, 16)), 9, 2),
    ("tgv_dirichlet", lambda: cases.tgv(n=16, dt=0.05), None, 2),
    ("tgv_periodic", lambda: cases.tgv(n=16, periodic=True, dt=0.05), None, 2),
    ("channel2d_outlet_timedep", lambda: cases.channel2d(n=(32, 16), pout=0.3, time_dependent=True), 7, 2),
    ("channel3d_outlet", lambda: cases.channel3d(n=(16, 12, 12), pout=0.2, dt=0.05), 11, 2),
    ("channel3d_periodic_z", lambda: cases.channel3d(n=(16, 12, 12), periodic_z=True, dt=0.05), 13, 2),
    ("ragged_sizes", lambda: cases.cavity3d_full(n=(13, 9, 7)), 17, 2),
    # large enough in x / y for the TMA-staged tile kernels (>= 32 x 8 cells per plane); ragged extents put the
    # shifted last tile, the in-tile one-sided wall rows and the direct-load z-wall planes all on the path
    ("tma_cavity_ragged_nonuniform", lambda: _st(cases.cavity3d_full(n=(37, 13, 9))), 19, 2),
    ("tma_cavity_sym", lambda: cases.cavity3d(n=(48, 24, 10)), 23, 1),
    ("tma_channel_outlet", lambda: cases.channel3d(n=(40, 16, 10), pout=0.2, dt=0.05), 29, 1),
    ("tma_channel_periodic_z", lambda: cases.channel3d(n=(33, 9, 8), periodic_z=True, dt=0.05), 31, 1),
    # periodic x on the tile path (BASELINE config 5's boundary set): wrap columns of the first / last tile of a row; one tile
    # per row (both wraps in one CTA), two tiles, a shifted last tile (odd nx), and two tiled multigrid levels
    ("direct_channel5_periodic_xz", lambda: cases.channel_bench_case((16, 12, 12), periodic_z=True), 33, 2),
    ("tma_channel5_one_tile_per_row", lambda: cases.channel_bench_case((32, 16, 8), periodic_z=True), 35, 1),
    ("tma_channel5_two_tiles_sym_z", lambda: cases.channel_bench_case((64, 16, 8), periodic_z=False), 37, 1),
    ("tma_channel5_shifted_last_tile", lambda: cases.channel_bench_case((50, 9, 6), periodic_z=True), 39, 1),
    ("tma_channel5_two_mg_levels", lambda: cases.channel_bench_case((64, 32, 8), periodic_z=True), 41, 1),
]


@pytest.mark.parametrize("mode", ["fractional", "coupled"])
@pytest.mark.parametrize("name,mk,seed,nsteps", CASES, ids=[c[0] for c in CASES])
def test_step_matches_oracle(lib, name, mk, seed, nsteps, mode):
    out = parity.compare_steps(mk(), lib, mode=mode, nsteps=nsteps, seed=seed, tol=1e-10)
    if mode == "coupled":
        for o in out:  # KSP residual histories track the oracle's (exact inner solves)
            assert abs(o["outer"][0] - o["outer"][1]) <= 1
            n = min(len(o["hist_gpu"]), len(o["hist_orc"]), 6)
            for a, b in zip(o["hist_gpu"][:n], o["hist_orc"][:n]):
                assert a == pytest.approx(b, rel=1e-6, abs=1e-12 * o["hist_orc"][0])


@pytest.mark.parametrize("n", [(16, 12, 12), (70, 19, 6)], ids=["direct", "tma"])
def test_operator_level_parity(lib, n):
    case = cases.channel3d(n=n, pout=0.2, dt=0.05)
    orc = cases.make_oracle(case)
    state = case.initial_state(seed=21)
    orc.set_state(*state)
    ns = parity.make_ns(case, lib, "coupled", **parity.TIGHT)
    parity.set_initial(ns, state)
    rhs = orc.prepare_step()
    rm, ri, rc = ns.ops["formfunction"](ns)
    ov, oU, _ = orc.split(rhs)
    assert parity.rel(rm, ov) < 1e-13 and parity.relU(ri, oU) < 1e-13 and np.abs(rc).max() == 0.0
    s = fb.NSB200GetSolver(ns)
    rng = np.random.default_rng(0)
    x = rng.standard_normal(orc.nsol)
    xv, xU, xp = orc.split(x)
    A, S = orc.matrix("A"), orc.matrix("S")
    assert parity.rel(s.apply_momentum(xv).ravel(), A @ xv.ravel()) < 1e-13
    assert parity.rel(s.apply_schur(xp).ravel(), S @ xp.ravel()) < 1e-12
    G, negT, negR, D = (orc.matrix(k) for k in ("G", "negT", "negR", "D"))
    xUc = np.concatenate([u.ravel() for u in xU])
    gv, gU, gp = s.apply_coupled(xv, xU, xp)
    assert parity.rel(gv.ravel(), A @ xv.ravel() + G @ xp.ravel()) < 1e-13
    assert parity.rel(np.concatenate([u.ravel() for u in gU]), negT @ xv.ravel() + xUc + negR @ xp.ravel()) < 1e-13
    assert parity.rel(gp.ravel(), D @ xUc) < 1e-13
    xo, _ = orc.abf_apply(x, O.default_options(**parity.ORC_TIGHT))
    av, aU, ap, st = s.apply_abf(xv, xU, xp)
    o_v, o_U, o_p = orc.split(xo)
    assert parity.rel(av, o_v) < 1e-10 and parity.relU(aU, o_U) < 1e-10 and parity.rel(ap, o_p) < 1e-9


def test_reductions_are_bitwise_reproducible(lib):
    case = cases.cavity3d_full(n=(24, 20, 12))
    res = []
    for _ in range(2):
        ns = parity.make_ns(case, lib, "coupled")
        parity.set_initial(ns, case.initial_state(seed=4))
        for _ in range(2):
            fb.NSStep(ns)
        res.append(fb.NSB200GetSolver(ns).get_state()["v"].copy())
        fb.NSDestroy(ns)
    assert np.array_equal(res[0], res[1])


def test_full_size_properties_256cube_fractional_step(lib):
    """BASELINE config 3 at full size (256^3): size-independent properties instead of the oracle --
    discrete continuity D U = 0 after the projection, zero-mean pressure increment, bounded solve counts."""
    n = 256
    case = cases.cavity3d_full(n=(n, n, n), Re=400.0)
    ns = parity.make_ns(case, lib, "fractional", ns_abf_momentum_ksp_rtol=1e-10, ns_abf_schur_ksp_rtol=1e-10)
    for _ in range(2):
        fb.NSStep(ns)
    st = fb.NSB200GetStats(ns)
    assert st.mom_its <= 12 and st.schur_its <= 14  # mesh-independent multigrid
    s = fb.NSB200GetSolver(ns).get_state()
    h = 1.0 / n
    U = s["U"]
    div = (U[0][:, :, 1:] - U[0][:, :, :-1] + U[1][:, 1:, :] - U[1][:, :-1, :] + U[2][1:, :, :] - U[2][:-1, :, :]) / h
    scale = np.abs(U[0]).max() / h
    assert np.abs(div).max() <= 1e-7 * scale
    # walls: face-normal velocity equals the boundary data exactly
    assert np.abs(U[0][:, :, 0]).max() == 0.0 and np.abs(U[1][:, 0, :]).max() == 0.0 and np.abs(U[1][:, -1, :]).max() == 0.0
    assert np.isfinite(s["p"]).all() and abs(s["p"].mean()) < 1e-8 * max(np.abs(s["p"]).max(), 1e-30)
    fb.NSDestroy(ns)


@pytest.mark.parametrize("mk", [lambda: cases.cavity3d_full(n=(64, 32, 16)), lambda: _st(cases.channel3d(n=(70, 19, 12), pout=0.2, dt=0.05)), lambda: cases.channel3d(n=(40, 16, 8), periodic_z=True, dt=0.05)], ids=["uniform", "stretched_outlet", "periodic_z"])
def test_fused_presmoothing_equals_two_sweeps(lib, mk, monkeypatch):
    """the one-pass kernel for the first two Jacobi sweeps of a V-cycle (mg.cu MGFirstTwoTile) reproduces the two separate
    sweeps: same V-cycle output to round-off, on uniform (constant-row path) and stretched / outlet / periodic levels"""
    case = mk()
    ns = parity.make_ns(case, lib, "fractional")
    s = fb.NSB200GetSolver(ns)
    r = np.random.default_rng(3).standard_normal(s.cell_shape)
    monkeypatch.setenv("FLUCA_B200_NO_MG_FUSION", "1")
    z0 = s.apply_vcycle(r)
    monkeypatch.delenv("FLUCA_B200_NO_MG_FUSION")
    z1 = s.apply_vcycle(r)
    assert parity.rel(z1, z0) < 1e-12
    fb.NSDestroy(ns)


def test_config1_cavity2d_128_matches_oracle(lib):
    """BASELINE config 1 at its full size (2-D lid-driven cavity Re=100, 128x128, dt=0.5h): three coupled steps at tight
    tolerances against the oracle"""
    out = parity.compare_steps(cases.cavity2d(n=128), lib, mode="coupled", nsteps=3, tol=1e-10)
    assert all(o["outer"][0] == o["outer"][1] for o in out)


def test_config2_cylinder2d_full_size_properties(lib):
    """BASELINE config 2 at its full size (2-D cylinder by IBM, 1024x512 on [-8,24]x[-8,8], 1024 markers, inflow / outlet /
    symmetry): size-independent properties -- discrete continuity after the projection, the forcing reduces the slip at the
    markers, the body decelerates the fluid"""
    n = (1024, 512)
    c = cases.channel2d(n=n, Re=100.0, dt=0.5 / 32.0)
    c.lo, c.hi = (-8.0, -8.0), (24.0, 8.0)
    c.bcs[0]["velocity"] = cases._const(1.0, 0.0)
    c.bcs[1]["pressure"] = cases._constp(0.0)
    ns = parity.make_ns(c, lib, "fractional", ns_abf_momentum_ksp_rtol=1e-10, ns_abf_schur_ksp_rtol=1e-10)
    cell, face = c.shapes()
    v = np.zeros((2,) + cell)
    v[0] = 1.0
    U = [np.zeros(s) for s in face]
    U[0][...] = 1.0
    parity.set_initial(ns, (v, U, np.zeros(cell)))
    mk = cases.cylinder_markers((0.0, 0.0), 1.0, 1024, 1.0 / 32.0)
    fb.NSB200SetMarkers(ns, mk["X"], mk["Ud"], mk["dV"], 4, iterations=3)
    s = fb.NSB200GetSolver(ns)
    slips = []
    for _ in range(3):
        fb.NSStep(ns)
        slips.append(float(np.sqrt((s.ibm_interpolate(s.get_state()["v"]) ** 2).sum(0)).mean()))
    st = s.get_state()
    Ux, Uy = st["U"][0][0], st["U"][1][0]
    h = 1.0 / 32.0
    div = (Ux[:, 1:] - Ux[:, :-1] + Uy[1:, :] - Uy[:-1, :]) / h
    assert np.abs(div).max() <= 1e-6 * np.abs(Ux).max() / h
    F, _ = fb.NSB200GetMarkerForces(ns)
    assert slips[2] < slips[1] < slips[0] < 0.8 and slips[2] < 0.55, slips  # free stream is 1
    assert F[0].sum() < 0.0 and np.isfinite(st["p"]).all()
    fb.NSDestroy(ns)


def test_config4_sphere512_full_size_properties(lib):
    """BASELINE config 4 at its full size (512^3, 100k markers) in fractional mode: one step; discrete continuity, finite
    fields, forcing acts against the free stream"""
    n = 512
    c = cases.sphere_bench_case(n, n)
    ns = parity.make_ns(c, lib, "fractional", ns_abf_momentum_ksp_rtol=1e-8, ns_abf_schur_ksp_rtol=1e-8)
    s = fb.NSB200GetSolver(ns)
    v, U, p = cases.uniform_inflow_state(c)
    s.set_state(v=v, U=U, p=p, phalf=p)
    del v, U, p
    mk = cases.sphere_markers((0.0, 0.0, 0.0), 1.0, 100000, 16.0 / n)
    fb.NSB200SetMarkers(ns, mk["X"], mk["Ud"], mk["dV"], 4)
    fb.NSStep(ns)
    stt = fb.NSB200GetStats(ns)
    assert stt.mom_its <= 16 and stt.schur_its <= 16
    st = s.get_state()
    U = st["U"]
    h = 16.0 / n
    k = slice(200, 312)  # a slab through the sphere is enough for the host-side check
    div = (U[0][k, :, 1:] - U[0][k, :, :-1] + U[1][k, 1:, :] - U[1][k, :-1, :] + U[2][201:313, :, :] - U[2][200:312, :, :]) / h
    assert np.abs(div).max() <= 1e-5 / h
    F, Um = fb.NSB200GetMarkerForces(ns)
    assert F[0].sum() < 0.0 and np.isfinite(st["p"]).all() and np.abs(Um[0]).max() <= 1.5
    fb.NSDestroy(ns)


@pytest.mark.parametrize("restart", [1, 3])
def test_restarted_outer_gmres_matches_oracle(lib, restart):
    """short restarts (BASELINE config 5 runs GMRES(1) for memory): the restart residual is formed in the last basis vector"""
    case = cases.cavity3d(n=(48, 24, 10))
    state = case.initial_state(seed=23)
    orc = cases.make_oracle(case)
    orc.set_state(*state)
    ns = parity.make_ns(case, lib, "coupled", **dict(parity.TIGHT, ns_ksp_gmres_restart=restart, ns_ksp_max_it=400))
    parity.set_initial(ns, state)
    orc.step(O.default_options(mode=0, **parity.ORC_TIGHT))
    fb.NSStep(ns)
    st = fb.NSB200GetStats(ns)
    assert st.converged and st.outer_its > restart
    a, b = orc.get_state(), fb.NSB200GetSolver(ns).get_state()
    assert parity.rel(b["v"], a["v"]) < 1e-10 and parity.relU(b["U"], a["U"]) < 1e-10 and parity.rel(b["p"], a["p"]) < 1e-9
    fb.NSDestroy(ns)
