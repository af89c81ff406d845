"""Immersed-boundary coupling.  The reference holds no IBM code (README.md:14 advertises it,
THEORY_GUIDE.md:130-132 is a TODO), so parity here is UNPINNED: the oracle's IBM section is this repo's own
definition of the method.  What these tests establish:
  * the definition has the properties the method needs (partition of unity, exact linear interpolation,
    interpolation and spreading adjoint, total force conserved) -- CPU, oracle only;
  * the product kernels (sorted markers, warp-per-marker gather, segment-reduced atomic scatter) and the step that
    uses them agree with the plain-loop definition to fp64 round-off -- through the C ABI; CPU tests run the
    host-logic test double, -m gpu tests the CUDA library."""
import numpy as np
import pytest

import fluca_b200 as fb
from tests import cases, parity


def _random_markers(case, n, seed, npts, span=(0.2, 0.8)):
    rng = np.random.default_rng(seed)
    lo, hi = np.array(case.lo, dtype=float), np.array(case.hi, dtype=float)
    X = lo[:, None] + (hi - lo)[:, None] * (span[0] + (span[1] - span[0]) * rng.random((case.dim, n)))
    # clusters of markers inside one cell exercise the segment reduction of the scatter
    X[:, n // 2 :] = X[:, : n - n // 2] + 1e-3 * rng.standard_normal((case.dim, n - n // 2))
    return dict(X=X, Ud=rng.standard_normal((case.dim, n)), dV=0.01 * (0.5 + rng.random(n)), npts=npts)


def _vol(case):
    h = [np.diff(x) for x in case.faces()]
    v = h[0][None, None, :] * h[1][None, :, None]
    return v * h[2][:, None, None] if case.dim == 3 else v


CASES = {
    "2d": lambda: cases.channel2d(n=(32, 16)),
    "3d": lambda: cases.channel3d(n=(16, 12, 12)),
    "3d_periodic_z": lambda: cases.channel3d(n=(16, 12, 12), periodic_z=True),
}


@pytest.mark.parametrize("npts", [3, 4])
@pytest.mark.parametrize("name", ["2d", "3d"])
def test_definition_properties(name, npts):
    case = CASES[name]()
    orc = cases.make_oracle(case)
    mk = _random_markers(case, 60, 1, npts)
    orc.set_markers(mk["X"], mk["Ud"], mk["dV"], npts)
    cell = orc.cell_shape
    assert np.abs(orc.ibm_interpolate(np.ones((case.dim,) + cell)) - 1.0).max() < 1e-13  # partition of unity
    g = cases._mesh(case.centres(), case.dim)
    lin = sum((d + 1.5) * g[d] for d in range(case.dim)) + 0 * g[0]
    exact = sum((d + 1.5) * mk["X"][d] for d in range(case.dim))
    assert np.abs(orc.ibm_interpolate(np.stack([np.broadcast_to(lin, cell)] * case.dim))[0] - exact).max() < 1e-12  # first moment
    rng = np.random.default_rng(2)
    v, F = rng.standard_normal((case.dim,) + cell), rng.standard_normal((case.dim, 60))
    Um, f = orc.ibm_interpolate(v), orc.ibm_spread(F)
    assert (Um * F * mk["dV"]).sum() == pytest.approx((f * v * _vol(case)).sum(), rel=1e-12)  # adjoint pair
    assert np.allclose((f * _vol(case)).reshape(case.dim, -1).sum(1), (F * mk["dV"]).sum(1), rtol=1e-12)  # total force


def _transfer_parity(lib, name, npts, stretch=0.0):
    case = CASES[name]()
    case.stretch = stretch
    orc = cases.make_oracle(case)
    span = (0.02, 0.98) if name == "3d_periodic_z" else (0.2, 0.8)
    mk = _random_markers(case, 200, 3, npts, span)
    if name == "3d_periodic_z":  # keep x, y inside; z may wrap
        lo, hi = np.array(case.lo), np.array(case.hi)
        for d in (0, 1):
            mk["X"][d] = np.clip(mk["X"][d], lo[d] + 0.2 * (hi[d] - lo[d]), hi[d] - 0.2 * (hi[d] - lo[d]))
        mk["X"][2] += 3.0 * (hi[2] - lo[2])  # several periods away: positions are folded back
    orc.set_markers(mk["X"], mk["Ud"], mk["dV"], npts)
    ns = parity.make_ns(case, lib, "fractional")
    fb.NSB200SetMarkers(ns, mk["X"], mk["Ud"], mk["dV"], npts)
    s = fb.NSB200GetSolver(ns)
    rng = np.random.default_rng(5)
    v, F = rng.standard_normal((case.dim,) + orc.cell_shape), rng.standard_normal((case.dim, 200))
    assert parity.rel(s.ibm_interpolate(v), orc.ibm_interpolate(v)) < 1e-13
    assert parity.rel(s.ibm_spread(F), orc.ibm_spread(F)) < 1e-13
    fb.NSDestroy(ns)


def _step_cases():
    c2 = cases.channel2d(n=(48, 24))  # [-2,6] x [-2,2], h = 1/6
    c3 = cases.channel3d(n=(24, 16, 16))  # [-2,4] x [-2,2]^2, h = 1/4
    return {
        "cylinder2d": (c2, cases.cylinder_markers((0.0, 0.0), 1.0, 64, 1.0 / 6.0)),
        "cylinder2d_roma_moving": (c2, cases.cylinder_markers((0.3, 0.1), 1.0, 40, 1.0 / 6.0, Ud=(0.2, -0.1), npts=3)),
        "sphere3d": (c3, cases.sphere_markers((0.0, 0.0, 0.0), 1.5, 150, 0.25)),
        "sphere3d_multidirect": (c3, dict(cases.sphere_markers((0.1, 0.0, -0.1), 1.5, 150, 0.25), iterations=3)),
    }


def _step_parity(lib, name, mode):
    case, mk = _step_cases()[name]
    out = parity.compare_steps(case, lib, mode=mode, nsteps=2, seed=41, tol=1e-10, markers=mk)
    assert all(o["Um"] <= 1e-10 for o in out)
    return out


# ------------------------------------------------------------------ CPU: host logic through the test double
@pytest.fixture(scope="module")
def emu():
    return parity.hostemu_library()


@pytest.mark.parametrize("npts", [3, 4])
@pytest.mark.parametrize("name", list(CASES))
def test_hostlogic_transfer_matches_definition(emu, name, npts):
    _transfer_parity(emu, name, npts, stretch=0.4 if name == "3d" else 0.0)


@pytest.mark.parametrize("mode", ["fractional", "coupled"])
@pytest.mark.parametrize("name", ["cylinder2d", "cylinder2d_roma_moving", "sphere3d", "sphere3d_multidirect"])
def test_hostlogic_step_with_markers_matches_definition(emu, name, mode):
    _step_parity(emu, name, mode)


def test_forcing_drives_marker_velocity_to_target(emu):
    """physical sanity of the coupling: the interpolated velocity at the markers approaches the prescribed one"""
    case, mk = _step_cases()["cylinder2d"]
    ns = parity.make_ns(case, emu, "fractional")
    v, U, p = case.initial_state()
    v[0] += 1.0
    U[0] += 1.0
    parity.set_initial(ns, (v, U, p))
    fb.NSB200SetMarkers(ns, mk["X"], mk["Ud"], mk["dV"], 4, iterations=4)
    s = fb.NSB200GetSolver(ns)
    slip = []
    for _ in range(4):
        fb.NSStep(ns)
        slip.append(float(np.sqrt((s.ibm_interpolate(s.get_state()["v"]) ** 2).sum(0)).mean()))
    F, _ = fb.NSB200GetMarkerForces(ns)
    assert slip[-1] < 0.25 and slip[-1] < slip[0], slip  # free stream is 1
    assert F[0].sum() < 0.0  # the body decelerates the fluid: drag on the body is positive
    fb.NSDestroy(ns)


# ------------------------------------------------------------------ GPU: the CUDA kernels
@pytest.fixture(scope="module")
def cuda():
    L = fb._lib.load()
    assert L.fluca_b200_is_host_emulation() == 0
    return L


@pytest.mark.gpu
@pytest.mark.parametrize("npts", [3, 4])
@pytest.mark.parametrize("name", list(CASES))
def test_gpu_transfer_matches_definition(cuda, name, npts):
    _transfer_parity(cuda, name, npts, stretch=0.4 if name == "3d" else 0.0)


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["fractional", "coupled"])
@pytest.mark.parametrize("name", ["cylinder2d", "cylinder2d_roma_moving", "sphere3d", "sphere3d_multidirect"])
def test_gpu_step_with_markers_matches_definition(cuda, name, mode):
    _step_parity(cuda, name, mode)


@pytest.mark.gpu
def test_gpu_many_markers_per_cell_conserve_force(cuda):
    """BASELINE config 4 density (~30 markers per surface cell): the segment-reduced scatter conserves the total force
    and matches the plain-loop definition"""
    case = cases.channel3d(n=(48, 32, 32))
    mk = cases.sphere_markers((0.0, 0.0, 0.0), 1.0, 20000, 0.125)
    orc = cases.make_oracle(case)
    orc.set_markers(mk["X"], mk["Ud"], mk["dV"], 4)
    ns = parity.make_ns(case, cuda, "fractional")
    fb.NSB200SetMarkers(ns, mk["X"], mk["Ud"], mk["dV"], 4)
    s = fb.NSB200GetSolver(ns)
    rng = np.random.default_rng(9)
    F = rng.standard_normal((3, 20000))
    f = s.ibm_spread(F)
    assert parity.rel(f, orc.ibm_spread(F)) < 1e-12
    assert np.allclose((f * _vol(case)).reshape(3, -1).sum(1), (F * mk["dV"]).sum(1), rtol=1e-11)
    v = rng.standard_normal((3,) + orc.cell_shape)
    assert parity.rel(s.ibm_interpolate(v), orc.ibm_interpolate(v)) < 1e-13
    fb.NSDestroy(ns)


def test_marker_count_may_change_between_calls(emu):
    """set_markers again with fewer / more markers (a body that is re-meshed): the arrays are re-laid out and results still
    match the definition"""
    case = CASES["3d"]()
    orc = cases.make_oracle(case)
    ns = parity.make_ns(case, emu, "fractional")
    s = fb.NSB200GetSolver(ns)
    rng = np.random.default_rng(11)
    v = rng.standard_normal((3,) + orc.cell_shape)
    for n in (120, 40, 200):
        mk = _random_markers(case, n, n, 4)
        orc.set_markers(mk["X"], mk["Ud"], mk["dV"], 4)
        fb.NSB200SetMarkers(ns, mk["X"], mk["Ud"], mk["dV"], 4)
        assert parity.rel(s.ibm_interpolate(v), orc.ibm_interpolate(v)) < 1e-13
        F = rng.standard_normal((3, n))
        assert parity.rel(s.ibm_spread(F), orc.ibm_spread(F)) < 1e-13
    fb.NSB200SetMarkers(ns, np.zeros((3, 0)), np.zeros((3, 0)), np.zeros(0), 4)  # n = 0 removes the markers
    fb.NSStep(ns)
    fb.NSDestroy(ns)
