"""FlucaFDApply on the device (csrc/fd.cu) against the point-by-point definition of fdapply.c:85-106 evaluated with the
golden-pinned stencil layer (tests/test_fd_stencils.py) at EVERY output point, boundary closures and corners included:
v1, the class-table kernel (derivative / sum / constant scale / composition on uniform coordinates), and v2, the assembled
(ELL) apply that serves everything else -- vector scale, second-order TVD with its limiters (the operator of
tutorials/fd/ex1.c), non-uniform coordinates, wide periodic stencils -- and FlucaFDGetOperator (fdapply.c:123-180).

CPU: the host-emulation build of the same sources; -m gpu: the CUDA library.  (The file sorts last on purpose.)"""
import numpy as np
import pytest

import fluca_b200 as fb
from fluca_b200 import fd as FD
from tests import parity

E, L, D, B = FD.DMSTAG_ELEMENT, FD.DMSTAG_LEFT, FD.DMSTAG_DOWN, FD.DMSTAG_BACK


def reference_apply(op, grid, fields, out_loc, bc_values):
    """fdapply.c:85-106 in plain loops over FlucaFDGetStencil"""
    shape = grid.field_shape(out_loc)
    full = (1,) * (3 - len(shape)) + shape
    out = np.zeros(full)
    per = list(grid.periodic) + [False] * (3 - grid.dim)
    for k in range(full[0]):
        for j in range(full[1]):
            for i in range(full[2]):
                r = 0.0
                for (ci, cj, ck, loc, c), v in op.GetStencil(i, j, k):
                    if c >= 0:
                        f = fields[(loc, c)]
                        f3 = f.reshape((1,) * (3 - f.ndim) + f.shape)
                        idx = [ck, cj, ci]
                        for a, d in ((2, 0), (1, 1), (0, 2)):
                            if per[d]:
                                idx[a] %= f3.shape[a]
                        r += v * f3[idx[0], idx[1], idx[2]]
                    elif c == FD.FLUCAFD_CONSTANT:
                        r += v
                    else:
                        r += v * bc_values[-c - 1]
                out[k, j, i] = r
    return out.reshape(shape)


def field(grid, loc, seed):
    return np.random.default_rng(seed).standard_normal(grid.field_shape(loc))


def deriv(g, d, order, accu, il, ol):
    return FD.FlucaFDDerivativeCreate(g, d, order, accu, il, 0, ol, 0)


def case_second_derivative_1d(lib):
    g = FD.FDGrid.uniform([16], [0.0], [2.0], library=lib)
    op = deriv(g, 0, 2, 2, E, E)
    bc = [0.0] * 6
    op.SetBoundaryCondition(0, FD.FLUCAFD_BC_DIRICHLET, 0.7), op.SetBoundaryCondition(1, FD.FLUCAFD_BC_NEUMANN, -1.3)
    bc[0], bc[1] = 0.7, -1.3
    return g, op.SetUp(), E, bc


def case_laplacian_3d_mixed(lib):
    g = FD.FDGrid.uniform([10, 9, 8], [0.0, 0.0, 0.0], [1.0, 2.0, 1.5], periodic=[False, False, True], library=lib)
    ops = [deriv(g, d, 2, 2, E, E).SetUp() for d in range(3)]
    s = FD.FlucaFDSumCreate(ops)
    bc = [0.3, 0.0, -0.4, 1.1, 0.0, 0.0]
    s.SetBoundaryCondition(0, FD.FLUCAFD_BC_DIRICHLET, bc[0]), s.SetBoundaryCondition(2, FD.FLUCAFD_BC_NEUMANN, bc[2]), s.SetBoundaryCondition(3, FD.FLUCAFD_BC_DIRICHLET, bc[3])
    return g, s.SetUp(), E, bc


def case_divergence_of_face_fields(lib):
    g = FD.FDGrid.uniform([10, 9, 8], [0.0] * 3, [1.0] * 3, library=lib)
    ops = [deriv(g, d, 1, 2, loc, E).SetUp() for d, loc in enumerate((L, D, B))]
    return g, FD.FlucaFDSumCreate(ops).SetUp(), E, [0.0] * 6


def case_compact_laplacian_2d(lib):
    g = FD.FDGrid.uniform([12, 10], [0.0, 0.0], [1.0, 1.0], library=lib)
    inner = deriv(g, 0, 1, 2, E, L).SetUp()
    outer = deriv(g, 0, 1, 2, L, E).SetUp()
    comp = FD.FlucaFDCompositionCreate(inner, outer)
    comp.SetBoundaryCondition(0, FD.FLUCAFD_BC_DIRICHLET, 0.25), comp.SetBoundaryCondition(1, FD.FLUCAFD_BC_NEUMANN, 2.0)
    sc = FD.FlucaFDScaleCreateConstant(comp.SetUp(), -0.5)
    sc.SetBoundaryCondition(0, FD.FLUCAFD_BC_DIRICHLET, 0.25), sc.SetBoundaryCondition(1, FD.FLUCAFD_BC_NEUMANN, 2.0)
    return g, sc.SetUp(), E, [0.25, 2.0, 0, 0, 0, 0]


def case_gradient_to_faces_periodic(lib):
    g = FD.FDGrid.uniform([9, 8], [0.0, 0.0], [3.0, 1.0], periodic=[True, False], library=lib)
    op = deriv(g, 1, 1, 2, E, D)
    op.SetBoundaryCondition(2, FD.FLUCAFD_BC_DIRICHLET, 0.5)
    return g, op.SetUp(), D, [0, 0, 0.5, 0, 0, 0]


CASES = [case_second_derivative_1d, case_laplacian_3d_mixed, case_divergence_of_face_fields, case_compact_laplacian_2d, case_gradient_to_faces_periodic]


def _check(lib, make):
    g, op, out_loc, bc = make(lib)
    order = op.ApplyInputs()
    fields = {key: field(g, key[0], 10 + n) for n, key in enumerate(order)}
    got = op.Apply(fields, out_loc)
    want = reference_apply(op, g, fields, out_loc, bc)
    assert got.shape == want.shape and np.abs(got - want).max() <= 1e-11 * max(1.0, np.abs(want).max()), (make.__name__, np.abs(got - want).max())


@pytest.mark.parametrize("make", CASES, ids=[c.__name__[5:] for c in CASES])
def test_apply_host_emulation(make):
    _check(parity.hostemu_library(), make)


def test_device_resident_form_and_plan_rebuild_host_emulation():
    """apply_device takes device pointers (host pointers in the test double) and reuses the cached tables; changing a
    boundary value rebuilds them (the values are folded into the per-variant constants)."""
    lib = parity.hostemu_library()
    g, op, out_loc, bc = case_laplacian_3d_mixed(lib)
    order = op.ApplyInputs()
    fields = {key: field(g, key[0], 3) for key in order}
    out = np.zeros(g.field_shape(out_loc))
    for _ in range(2):
        op.ApplyDevice([fields[k].ctypes.data for k in order], out.ctypes.data)
        op.Sync()
        assert np.abs(out - reference_apply(op, g, fields, out_loc, bc)).max() < 1e-11
    bc2 = list(bc)
    bc2[0] = -2.5
    op.SetBoundaryCondition(0, FD.FLUCAFD_BC_DIRICHLET, bc2[0])
    want = reference_apply(op, g, fields, out_loc, bc2)
    op.ApplyDevice([fields[k].ctypes.data for k in order], out.ctypes.data)
    op.Sync()
    assert np.abs(out - want).max() < 1e-11 and np.abs(want - reference_apply(op, g, fields, out_loc, bc)).max() > 1e-3


# ---- v2: what the class-table kernel cannot serve goes through the assembled (ELL) apply: vector scale, second-order TVD with
# its limiters, non-uniform coordinates, periodic stencils wider than the ghost width -- the operators of tutorials/fd/ex1-4.c
def case_vector_scale_1d(lib):
    g = FD.FDGrid.uniform([12], [0.0], [1.0], library=lib)
    d = deriv(g, 0, 1, 2, E, E)
    d.SetBoundaryCondition(0, FD.FLUCAFD_BC_DIRICHLET, 0.4)
    sv = FD.FlucaFDScaleCreateVector(d.SetUp(), 1.0 + np.arange(12.0) ** 1.5, E)
    sv.SetBoundaryCondition(0, FD.FLUCAFD_BC_DIRICHLET, 0.4)
    return g, sv.SetUp(), E, [0.4, 0, 0, 0, 0, 0]


def case_nonuniform_laplacian_2d(lib):
    g = FD.FDGrid([10, 8], [np.linspace(0, 1, 11) ** 1.5, 2.0 * np.linspace(0, 1, 9) ** 0.8], library=lib)
    ops = [deriv(g, d, 2, 2, E, E).SetUp() for d in range(2)]
    s_ = FD.FlucaFDSumCreate(ops)
    s_.SetBoundaryCondition(0, FD.FLUCAFD_BC_DIRICHLET, 0.3), s_.SetBoundaryCondition(3, FD.FLUCAFD_BC_NEUMANN, -0.7)
    return g, s_.SetUp(), E, [0.3, 0, 0, -0.7, 0, 0]


def case_wide_periodic_third_derivative(lib):
    g = FD.FDGrid.uniform([12], [0.0], [1.0], periodic=[True], stencil_width=1, library=lib)
    return g, deriv(g, 0, 3, 2, E, E).SetUp(), E, [0.0] * 6


def tutorial_convection_diffusion(lib, limiter, n=24, rho=1.3, gamma=0.05):
    """the operator of tutorials/fd/ex1.c:46-110: d/dx(rho u phi) by second-order TVD composed with a face-to-element derivative,
    minus d/dx(gamma d/dx phi); Dirichlet phi(0) = 1, phi(1) = 0"""
    g = FD.FDGrid.uniform([n], [0.0], [1.0], library=lib)

    def bcs(op):
        op.SetBoundaryCondition(0, FD.FLUCAFD_BC_DIRICHLET, 1.0), op.SetBoundaryCondition(1, FD.FLUCAFD_BC_DIRICHLET, 0.0)
        return op

    tvd = bcs(FD.FlucaFDSecondOrderTVDCreate(g, FD.FLUCAFD_X, 0, 0))
    FD.FlucaFDSecondOrderTVDSetLimiter(tvd, limiter)
    tvd.SetUp()
    conv = bcs(FD.FlucaFDCompositionCreate(FD.FlucaFDScaleCreateConstant(tvd, rho).SetUp(), deriv(g, 0, 1, 2, L, E).SetUp())).SetUp()
    inner = deriv(g, 0, 1, 2, E, L).SetUp()
    diff = bcs(FD.FlucaFDCompositionCreate(FD.FlucaFDScaleCreateConstant(inner, gamma).SetUp(), deriv(g, 0, 1, 2, L, E).SetUp())).SetUp()
    op = bcs(FD.FlucaFDSumCreate([conv, FD.FlucaFDScaleCreateConstant(diff, -1.0).SetUp()])).SetUp()
    return g, op, tvd


def tutorial_ex3_rhs_2d(lib, limiter, n=(14, 12), mu=0.01):
    """the right-hand side operator of tutorials/fd/ex3.c:55-125: -(d/dx)(u phi) - (d/dy)(v phi) + (d/dx)(mu dphi/dx) + (d/dy)(mu dphi/dy) on
    a doubly periodic square, the two convective fluxes by second-order TVD"""
    g = FD.FDGrid.uniform(list(n), [0.0, 0.0], [1.0, 1.0], periodic=[True, True], library=lib)
    tvds, terms = [], []
    for d, face in ((0, L), (1, D)):
        t = FD.FlucaFDSecondOrderTVDCreate(g, d, 0, 0)
        FD.FlucaFDSecondOrderTVDSetLimiter(t, limiter)
        tvds.append(t.SetUp())
        conv = FD.FlucaFDCompositionCreate(t, deriv(g, d, 1, 2, face, E).SetUp()).SetUp()
        terms.append(FD.FlucaFDScaleCreateConstant(conv, -1.0).SetUp())
        inner = FD.FlucaFDScaleCreateConstant(deriv(g, d, 1, 2, E, face).SetUp(), mu).SetUp()
        terms.append(FD.FlucaFDCompositionCreate(inner, deriv(g, d, 1, 2, face, E).SetUp()).SetUp())
    return g, FD.FlucaFDSumCreate(terms).SetUp(), tvds


def _ex3_check(lib, limiter):
    g, op, tvds = tutorial_ex3_rhs_2d(lib, limiter)
    rng = np.random.default_rng(7)
    FD.FlucaFDSecondOrderTVDSetVelocity(tvds[0], 1.0 + 0.2 * rng.standard_normal(g.field_shape(L)))
    FD.FlucaFDSecondOrderTVDSetVelocity(tvds[1], -0.7 + 0.2 * rng.standard_normal(g.field_shape(D)))
    for seed in (0, 1):
        phi = np.random.default_rng(seed).uniform(0.0, 1.0, g.field_shape(E))
        for t in tvds:
            FD.FlucaFDSecondOrderTVDSetCurrentSolution(t, phi)
        got = op.Apply({(E, 0): phi}, E)
        want = reference_apply(op, g, {(E, 0): phi}, E, [0.0] * 6)
        assert np.abs(got - want).max() <= 1e-11 * max(1.0, np.abs(want).max()), (limiter, seed)


@pytest.mark.parametrize("limiter", ["minmod", "vanleer", "mc"])
def test_tutorial_ex3_2d_periodic_tvd_operator_host_emulation(limiter):
    _ex3_check(parity.hostemu_library(), limiter)


V2_CASES = [case_vector_scale_1d, case_nonuniform_laplacian_2d, case_wide_periodic_third_derivative]


@pytest.mark.parametrize("make", V2_CASES, ids=[c.__name__[5:] for c in V2_CASES])
def test_assembled_apply_host_emulation(make):
    _check(parity.hostemu_library(), make)


@pytest.mark.parametrize("make", CASES[:3], ids=[c.__name__[5:] for c in CASES[:3]])
def test_assembled_apply_equals_the_class_table_kernel(make, monkeypatch):
    """the operators v1 serves give the same result through the assembled path (FLUCA_B200_FD_ASSEMBLED forces it)"""
    lib = parity.hostemu_library()
    g, op, out_loc, bc = make(lib)
    order = op.ApplyInputs()
    fields = {key: field(g, key[0], 20 + n) for n, key in enumerate(order)}
    a = op.Apply(fields, out_loc)
    monkeypatch.setenv("FLUCA_B200_FD_ASSEMBLED", "1")
    g2, op2, _, _ = make(lib)
    assert op2.ApplyInputs() == order
    b = op2.Apply(fields, out_loc)
    assert np.abs(a - b).max() <= 1e-12 * max(1.0, np.abs(a).max())


def _tvd_check(lib, limiter):
    g, op, tvd = tutorial_convection_diffusion(lib, limiter)
    n = g.n[0]
    FD.FlucaFDSecondOrderTVDSetVelocity(tvd, 1.0 + 0.3 * np.cos(np.arange(n + 1.0)))
    bc = [1.0, 0.0, 0, 0, 0, 0]
    for seed in (0, 1):  # the stencil depends on the current solution: a second field re-assembles the table
        phi = np.random.default_rng(seed).uniform(0.0, 1.0, n)
        FD.FlucaFDSecondOrderTVDSetCurrentSolution(tvd, phi)
        assert op.ApplyInputs() == [(E, 0)]
        got = op.Apply({(E, 0): phi}, E)
        want = reference_apply(op, g, {(E, 0): phi}, E, bc)
        assert np.abs(got - want).max() <= 1e-11 * max(1.0, np.abs(want).max()), (limiter, seed)


@pytest.mark.parametrize("limiter", ["minmod", "vanleer", "superbee", "mc", "vanalbada"])
def test_tutorial_tvd_operator_host_emulation(limiter):
    _tvd_check(parity.hostemu_library(), limiter)


def test_get_operator_is_the_matrix_of_apply():
    """FlucaFDGetOperator: CSR of the interior stencil points; apply(x) = A x + apply(0) (the boundary / constant part)"""
    lib = parity.hostemu_library()
    g, op, out_loc, bc = case_laplacian_3d_mixed(lib)
    rowptr, cols, vals = op.GetOperator()
    shape = g.field_shape(E)
    nr = int(np.prod(shape))
    assert len(rowptr) == nr + 1 and rowptr[-1] == len(cols) == len(vals)
    x = field(g, E, 5)
    y = np.zeros(nr)
    n = list(g.n) + [1] * (3 - g.dim)
    for r in range(nr):
        for q in range(rowptr[r], rowptr[r + 1]):
            ci, cj, ck, loc, c = cols[q]
            assert loc == E and c == 0
            y[r] += vals[q] * x[ck % n[2], cj, ci]  # z is periodic in this case: ghost columns wrap
    zero = op.Apply({(E, 0): np.zeros(shape)}, E)
    full = op.Apply({(E, 0): x}, E)
    assert np.abs(full.ravel() - (y + zero.ravel())).max() <= 1e-10 * np.abs(full).max()
    # the stencil at one interior point is that row
    i, j, k = 4, 3, 2
    r = i + shape[2] * (j + shape[1] * k)
    st = [(cv, v) for cv, v in op.GetStencil(i, j, k) if cv[4] >= 0]
    assert sorted((cols[q], vals[q]) for q in range(rowptr[r], rowptr[r + 1])) == sorted(st)


def test_product_library_refuses_to_apply_without_a_device():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    lib = fb._lib.load()
    g = FD.FDGrid.uniform([8], [0.0], [1.0], library=lib)
    d = deriv(g, 0, 1, 2, E, E).SetUp()
    assert d.ApplyInputs() == [(E, 0)]  # planning is host work
    with pytest.raises(FD.FlucaFDError, match="no CPU fallback"):
        d.Apply([np.zeros(8)], E)


@pytest.mark.gpu
@pytest.mark.parametrize("make", CASES + V2_CASES, ids=[c.__name__[5:] for c in CASES + V2_CASES])
def test_apply_cuda(make):
    L_ = fb._lib.load()
    assert L_.fluca_b200_is_host_emulation() == 0
    _check(L_, make)


@pytest.mark.gpu
@pytest.mark.parametrize("limiter", ["minmod", "superbee"])
def test_tutorial_tvd_operator_cuda(limiter):
    L_ = fb._lib.load()
    assert L_.fluca_b200_is_host_emulation() == 0
    _tvd_check(L_, limiter)
    _ex3_check(L_, limiter)
