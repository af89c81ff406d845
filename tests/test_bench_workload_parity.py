"""Parity on what bench.py actually runs: the workload builders of BASELINE configs 3 and 4
themselves -- sphere + pressure outlet + symmetry + immersed boundary, and the 3-D cavity -- in COUPLED mode,
(i) at tight tolerances against the oracle (fields <= 1e-10, outer residual histories equal), on sizes that put the TMA tile
    kernels and one / three tiled multigrid levels on the path;
(ii) at the reference's default tolerances with the inexact-Krylov rule on, as the bench runs them: the true residual formed
    with the oracle's assembled operators meets 1e-5 |b| and the fields are within 1e-4 of the tight answer
(reference: NSStep_CNLinear_Cart3d_Internal cnlinearcart3d.c:2807-2863, tolerances nssol.c:21-29)."""
import os

import pytest

import fluca_b200 as fb
from tests import cases, parity

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib():
    L = fb._lib.load()
    assert L.fluca_b200_is_host_emulation() == 0
    return L


def sphere_workload(n3):
    """bench.py's sphere case (config 4) at n3 cells: same domain, boundary set, dt = 0.5 h_x rule, markers scaled with the
    surface cell count as bench.py's CPU sample does."""
    nx, ny, nz = n3
    c = cases.sphere_bench_case(ny, nz)
    c.n = (nx, ny, nz)
    c.dt = 0.5 * 16.0 / nx
    mk = cases.sphere_markers((0.0, 0.0, 0.0), 1.0, max(64, int(100000 * (ny / 512.0) ** 2)), 16.0 / ny)
    return c, mk


@pytest.mark.parametrize("n3", [(48, 48, 48), (128, 32, 32)], ids=["48cube_1_tiled_level", "128x32x32_3_tiled_levels"])
def test_sphere_workload_tight_matches_oracle(lib, n3):
    case, mk = sphere_workload(n3)
    out = parity.compare_steps(case, lib, mode="coupled", nsteps=2, tol=1e-10, markers=mk, state=cases.uniform_inflow_state(case), fast_oracle=True, ilu_blocks=os.cpu_count() or 1)
    parity.assert_histories_track(out)


def test_cavity_workload_tight_matches_oracle(lib):
    """config 3's builder at 64 x 32 x 32 (two tiled multigrid levels), three coupled steps"""
    case = cases.cavity_bench_case(32, 32)
    case.n, case.dt = (64, 32, 32), 0.5 / 64
    out = parity.compare_steps(case, lib, mode="coupled", nsteps=3, tol=1e-10, fast_oracle=True, ilu_blocks=os.cpu_count() or 1)
    parity.assert_histories_track(out)


def test_sphere_workload_default_tolerances_64cube(lib):
    case, mk = sphere_workload((64, 64, 64))
    out = parity.default_tolerance_check(case, lib, cases.uniform_inflow_state(case), markers=mk, nsteps=3, orc_steps=3)
    print("sphere 64^3 at default tolerances:", out)


def test_cavity_workload_default_tolerances_64cube(lib):
    case = cases.cavity_bench_case(64, 64)
    out = parity.default_tolerance_check(case, lib, case.initial_state(), nsteps=3, orc_steps=3)
    print("cavity 64^3 at default tolerances:", out)


def channel_workload(n3, nspheres=2):
    """bench.py's config-5 builder (periodic x and z, no-slip walls in y, several spheres) at n3 cells; one sphere sits across the
    periodic x boundary so that marker supports wrap"""
    case = cases.channel_bench_case(n3, periodic_z=True)
    h = case.hi[0] / n3[0]
    Lx, Ly, Lz = case.hi
    centres = [(0.02 * Lx, 0.5 * Ly, 0.45 * Lz), (0.6 * Lx, 0.4 * Ly, 0.98 * Lz)][:nspheres]
    D = 0.25 * Ly
    mk = cases.multi_sphere_markers(centres, D, 150, h)
    return case, mk


@pytest.mark.parametrize("n3", [(64, 16, 8), (40, 12, 12)], ids=["tiled_64x16x8", "tiled_shifted_40x12x12"])
def test_channel_workload_with_wrapping_markers_matches_oracle(lib, n3):
    """BASELINE config 5's code path: periodic-x tile kernels + immersed-boundary supports that wrap in x and z, coupled mode"""
    case, mk = channel_workload(n3)
    out = parity.compare_steps(case, lib, mode="coupled", nsteps=2, tol=1e-10, markers=mk, state=cases.uniform_inflow_state(case), fast_oracle=True, ilu_blocks=os.cpu_count() or 1)
    parity.assert_histories_track(out)
