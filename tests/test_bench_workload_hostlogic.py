"""CPU twin of tests/test_bench_workload_parity.py: the same checks of bench.py's workload builders (sphere + outlet +
symmetry + immersed boundary; cavity) in coupled mode, small, on the host-emulation test double -- so that the harness and
the host logic of those paths are exercised without a GPU."""
import pytest

import fluca_b200 as fb  # noqa: F401
from tests import cases, parity
from tests.test_bench_workload_parity import channel_workload, sphere_workload


@pytest.fixture(scope="module")
def lib():
    return parity.hostemu_library()


def test_sphere_workload_tight_small(lib):
    case, mk = sphere_workload((24, 16, 16))
    out = parity.compare_steps(case, lib, mode="coupled", nsteps=2, tol=1e-10, markers=mk, state=cases.uniform_inflow_state(case), fast_oracle=True)
    parity.assert_histories_track(out)


def test_sphere_workload_default_tolerances_small(lib):
    case, mk = sphere_workload((24, 16, 16))
    out = parity.default_tolerance_check(case, lib, cases.uniform_inflow_state(case), markers=mk, nsteps=3, orc_steps=3)
    assert out["true_rel"] <= 1e-5


def test_cavity_workload_default_tolerances_small(lib):
    case = cases.cavity_bench_case(16, 16)
    out = parity.default_tolerance_check(case, lib, case.initial_state(), nsteps=3, orc_steps=3)
    assert out["true_rel"] <= 1e-5


def test_channel_workload_with_wrapping_markers_small(lib):
    case, mk = channel_workload((16, 8, 8))
    out = parity.compare_steps(case, lib, mode="coupled", nsteps=2, tol=1e-10, markers=mk, state=cases.uniform_inflow_state(case), fast_oracle=True)
    parity.assert_histories_track(out)
