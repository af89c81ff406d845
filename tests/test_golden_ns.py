"""Committed NS-step fixtures (tests/golden/ns_steps.npz, written by tests/golden/make_ns_golden.py from the oracle; the
reference stores none -- SURVEY.md F5, 8c).  CPU: the oracle still reproduces them (pins it against accidental change) and
the host logic of the product sources matches them; GPU: the CUDA library matches them without the oracle in the loop."""
import importlib.util
import os

import numpy as np
import pytest

import fluca_b200 as fb
from oracle import oracle as O
from tests import cases, parity

HERE = os.path.dirname(os.path.abspath(__file__))
G = np.load(os.path.join(HERE, "golden", "ns_steps.npz"))
_spec = importlib.util.spec_from_file_location("make_ns_golden", os.path.join(HERE, "golden", "make_ns_golden.py"))
_gen = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_gen)
FIX = _gen.fixtures()
NAMES = list(FIX)


def _inputs(name, dim):
    return G[f"{name}/in_v"], [G[f"{name}/in_U{d}"] for d in range(dim)], G[f"{name}/in_p"]


@pytest.mark.parametrize("name", NAMES)
def test_oracle_reproduces_the_committed_fixtures(name):
    case, seed, markers, ainv = FIX[name]
    for mode, tag in ((0, "coupled"), (1, "fractional")):
        state, rhs, out, its = _gen.run(case, seed, markers, mode, ainv=ainv)
        if tag == "coupled":
            assert np.array_equal(state[0], G[f"{name}/in_v"])  # the seeded inputs themselves
            assert parity.rel(rhs, G[f"{name}/rhs"]) < 1e-13
        assert parity.rel(out["v"], G[f"{name}/{tag}/v"]) < 1e-11 and parity.rel(out["p"], G[f"{name}/{tag}/p"]) < 1e-10
        assert list(its) == list(G[f"{name}/{tag}/outer_its"])


def _product_matches(lib, name, mode):
    case, seed, markers, ainv = FIX[name]
    ns = parity.make_ns(case, lib, mode, ns_pc_abf_schur_ainv_type=parity.AINV_OPTION[ainv[0]], ns_pc_abf_upper_ainv_type=parity.AINV_OPTION[ainv[1]], **parity.TIGHT)
    parity.set_initial(ns, _inputs(name, case.dim))
    if markers is not None:
        fb.NSB200SetMarkers(ns, markers["X"], markers["Ud"], markers["dV"], markers.get("npts", 4))
    for _ in range(2):
        fb.NSStep(ns)
    got = fb.NSB200GetSolver(ns).get_state()
    k = f"{name}/{mode}"
    assert parity.rel(got["v"], G[f"{k}/v"]) < 1e-10
    assert parity.relU(got["U"], [G[f"{k}/U{d}"] for d in range(case.dim)]) < 1e-10
    assert parity.rel(got["p"], G[f"{k}/p"]) < 1e-9 and parity.rel(got["phalf"], G[f"{k}/phalf"]) < 1e-9
    fb.NSDestroy(ns)


@pytest.mark.parametrize("mode", ["coupled", "fractional"])
@pytest.mark.parametrize("name", NAMES)
def test_hostlogic_matches_the_fixtures(name, mode):
    _product_matches(parity.hostemu_library(), name, mode)


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["coupled", "fractional"])
@pytest.mark.parametrize("name", NAMES)
def test_cuda_library_matches_the_fixtures(name, mode):
    L = fb._lib.load()
    assert L.fluca_b200_is_host_emulation() == 0
    _product_matches(L, name, mode)
