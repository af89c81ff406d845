"""Asynchronous solution views (fluca_b200_stage_state / fluca_b200_staged_state): the staged pinned-host
copy must be exactly the state at the moment of staging, whatever the solver does between staging and the host's wait --
on the GPU the copy runs on its own stream while the next steps compute."""
import numpy as np
import pytest

import fluca_b200 as fb
from tests import cases, parity


def _same(a, b):
    return all(np.array_equal(a[k], b[k]) for k in ("v", "p", "phalf")) and all(np.array_equal(x, y) for x, y in zip(a["U"], b["U"]))


def _check(lib, case, mode, steps_between):
    ns = parity.make_ns(case, lib, mode)
    parity.set_initial(ns, case.initial_state(seed=2))
    s = fb.NSB200GetSolver(ns)
    with pytest.raises(fb._lib.FlucaB200Error):
        s.staged_state()  # nothing staged yet
    fb.NSStep(ns)
    ref = s.get_state()
    s.stage_state()
    for _ in range(steps_between):  # the time loop goes on while the copy is in flight
        fb.NSStep(ns)
    got = s.staged_state()
    assert _same(got, ref)
    assert not _same(s.get_state(), ref) or steps_between == 0
    # the buffers are the library's own and stay put until the next staging
    a, b = s.staged_state(copy=False), s.staged_state(copy=False)
    assert a["v"].ctypes.data == b["v"].ctypes.data and _same(a, ref)
    # staging again replaces them with the newer state; overwriting the live state (set_state) must not disturb a copy in flight
    now = s.get_state()
    s.stage_state()
    s.set_state(v=np.zeros_like(now["v"]), p=np.ones_like(now["p"]))
    assert _same(s.staged_state(), now)
    fb.NSDestroy(ns)


CASES = [
    ("cavity2d", lambda: cases.cavity2d(n=16), "coupled", 2),
    ("cavity3d_sym", lambda: cases.cavity3d(n=(8, 8, 4)), "fractional", 3),
    ("channel3d_outlet", lambda: cases.channel3d(n=(8, 6, 6), pout=0.2, dt=0.05), "coupled", 1),
    ("no_steps_between", lambda: cases.cavity3d_full(n=(8, 6, 5)), "fractional", 0),
]


@pytest.mark.parametrize("name,mk,mode,between", CASES, ids=[c[0] for c in CASES])
def test_staged_view_host_emulation(name, mk, mode, between):
    _check(parity.hostemu_library(), mk(), mode, between)


def test_ns_api_stage_and_sync():
    """The application-level pair of glue/nsb200.c: stage after step n, keep stepping, sync before the viewer runs."""
    case = cases.cavity2d(n=12)
    ns = parity.make_ns(case, parity.hostemu_library(), "fractional")
    parity.set_initial(ns, case.initial_state())
    fb.NSStep(ns)
    want = {k: fb.NSGetSolutionSubVector(ns, k) for k in (fb.NS_FIELD_VELOCITY, fb.NS_FIELD_PRESSURE)}
    fb.NSB200StageSolution(ns)
    step_staged = fb.NSGetTimeStep(ns)
    fb.NSStep(ns)
    fb.NSStep(ns)
    got, (step, t) = fb.NSB200SyncSolution(ns)
    assert step == step_staged and np.array_equal(got[fb.NS_FIELD_VELOCITY], want[fb.NS_FIELD_VELOCITY]) and np.array_equal(got[fb.NS_FIELD_PRESSURE], want[fb.NS_FIELD_PRESSURE])
    # without an earlier staging the sync stages the current state itself
    got, (step, t) = fb.NSB200SyncSolution(ns)
    assert step == fb.NSGetTimeStep(ns) and np.array_equal(got[fb.NS_FIELD_VELOCITY], fb.NSGetSolutionSubVector(ns, fb.NS_FIELD_VELOCITY))
    fb.NSDestroy(ns)


GPU_CASES = [
    ("cavity2d_64", lambda: cases.cavity2d(n=64), "coupled", 2),
    # large enough for the copy (7 fields of 2.4 MB) to still be in flight when the next steps start
    ("tma_cavity3d_96x64x48", lambda: cases.cavity3d_full(n=(96, 64, 48)), "fractional", 3),
    ("channel3d_outlet", lambda: cases.channel3d(n=(40, 16, 10), pout=0.2, dt=0.05), "coupled", 1),
]


@pytest.mark.gpu
@pytest.mark.parametrize("name,mk,mode,between", GPU_CASES, ids=[c[0] for c in GPU_CASES])
def test_staged_view_cuda(name, mk, mode, between):
    L = fb._lib.load()
    assert L.fluca_b200_is_host_emulation() == 0
    _check(L, mk(), mode, between)
