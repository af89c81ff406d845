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
HOSTEMU = os.path.join(ROOT, "tests", "hostemu", "_build", "libfluca_b200_hostemu.so")


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


def make_ns(case, library=None, mode="coupled", comm=None, **opts):
    """Build the NS object the way the reference drivers do (cavity_flow_2d.c:38-71)."""
    bnd = [fb.MESHCART_BOUNDARY_PERIODIC if p else fb.MESHCART_BOUNDARY_NONE for p in case.periodic()]
    if case.dim == 2:
        mesh = fb.MeshCartCreate2d(None, bnd[0], bnd[1], *case.n)
    else:
        mesh = fb.MeshCartCreate3d(None, bnd[0], bnd[1], bnd[2], *case.n)
    fb.MeshSetUp(mesh)
    fb.MeshCartSetCoordinates(mesh, case.faces())
    ns = fb.NSCreate(comm)
    fb.NSSetType(ns, fb.NSB200)
    if library is not None:
        fb.NSB200SetLibrary(ns, library)
    fb.NSSetMesh(ns, mesh)
    fb.NSSetDensity(ns, case.rho)
    fb.NSSetViscosity(ns, case.mu)
    for b, bc in enumerate(case.bcs):
        fb.NSSetBoundaryCondition(ns, b, fb.NSBoundaryCondition(type=bc["type"], velocity=bc["velocity"], pressure=bc["pressure"]))
    o = {"ns_time_step_size": case.dt, "ns_b200_mode": mode}
    o.update(opts)
    fb.NSSetFromOptions(ns, o)
    fb.NSSetUp(ns)
    return ns


TIGHT = {"ns_ksp_rtol": 1e-13, "ns_abf_momentum_ksp_rtol": 1e-13, "ns_abf_schur_ksp_rtol": 1e-13, "ns_ksp_max_it": 60}
ORC_TIGHT = dict(outer_rtol=1e-13, mom_rtol=1e-13, schur_rtol=1e-13)


def set_initial(ns, state):
    v, U, p = state
    fb.NSSetSolutionSubVector(ns, fb.NS_FIELD_VELOCITY, v)
    fb.NSSetSolutionSubVector(ns, fb.NS_FIELD_FACE_NORMAL_VELOCITY, U)
    fb.NSSetSolutionSubVector(ns, fb.NS_FIELD_PRESSURE, p)


AINV_OPTION = {0: "ID", 1: "DIAG", 2: "ROWSUM"}  # PCABFAinvTypes[], abfpc.c:4


def compare_steps(case, library=None, mode="coupled", nsteps=2, seed=None, tol=1e-10, ptol=None, markers=None, ainv=(0, 0)):
    """K steps on both sides at tight tolerances; returns the per-step relative L2 differences.
    ainv = (schur, upper) PCABFAinvType of both sides (-ns_pc_abf_{schur,upper}_ainv_type)."""
    orc = cases.make_oracle(case)
    state = case.initial_state(seed=seed)
    orc.set_state(*state)
    ns = make_ns(case, library, mode, ns_pc_abf_schur_ainv_type=AINV_OPTION[ainv[0]], ns_pc_abf_upper_ainv_type=AINV_OPTION[ainv[1]], **TIGHT)
    set_initial(ns, state)
    if markers is not None:  # immersed boundary: the same marker list on both sides
        orc.set_markers(markers["X"], markers["Ud"], markers["dV"], markers.get("npts", 4), markers.get("iterations", 1))
        fb.NSB200SetMarkers(ns, markers["X"], markers["Ud"], markers["dV"], markers.get("npts", 4), markers.get("iterations", 1))
    out = []
    oopt = O.default_options(mode=0 if mode == "coupled" else 1, schur_ainv=ainv[0], upper_ainv=ainv[1], **ORC_TIGHT)
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
