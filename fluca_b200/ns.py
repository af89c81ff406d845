"""Host-side mirror of the reference's NS / MeshCart interface for the time-step path.

The reference's host language is C on PETSc; PETSc is not in this image, so the PETSc-facing
glue (glue/nsb200.c, shown in INTEGRATION.md) cannot be compiled here.  This module is the same
host logic written above the C ABI in Python, with the reference's names, argument meaning and
error behaviour, so that the parity tests read like the reference's own drivers
(fluca/tests/cavity_flow/cavity_flow_2d.c, fluca/tests/taylor_green_vortex/taylor_green_vortex.c):

    mesh = MeshCartCreate3d(...); MeshSetUp(mesh); MeshCartSetUniformCoordinates(mesh, ...)
    ns = NSCreate(); NSSetType(ns, NSB200); NSSetMesh(ns, mesh); NSSetDensity(ns, rho) ...
    NSSetBoundaryCondition(ns, ileftb, NSBoundaryCondition(NS_BC_VELOCITY, velocity=f))
    NSSetFromOptions(ns, {"ns_time_step_size": dt, "ns_max_steps": 10}); NSSetUp(ns); NSSolve(ns)

It holds no arithmetic: fields live on the GPU, every operator is a CUDA kernel behind
include/fluca_b200.h.  What stays on the host is what the reference keeps on the host:
options, the time loop (nsbasic.c:325-351), monitors (nsmon.c), failure policy
(nsbasic.c:288-297) and the evaluation of the user's boundary callbacks (flucansbc.h:14).
"""
from __future__ import annotations

import math
import sys
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Sequence

import numpy as np

from . import _lib
from .solver import (BC_NONE, BC_PERIODIC, BC_PRESSURE_OUTLET, BC_SYMMETRY, BC_VELOCITY, MODE_COUPLED, MODE_FRACTIONAL, Comm, Solver, slab_partition)

# ---- enums / names of the reference (flucansbc.h:5-11, flucans.h:11-23, flucameshcart.h:5-19) ----
NS_BC_NONE, NS_BC_VELOCITY, NS_BC_PRESSURE_OUTLET, NS_BC_PERIODIC, NS_BC_SYMMETRY = BC_NONE, BC_VELOCITY, BC_PRESSURE_OUTLET, BC_PERIODIC, BC_SYMMETRY
MESHCART_BOUNDARY_NONE, MESHCART_BOUNDARY_PERIODIC = 0, 1
MESHCART_LEFT, MESHCART_RIGHT, MESHCART_DOWN, MESHCART_UP, MESHCART_BACK, MESHCART_FRONT = range(6)
NS_CONVERGED_ITERATING, NS_CONVERGED_TIME, NS_CONVERGED_ITS, NS_DIVERGED_NONLINEAR_SOLVE = 0, 1, 2, -1
NS_FIELD_VELOCITY, NS_FIELD_FACE_NORMAL_VELOCITY, NS_FIELD_PRESSURE = "Velocity", "FaceNormalVelocity", "Pressure"
NSB200 = "b200"  # the new type name, next to NSCNLINEAR "cnlinear" (flucans.h:11)
PETSC_DECIDE = -1


class FlucaError(RuntimeError):
    """Stands in for a non-zero PetscErrorCode."""


# ------------------------------------------------------------------ MeshCart (the part NS reads)
@dataclass
class Mesh:
    dim: int
    n: List[int]
    periodic: List[bool]
    ranks_z: int = 1
    xf: Optional[List[np.ndarray]] = None
    setupcalled: bool = False


def MeshCartCreate2d(comm, bndx, bndy, M, N, m=PETSC_DECIDE, n=PETSC_DECIDE, lx=None, ly=None) -> Mesh:
    return Mesh(2, [int(M), int(N)], [bndx == MESHCART_BOUNDARY_PERIODIC, bndy == MESHCART_BOUNDARY_PERIODIC])


def MeshCartCreate3d(comm, bndx, bndy, bndz, M, N, P, m=PETSC_DECIDE, n=PETSC_DECIDE, p=PETSC_DECIDE, lx=None, ly=None, lz=None) -> Mesh:
    return Mesh(3, [int(M), int(N), int(P)], [b == MESHCART_BOUNDARY_PERIODIC for b in (bndx, bndy, bndz)])


def MeshSetFromOptions(mesh: Mesh, options: Optional[Dict] = None):
    o = options or {}
    for d, key in enumerate(("cart_grid_x", "cart_grid_y", "cart_grid_z")[: mesh.dim]):  # cart.c:22-42
        if key in o:
            mesh.n[d] = int(o[key])


def MeshSetUp(mesh: Mesh):
    mesh.setupcalled = True


def MeshCartSetUniformCoordinates(mesh: Mesh, xmin, xmax, ymin, ymax, zmin=0.0, zmax=0.0):
    if not mesh.setupcalled:  # cart.c:462
        raise FlucaError("This function must be called after MeshSetUp()")
    lims = [(xmin, xmax), (ymin, ymax), (zmin, zmax)][: mesh.dim]
    mesh.xf = [lo + (hi - lo) * np.arange(n + 1, dtype=np.float64) / n for (lo, hi), n in zip(lims, mesh.n)]


def MeshCartSetCoordinates(mesh: Mesh, faces: Sequence[np.ndarray]):
    """Non-uniform product coordinates (what a CGNS-loaded mesh provides, cart.c:133-140)."""
    mesh.xf = [np.ascontiguousarray(f, dtype=np.float64) for f in faces]


def MeshCartGetBoundaryIndex(mesh: Mesh, loc: int) -> int:  # cart.c:564-591
    if not 0 <= loc < 6:
        raise FlucaError("Invalid boundary location")
    return loc


def MeshDestroy(mesh):
    return None


# ------------------------------------------------------------------ boundary conditions
@dataclass
class NSBoundaryCondition:
    """flucansbc.h:16-22.  velocity(dim, t, x, ctx) -> dim values; pressure(dim, t, x, ctx) -> value.

    A callback may set attribute `vectorized = True`: it is then called once per boundary with x a
    list of coordinate arrays and must return arrays (host evaluation of a 512^2 plane point by point
    in Python would dominate the step).
    """

    type: int = NS_BC_NONE
    velocity: Optional[Callable] = None
    ctx_velocity: object = None
    pressure: Optional[Callable] = None
    ctx_pressure: object = None


def constant_velocity(*vals):
    def f(dim, t, x, ctx=None):
        shape = np.shape(x[0])
        return [np.full(shape, v) if shape else v for v in vals[:dim]]

    f.vectorized = True
    f.time_independent = True
    return f


def constant_pressure(val):
    def f(dim, t, x, ctx=None):
        shape = np.shape(x[0])
        return np.full(shape, val) if shape else val

    f.vectorized = True
    f.time_independent = True
    return f


def _call(fn, dim, t, x, ctx):
    nargs = getattr(fn, "_fluca_nargs", None)
    if nargs is None:
        import inspect

        try:
            nargs = len(inspect.signature(fn).parameters)
        except (TypeError, ValueError):
            nargs = 4
        try:
            fn._fluca_nargs = nargs
        except AttributeError:
            pass
    return fn(dim, t, x, ctx) if nargs >= 4 else fn(dim, t, x)


# ------------------------------------------------------------------ NS object
_NSList: Dict[str, Callable] = {}


def NSRegister(type_name: str, create: Callable):  # nsreg.c:5-11
    _NSList[type_name] = create


@dataclass
class NS:
    rho: float = 0.0
    mu: float = 0.0
    dt: float = 0.0
    max_time: float = math.inf
    max_steps: int = 2**31 - 1
    step: int = 0
    t: float = 0.0
    mesh: Optional[Mesh] = None
    bcs: List[NSBoundaryCondition] = field(default_factory=list)
    type_name: Optional[str] = None
    ops: Dict[str, Callable] = field(default_factory=dict)
    data: object = None
    errorifstepfailed: bool = True
    reason: int = NS_CONVERGED_ITERATING
    setupcalled: bool = False
    monitors: List[Callable] = field(default_factory=list)
    options: Dict = field(default_factory=dict)
    comm: object = None  # None, or a dict(rank=, nranks=, make_comm=callable) for multi-GPU


def NSCreate(comm=None) -> NS:  # nsbasic.c:19-53
    return NS(comm=comm)


def NSSetType(ns: NS, type_name: str):  # nsbasic.c:55-79
    if ns.type_name == type_name:
        return
    create = _NSList.get(type_name)
    if create is None:
        raise FlucaError(f"Unknown ns type: {type_name}")
    if ns.type_name and "destroy" in ns.ops:
        ns.ops["destroy"](ns)
    ns.ops = {}
    ns.type_name = type_name
    create(ns)


def NSSetMesh(ns, mesh):
    ns.mesh = mesh
    ns.bcs = [NSBoundaryCondition() for _ in range(2 * mesh.dim)]
    for d in range(mesh.dim):  # periodic mesh directions carry NS_BC_PERIODIC (nsopts.c)
        if mesh.periodic[d]:
            ns.bcs[2 * d].type = ns.bcs[2 * d + 1].type = NS_BC_PERIODIC


def NSSetDensity(ns, rho):
    ns.rho = float(rho)


def NSSetViscosity(ns, mu):
    ns.mu = float(mu)


def NSSetTimeStepSize(ns, dt):
    ns.dt = float(dt)


def NSSetMaxSteps(ns, n):
    ns.max_steps = int(n)


def NSSetMaxTime(ns, t):
    ns.max_time = float(t)


def NSSetTimeStep(ns, step):
    ns.step = int(step)


def NSSetTime(ns, t):
    ns.t = float(t)


def NSGetTime(ns):
    return ns.t


def NSGetTimeStep(ns):
    return ns.step


def NSGetConvergedReason(ns):
    return ns.reason


def NSSetErrorIfStepFailed(ns, flg):
    ns.errorifstepfailed = bool(flg)


def NSSetBoundaryCondition(ns, index, bc: NSBoundaryCondition):
    if ns.mesh is None:
        raise FlucaError("Mesh not set")
    if not 0 <= index < 2 * ns.mesh.dim:
        raise FlucaError("Boundary index out of range")
    ns.bcs[index] = bc


def NSSetFromOptions(ns, options: Optional[Dict] = None):  # nsopts.c:179-194
    o = dict(options or {})
    ns.options.update(o)
    if "ns_type" in o:
        NSSetType(ns, o["ns_type"])
    elif ns.type_name is None:
        NSSetType(ns, NSB200)
    if "ns_density" in o:
        ns.rho = float(o["ns_density"])
    if "ns_viscosity" in o:
        ns.mu = float(o["ns_viscosity"])
    if "ns_time_step_size" in o:
        ns.dt = float(o["ns_time_step_size"])
    if "ns_max_time" in o:
        ns.max_time = float(o["ns_max_time"])
    if "ns_max_steps" in o:
        ns.max_steps = int(o["ns_max_steps"])
    if "ns_error_if_step_failed" in o:
        ns.errorifstepfailed = bool(int(o["ns_error_if_step_failed"]))
    if o.get("ns_monitor"):
        NSMonitorSet(ns, NSMonitorDefault)
    if "setfromoptions" in ns.ops:
        ns.ops["setfromoptions"](ns, o)


def NSMonitorSet(ns, fn):  # nsmon.c:4-20 (MAXNSMONITORS = 10, nsimpl.h:9)
    if len(ns.monitors) >= 10:
        raise FlucaError("Too many monitors set")
    ns.monitors.append(fn)


def NSMonitorCancel(ns):
    ns.monitors.clear()


def NSMonitorDefault(ns):  # nsmon.c: "<step> NS dt <dt> time <t>"
    print(f"{ns.step} NS dt {ns.dt:g} time {ns.t:g}")


def NSMonitor(ns):
    for m in ns.monitors:
        m(ns)


def NSSetUp(ns):  # nsbasic.c:153-274
    if ns.setupcalled:
        return
    if ns.type_name is None:
        NSSetType(ns, NSB200)
    if ns.mesh is None:
        raise FlucaError("Mesh not set")
    ns.ops["setup"](ns)
    ns.setupcalled = True


def NSStep(ns):  # nsbasic.c:276-299
    ns.ops["step"](ns)
    if ns.reason >= 0:
        ns.step += 1
        ns.t += ns.dt
    if ns.reason < 0 and ns.errorifstepfailed:
        NSMonitorCancel(ns)
        raise FlucaError("NSStep has failed due to DIVERGED_NONLINEAR_SOLVE")


def NSSolve(ns):  # nsbasic.c:325-351
    if not (ns.max_time < math.inf or ns.max_steps != 2**31 - 1):
        raise FlucaError("At least one of max time or max steps must be specified")
    if ns.step >= ns.max_steps:
        ns.reason = NS_CONVERGED_ITS
    elif ns.t >= ns.max_time:
        ns.reason = NS_CONVERGED_TIME
    while ns.reason == NS_CONVERGED_ITERATING:
        NSMonitor(ns)
        NSStep(ns)
        if ns.reason == NS_CONVERGED_ITERATING:
            if ns.step >= ns.max_steps:
                ns.reason = NS_CONVERGED_ITS
            elif ns.t >= ns.max_time:
                ns.reason = NS_CONVERGED_TIME
    NSMonitor(ns)


def NSGetSolutionSubVector(ns, name):
    """Host copy of a field of ns->sol (nssol.c:110-119): numpy array(s) in the C-ABI layout."""
    return ns.ops["getfield"](ns, name)


def NSSetSolutionSubVector(ns, name, value):
    """Write a field of ns->sol (what tests do with DMStagVecSetValuesStencil, taylor_green_vortex.c:113-178)."""
    ns.ops["setfield"](ns, name, value)


def NSViewSolution(ns, viewer: dict):
    """Collect every field + the type's extra state ("PressureHalfStep", cnlinear.c:146-152) into a dict."""
    for name in (NS_FIELD_VELOCITY, NS_FIELD_FACE_NORMAL_VELOCITY, NS_FIELD_PRESSURE):
        viewer[name] = NSGetSolutionSubVector(ns, name)
    viewer["step"], viewer["time"] = ns.step, ns.t
    ns.ops["viewsolution"](ns, viewer)


def NSLoadSolution(ns, viewer: dict):  # nssol.c:176-204
    if not ns.setupcalled:
        raise FlucaError("This function must be called after NSSetUp()")
    for name in (NS_FIELD_VELOCITY, NS_FIELD_FACE_NORMAL_VELOCITY, NS_FIELD_PRESSURE):
        NSSetSolutionSubVector(ns, name, viewer[name])
    ns.ops["loadsolution"](ns, viewer)
    ns.step, ns.t = int(viewer["step"]), float(viewer["time"])


def NSDestroy(ns):
    if ns is not None and "destroy" in ns.ops:
        ns.ops["destroy"](ns)
        ns.ops = {}


# ------------------------------------------------------------------ the "b200" type
class _B200Data:
    solver: Optional[Solver] = None
    library = None
    last_stats = None
    history: List = None
    bc_cache: Dict = None
    pts: List = None


def _boundary_points(mesh: Mesh, s: Solver, b: int):
    """Boundary-face centres of boundary b on this rank's slab, as dim coordinate arrays of bc_shape[b]
    (the xb[] of every bc.velocity / bc.pressure call in cnlinearcart{2,3}d.c)."""
    d, side = b // 2, b % 2
    xc = [(f[:-1] + f[1:]) / 2.0 for f in mesh.xf]
    if mesh.dim == 3:
        xc[2] = xc[2][s.k0 : s.k0 + s.nzl]
    wall = mesh.xf[d][-1] if side else mesh.xf[d][0]
    shape = s.bc_shape[b]
    tang = [a for a in range(mesh.dim) if a != d]  # ascending: fastest index first in memory is tang[0]
    coords = [None] * mesh.dim
    coords[d] = np.full(shape, wall)
    if mesh.dim == 2:
        coords[tang[0]] = np.broadcast_to(xc[tang[0]][None, :], shape).copy()
    else:
        coords[tang[0]] = np.broadcast_to(xc[tang[0]][None, :], shape).copy()
        coords[tang[1]] = np.broadcast_to(xc[tang[1]][:, None], shape).copy()
    return coords


def _eval(fn, ctx, dim, t, pts, ncomp):
    shape = pts[0].shape
    if getattr(fn, "vectorized", False):
        out = _call(fn, dim, t, pts, ctx)
        if ncomp == 1:
            return np.broadcast_to(np.asarray(out, dtype=np.float64), shape).copy()
        return np.stack([np.broadcast_to(np.asarray(o, dtype=np.float64), shape) for o in out]).copy()
    res = np.empty((ncomp,) + shape)
    it = np.nditer(pts[0], flags=["multi_index"])
    for _ in it:
        idx = it.multi_index
        x = [float(p[idx]) for p in pts]
        val = _call(fn, dim, t, x, ctx)
        if ncomp == 1:
            res[(0,) + idx] = float(val)
        else:
            for c in range(ncomp):
                res[(c,) + idx] = val[c]
    return res[0] if ncomp == 1 else res


def _b200_setfromoptions(ns: NS, o: Dict):
    pass  # keys are read in setup: ns_b200_mode, ns_ksp_rtol, ns_abf_momentum_ksp_rtol, ns_abf_schur_ksp_rtol, ...


def _b200_setup(ns: NS):
    mesh, o, dat = ns.mesh, ns.options, ns.data
    if mesh.xf is None:
        raise FlucaError("Mesh coordinates not set")
    for b, bc in enumerate(ns.bcs):
        if bc.type == NS_BC_NONE:
            raise FlucaError("Unsupported boundary condition type")
        if bc.type == NS_BC_VELOCITY and bc.velocity is None:
            raise FlucaError(f"boundary {b}: NS_BC_VELOCITY needs a velocity callback")
        if bc.type == NS_BC_PRESSURE_OUTLET and bc.pressure is None:
            raise FlucaError(f"boundary {b}: NS_BC_PRESSURE_OUTLET needs a pressure callback")
    if not ns.dt > 0:
        raise FlucaError("Time step size must be set (-ns_time_step_size)")
    mode = {"coupled": MODE_COUPLED, "fractional": MODE_FRACTIONAL}[o.get("ns_b200_mode", "coupled")]
    kw = {}
    for key, name in (("ns_ksp_rtol", "outer_rtol"), ("ns_abf_momentum_ksp_rtol", "mom_rtol"), ("ns_abf_schur_ksp_rtol", "schur_rtol")):
        if key in o:
            kw[name] = float(o[key])
    for key, name in (("ns_ksp_max_it", "outer_maxit"), ("ns_ksp_gmres_restart", "outer_restart"), ("ns_abf_ksp_max_it", "inner_maxit"), ("ns_b200_mg_nu1", "mg_nu1"), ("ns_b200_mg_nu2", "mg_nu2"), ("ns_b200_mg_coarse_sweeps", "mg_coarse_sweeps"), ("ns_b200_no_bcg_quirk", "no_bcg_quirk"), ("ns_b200_no_t_outlet_quirk", "no_t_outlet_quirk")):
        if key in o:
            kw[name] = int(o[key])
    inner = [int(o[k]) for k in ("ns_abf_momentum_ksp_max_it", "ns_abf_schur_ksp_max_it") if k in o]  # PCABF's two KSPs (abfpc.c:33-46): one limit here
    if inner:
        kw["inner_maxit"] = max(inner)
    rank, nranks, comm = 0, 1, None
    if ns.comm:
        rank, nranks = ns.comm["rank"], ns.comm["nranks"]
    k0, nzl = 0, 1
    if mesh.dim == 3:
        k0, nzl = slab_partition(mesh.n[2], nranks)[rank]
    elif nranks != 1:
        raise FlucaError("2-D meshes run on one rank (the slab partition is along z)")
    lib = dat.library if dat.library is not None else _lib.load()
    if ns.comm and nranks > 1:
        comm = ns.comm["make_comm"](lib)
    try:
        dat.solver = Solver(mesh.n, mesh.xf, [bc.type for bc in ns.bcs], ns.rho, ns.mu, ns.dt, mode=mode, k0=k0, nzl=nzl, comm=comm, library=lib, **kw)
    except _lib.FlucaB200Error as e:
        raise FlucaError(str(e)) from e
    # -ns_pc_abf_schur_ainv_type / -ns_pc_abf_upper_ainv_type <ID|DIAG|ROWSUM> (abfpc.c:246-247 under the "ns_" prefix of nssol.c:17)
    ainv = []
    for key in ("ns_pc_abf_schur_ainv_type", "ns_pc_abf_upper_ainv_type"):
        name = str(o.get(key, "ID")).lower()
        if name not in _lib.AINV_NAMES:
            raise FlucaError(f"-{key}: unknown A-inverse type {o[key]!r} (ID, DIAG, ROWSUM)")
        ainv.append(_lib.AINV_NAMES[name])
    if any(ainv):
        dat.solver.set_abf_ainv_types(*ainv)
    # -ns_abf_momentum_ksp_monitor / -ns_abf_schur_ksp_monitor (the KSPs of PCABF, abfpc.c:33-46): KSPMonitorResidual's lines on rank 0
    want = (bool(o.get("ns_abf_momentum_ksp_monitor")), bool(o.get("ns_abf_schur_ksp_monitor")))
    if any(want) and rank == 0:
        out = o.get("ns_monitor_file", sys.stdout)

        def _inner(which, it, rnorm):
            if want[which]:
                if it == 0:
                    print(f"    Residual norms for ns_abf_{'schur' if which else 'momentum'}_ solve.", file=out)
                print(f"    {it:3d} KSP Residual norm {rnorm:14.12e}", file=out)

        dat.solver.set_inner_monitor(_inner)
    dat.pts = [_boundary_points(mesh, dat.solver, b) for b in range(2 * mesh.dim)]
    dat.bc_cache = {}
    dat.history = []


def _upload_bcs(ns: NS):
    """Evaluate the user's callbacks on the host at the times the step needs and ship the planes
    (<= 6 n^2 (dim + 1) doubles per time level; unchanged planes are not re-sent)."""
    dat, s, dim = ns.data, ns.data.solver, ns.mesh.dim
    tq = ns.t if ns.step == 0 else ns.t - 0.5 * ns.dt
    for b, bc in enumerate(ns.bcs):
        if bc.type == NS_BC_VELOCITY:
            for slot, t in ((0, ns.t), (1, ns.t + ns.dt)):
                key = (b, "v", slot)
                if getattr(bc.velocity, "time_independent", False) and key in dat.bc_cache:
                    continue
                vals = _eval(bc.velocity, bc.ctx_velocity, dim, t, dat.pts[b], dim)
                old = dat.bc_cache.get(key)
                if old is None or not np.array_equal(old, vals):
                    s.set_boundary_velocity(b, slot, vals)
                    dat.bc_cache[key] = vals
        elif bc.type == NS_BC_PRESSURE_OUTLET:
            for slot, t in ((0, tq), (1, ns.t + 0.5 * ns.dt)):
                key = (b, "p", slot)
                if getattr(bc.pressure, "time_independent", False) and key in dat.bc_cache:
                    continue
                vals = _eval(bc.pressure, bc.ctx_pressure, dim, t, dat.pts[b], 1)
                old = dat.bc_cache.get(key)
                if old is None or not np.array_equal(old, vals):
                    s.set_boundary_pressure(b, slot, vals)
                    dat.bc_cache[key] = vals


def _b200_step(ns: NS):
    dat = ns.data
    _upload_bcs(ns)
    try:
        st = dat.solver.step(ns.t, ns.step)
    except _lib.FlucaB200Error as e:
        if e.code == _lib.ERR_DIVERGED:  # NSCheckDiverged, nsbasic.c:425-436
            ns.reason = NS_DIVERGED_NONLINEAR_SOLVE
            return
        raise FlucaError(str(e)) from e
    dat.last_stats = st
    dat.history.append((st.outer_its, st.mom_its, st.schur_its, [st.hist[i] for i in range(st.nhist)]))


def _b200_prepare(ns: NS):
    """formfunction analogue: build the right-hand side of the current step without solving."""
    _upload_bcs(ns)
    ns.data.solver.prepare_step(ns.t, ns.step)
    return ns.data.solver.get_rhs()


def _b200_getfield(ns: NS, name: str):
    st = ns.data.solver.get_state()
    return {NS_FIELD_VELOCITY: st["v"], NS_FIELD_FACE_NORMAL_VELOCITY: st["U"], NS_FIELD_PRESSURE: st["p"], "PressureHalfStep": st["phalf"]}[name]


def _b200_setfield(ns: NS, name: str, value):
    s = ns.data.solver
    if name == NS_FIELD_VELOCITY:
        s.set_state(v=value)
    elif name == NS_FIELD_FACE_NORMAL_VELOCITY:
        s.set_state(U=value)
    elif name == NS_FIELD_PRESSURE:
        s.set_state(p=value)
    elif name == "PressureHalfStep":
        s.set_state(phalf=value)
    else:
        raise FlucaError(f'Field "{name}" not found')  # nssol.c:83


def _b200_viewsolution(ns: NS, viewer: dict):
    viewer["PressureHalfStep"] = _b200_getfield(ns, "PressureHalfStep")


def _b200_loadsolution(ns: NS, viewer: dict):
    _b200_setfield(ns, "PressureHalfStep", viewer["PressureHalfStep"])


def _b200_destroy(ns: NS):
    if ns.data is not None and ns.data.solver is not None:
        ns.data.solver.close()
        ns.data.solver = None


def _b200_view(ns: NS):
    s = ns.data.solver
    return f"NS type b200: mode {'coupled' if s.mode == MODE_COUPLED else 'fractional'}, slab k0={s.k0} nzl={s.nzl}"


def _b200_formjacobian(ns: NS, *a):
    raise FlucaError("NS type b200 is matrix-free: the Jacobian blocks are never assembled (use fluca_b200_apply_* for operator access)")


def NSCreate_B200(ns: NS):
    """Create function of the new type (model: NSCreate_CNLinear, cnlinear.c:164-187): fills all nine ops."""
    ns.data = _B200Data()
    ns.ops = dict(
        setfromoptions=_b200_setfromoptions,
        setup=_b200_setup,
        step=_b200_step,
        formjacobian=_b200_formjacobian,
        formfunction=_b200_prepare,
        destroy=_b200_destroy,
        view=_b200_view,
        viewsolution=_b200_viewsolution,
        loadsolution=_b200_loadsolution,
        getfield=_b200_getfield,
        setfield=_b200_setfield,
    )


NSRegister(NSB200, NSCreate_B200)


def NSB200SetLibrary(ns: NS, library):
    """Test hook: bind the type to an explicitly loaded C-ABI library (the CPU tests pass the
    host-emulation test double here; the default is the CUDA product library)."""
    ns.data.library = library


def NSB200GetSolver(ns: NS) -> Solver:
    return ns.data.solver


def NSB200SetMarkers(ns: NS, X, Ud, dV, delta_points: int = 4, iterations: int = 1):
    """Immersed-boundary markers of the b200 type (no counterpart in the reference, which only plans IBM:
    README.md:14, THEORY_GUIDE.md:130-132).  Call after NSSetUp; call again whenever the body moves."""
    if not ns.setupcalled:
        raise FlucaError("This function must be called after NSSetUp()")
    ns.data.solver.set_markers(X, Ud, dV, delta_points, iterations)


def NSB200GetMarkerForces(ns: NS):
    """(F, Um): force of every marker on the fluid and the interpolated predictor velocity of the last step."""
    return ns.data.solver.marker_forces()


def PCABFSetSchurComplementAinvType(ns: NS, type_: int):
    """flucans.h:106 (abfpc.c:300-308), addressed through the NS object because the b200 type owns its ABF factors."""
    if not ns.setupcalled:
        raise FlucaError("This function must be called after NSSetUp()")
    s = ns.data.solver
    s.set_abf_ainv_types(int(type_), s.ainv_types[1])


def PCABFSetUpperTriangularAinvType(ns: NS, type_: int):
    """flucans.h:107 (abfpc.c:310-318)."""
    if not ns.setupcalled:
        raise FlucaError("This function must be called after NSSetUp()")
    s = ns.data.solver
    s.set_abf_ainv_types(s.ainv_types[0], int(type_))


def NSB200StageSolution(ns: NS):
    """Start an asynchronous download of the current state into pinned host memory; the time loop may go on
    (glue/nsb200.c NSB200StageSolution; fluca_b200_stage_state)."""
    if not ns.setupcalled:
        raise FlucaError("This function must be called after NSSetUp()")
    ns.data.solver.stage_state()
    ns.data.staged_step = (ns.step, ns.t)


def NSB200SyncSolution(ns: NS):
    """Wait for the staged download; returns ({field name: array}, (step, t)) of the staged state -- what the C glue
    unpacks into ns->sol before NSViewSolution."""
    if not ns.setupcalled:
        raise FlucaError("This function must be called after NSSetUp()")
    if getattr(ns.data, "staged_step", None) is None:
        NSB200StageSolution(ns)
    st = ns.data.solver.staged_state()
    tag, ns.data.staged_step = ns.data.staged_step, None
    return {NS_FIELD_VELOCITY: st["v"], NS_FIELD_FACE_NORMAL_VELOCITY: st["U"], NS_FIELD_PRESSURE: st["p"], "PressureHalfStep": st["phalf"]}, tag


def NSB200GetStats(ns: NS):
    return ns.data.last_stats
