"""ctypes front-end of the CPU oracle (oracle/fluca_oracle.h).

TEST INFRASTRUCTURE, NOT PRODUCT CODE: only tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py import this module.  The product package
(fluca_b200/) never does.  Parity status of the oracle (oracle/fluca_oracle.h): PINNED to the reference's own NS sources compiled
on a PETSc model (oracle/ref.py, tests/test_oracle_vs_reference.py) for the discretisation, right-hand side, ABF factors and the
step; UNPINNED for PETSc's solver arithmetic (absent from the image) and for the immersed-boundary section (no reference code).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass
from typing import Callable, Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libfluca_oracle.so")

BC_NONE, BC_VELOCITY, BC_PRESSURE_OUTLET, BC_PERIODIC, BC_SYMMETRY = range(5)

_BCFN = C.CFUNCTYPE(C.c_int, C.c_int, C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_void_p)


class _OrcBC(C.Structure):
    _fields_ = [
        ("type", C.c_int),
        ("velocity", _BCFN),
        ("ctx_velocity", C.c_void_p),
        ("pressure", _BCFN),
        ("ctx_pressure", C.c_void_p),
    ]


class OrcOptions(C.Structure):
    _fields_ = [
        ("mode", C.c_int),
        ("outer_rtol", C.c_double),
        ("outer_maxit", C.c_int),
        ("mom_rtol", C.c_double),
        ("schur_rtol", C.c_double),
        ("inner_maxit", C.c_int),
        ("ilu_blocks", C.c_int),
        ("exact_schur", C.c_int),
        ("quirk_bcg_scale", C.c_int),
        ("schur_ainv", C.c_int),
        ("upper_ainv", C.c_int),
    ]


class OrcStepInfo(C.Structure):
    _fields_ = [
        ("outer_its", C.c_int),
        ("mom_its", C.c_int),
        ("schur_its", C.c_int),
        ("abf_applies", C.c_int),
        ("converged", C.c_int),
        ("outer_rnorm0", C.c_double),
        ("outer_rnorm", C.c_double),
        ("nhist", C.c_int),
        ("hist", C.c_double * 512),
    ]


def build(force: bool = False) -> str:
    """Compile the oracle with the committed Makefile (gcc only)."""
    srcs = [os.path.join(_HERE, "src", f) for f in ("ns.c", "sparse.c", "sparse.h")] + [os.path.join(_HERE, "fluca_oracle.h")]
    stale = force or not os.path.exists(_LIB_PATH) or any(os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs)
    if stale:
        subprocess.run(["make", "-C", _HERE], check=True, stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.POINTER(C.c_double)), C.c_double, C.c_double, C.c_double, C.POINTER(_OrcBC)]
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_sizes.argtypes = [C.c_void_p, C.POINTER(C.c_long), C.POINTER(C.c_long)]
        L.orc_set_state.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p), C.c_void_p, C.c_void_p, C.c_int, C.c_double]
        L.orc_get_state.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p), C.c_void_p, C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_double)]
        L.orc_step.restype = C.c_int
        L.orc_step.argtypes = [C.c_void_p, C.POINTER(OrcOptions), C.POINTER(OrcStepInfo)]
        L.orc_default_options.argtypes = [C.POINTER(OrcOptions)]
        L.orc_matrix.restype = C.c_int
        L.orc_matrix.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_long), C.POINTER(C.POINTER(C.c_int)), C.POINTER(C.POINTER(C.c_int)), C.POINTER(C.POINTER(C.c_double))]
        L.orc_prepare_step.argtypes = [C.c_void_p, C.POINTER(OrcOptions), C.c_void_p]
        L.orc_abf_apply.argtypes = [C.c_void_p, C.POINTER(OrcOptions), C.c_void_p, C.c_void_p, C.POINTER(OrcStepInfo)]
        L.orc_formula.restype = C.c_int
        L.orc_formula.argtypes = [C.c_char_p, C.POINTER(C.c_double), C.c_double, C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_int)]
        L.orc_set_markers.argtypes = [C.c_void_p, C.c_long, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.orc_set_ibm_iterations.argtypes = [C.c_void_p, C.c_int]
        L.orc_ibm_interpolate.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_ibm_spread.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_get_marker_forces.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_set_t_outlet_quirk.argtypes = [C.c_int]
        L.orc_set_threads.argtypes = [C.c_int]
        L.orc_set_threads.restype = C.c_int
        L.orc_get_threads.restype = C.c_int
        _lib = L
    return _lib


def set_t_outlet_quirk(on: bool) -> None:
    """Objects created from now on form the 3-D upper-outlet rows of T as cnlinearcart3d.c:1996 does (default) or as the 2-D file does."""
    lib().orc_set_t_outlet_quirk(1 if on else 0)


def set_threads(n: int) -> int:
    """Force the OpenMP thread count of the oracle (overrides an inherited OMP_NUM_THREADS); returns the count a parallel
    region actually gets."""
    lib().orc_set_threads(int(n))
    return int(lib().orc_get_threads())


def default_options(**kw) -> OrcOptions:
    o = OrcOptions()
    lib().orc_default_options(C.byref(o))
    for k, v in kw.items():
        if not hasattr(o, k):
            raise AttributeError(k)
        setattr(o, k, v)
    return o


def formula(name: str, xs: Sequence[float], h: float = 1.0, vf: float = 1.0):
    """One of the 24 cartdiscret.c formulas: returns (offsets, weights)."""
    x = (C.c_double * 8)(*list(xs) + [0.0] * (8 - len(xs)))
    w = (C.c_double * 4)()
    off = (C.c_int * 4)()
    n = lib().orc_formula(name.encode(), x, h, vf, w, off)
    if n < 0:
        raise KeyError(name)
    return [off[i] for i in range(n)], [w[i] for i in range(n)]


@dataclass
class BC:
    """Mirror of NSBoundaryCondition (fluca/include/flucansbc.h:16-22).

    velocity(dim, t, x) -> sequence of dim values; pressure(dim, t, x) -> float.
    """

    type: int
    velocity: Optional[Callable] = None
    pressure: Optional[Callable] = None
    # constants evaluated by the oracle's built-in C callbacks (timing runs: no Python in the loop)
    const_velocity: Optional[Sequence[float]] = None
    const_pressure: Optional[float] = None


class Oracle:
    def __init__(self, n: Sequence[int], xf: Sequence[np.ndarray], rho: float, mu: float, dt: float, bcs: Sequence[BC], periodic: Optional[Sequence[bool]] = None):
        L = lib()
        self.dim = len(n)
        assert self.dim in (2, 3) and len(bcs) == 2 * self.dim and len(xf) == self.dim
        self.n = tuple(int(a) for a in n) + (1,) * (3 - self.dim)
        if periodic is None:
            periodic = [bcs[2 * d].type == BC_PERIODIC for d in range(self.dim)]
        self.periodic = tuple(bool(a) for a in periodic) + (False,) * (3 - self.dim)
        self._keep = []
        cbcs = (_OrcBC * 6)()
        for b in range(6):
            if b >= 2 * self.dim:
                cbcs[b].type = BC_NONE
                continue
            bc = bcs[b]
            cbcs[b].type = bc.type
            if bc.const_velocity is not None:
                arr = (C.c_double * 3)(*(list(bc.const_velocity) + [0.0] * 3)[:3])
                self._keep.append(arr)
                cbcs[b].velocity = C.cast(L.orc_bc_constant, _BCFN)
                cbcs[b].ctx_velocity = C.cast(arr, C.c_void_p)
            elif bc.velocity is not None:
                cbcs[b].velocity = self._wrap_velocity(bc.velocity)
            if bc.const_pressure is not None:
                arr = (C.c_double * 1)(float(bc.const_pressure))
                self._keep.append(arr)
                cbcs[b].pressure = C.cast(L.orc_bc_constant_pressure, _BCFN)
                cbcs[b].ctx_pressure = C.cast(arr, C.c_void_p)
            elif bc.pressure is not None:
                cbcs[b].pressure = self._wrap_pressure(bc.pressure)
        self._xf = [np.ascontiguousarray(a, dtype=np.float64) for a in xf]
        for d in range(self.dim):
            assert self._xf[d].shape == (self.n[d] + 1,)
        cn = (C.c_int * 3)(*self.n)
        cper = (C.c_int * 3)(*[int(a) for a in self.periodic])
        cxf = (C.POINTER(C.c_double) * 3)()
        for d in range(self.dim):
            cxf[d] = self._xf[d].ctypes.data_as(C.POINTER(C.c_double))
        self._h = L.orc_create(self.dim, cn, cper, cxf, rho, mu, dt, cbcs)
        nc = C.c_long()
        nf = (C.c_long * 3)()
        L.orc_sizes(self._h, C.byref(nc), nf)
        self.ncell = nc.value
        self.nface = [nf[d] for d in range(3)]
        self.nsol = self.dim * self.ncell + sum(self.nface) + self.ncell
        self.face_shape = []
        for d in range(self.dim):
            shp = [self.n[2], self.n[1], self.n[0]]
            shp[2 - d] += 0 if self.periodic[d] else 1
            self.face_shape.append(tuple(shp))
        self.cell_shape = (self.n[2], self.n[1], self.n[0])

    def _wrap_velocity(self, fn):
        dim = self.dim

        def cb(d, t, x, val, ctx):
            out = fn(d, t, [x[i] for i in range(dim)])
            for i in range(dim):
                val[i] = out[i]
            return 0

        c = _BCFN(cb)
        self._keep.append(c)
        return c

    def _wrap_pressure(self, fn):
        dim = self.dim

        def cb(d, t, x, val, ctx):
            val[0] = float(fn(d, t, [x[i] for i in range(dim)]))
            return 0

        c = _BCFN(cb)
        self._keep.append(c)
        return c

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                lib().orc_destroy(self._h)
                self._h = None
        except Exception:
            pass

    # ---- state: v has shape (dim, nz, ny, nx); U[d] has face_shape[d]; p (nz, ny, nx) ----
    def set_state(self, v, U, p, phalf=None, step=0, t=0.0):
        v = np.ascontiguousarray(v, dtype=np.float64).reshape(self.dim, *self.cell_shape)
        p = np.ascontiguousarray(p, dtype=np.float64).reshape(self.cell_shape)
        Us = [np.ascontiguousarray(U[d], dtype=np.float64).reshape(self.face_shape[d]) for d in range(self.dim)]
        cU = (C.c_void_p * 3)()
        for d in range(self.dim):
            cU[d] = Us[d].ctypes.data
        ph = None
        if phalf is not None:
            ph = np.ascontiguousarray(phalf, dtype=np.float64).reshape(self.cell_shape)
        lib().orc_set_state(self._h, v.ctypes.data, cU, p.ctypes.data, ph.ctypes.data if ph is not None else None, int(step), float(t))

    def get_state(self):
        v = np.empty((self.dim,) + self.cell_shape)
        p = np.empty(self.cell_shape)
        ph = np.empty(self.cell_shape)
        Us = [np.empty(self.face_shape[d]) for d in range(self.dim)]
        cU = (C.c_void_p * 3)()
        for d in range(self.dim):
            cU[d] = Us[d].ctypes.data
        step = C.c_int()
        t = C.c_double()
        lib().orc_get_state(self._h, v.ctypes.data, cU, p.ctypes.data, ph.ctypes.data, C.byref(step), C.byref(t))
        return dict(v=v, U=Us, p=p, phalf=ph, step=step.value, t=t.value)

    def step(self, opt: Optional[OrcOptions] = None) -> OrcStepInfo:
        opt = opt or default_options()
        info = OrcStepInfo()
        lib().orc_step(self._h, C.byref(opt), C.byref(info))
        return info

    def prepare_step(self, opt: Optional[OrcOptions] = None) -> np.ndarray:
        opt = opt or default_options()
        rhs = np.empty(self.nsol)
        lib().orc_prepare_step(self._h, C.byref(opt), rhs.ctypes.data)
        return rhs

    def abf_apply(self, b: np.ndarray, opt: Optional[OrcOptions] = None):
        opt = opt or default_options()
        b = np.ascontiguousarray(b, dtype=np.float64)
        x = np.zeros(self.nsol)
        info = OrcStepInfo()
        lib().orc_abf_apply(self._h, C.byref(opt), b.ctypes.data, x.ctypes.data, C.byref(info))
        return x, info

    def matrix(self, name: str):
        """CSR copy of one of the assembled operators as a scipy.sparse.csr_matrix."""
        import scipy.sparse as sp

        nr, nc, nnz = C.c_int(), C.c_int(), C.c_long()
        ptr, idx = C.POINTER(C.c_int)(), C.POINTER(C.c_int)()
        val = C.POINTER(C.c_double)()
        rc = lib().orc_matrix(self._h, name.encode(), C.byref(nr), C.byref(nc), C.byref(nnz), C.byref(ptr), C.byref(idx), C.byref(val))
        if rc != 0:
            raise KeyError(name)
        p = np.ctypeslib.as_array(ptr, shape=(nr.value + 1,)).copy()
        i = np.ctypeslib.as_array(idx, shape=(max(nnz.value, 1),))[: nnz.value].copy()
        v = np.ctypeslib.as_array(val, shape=(max(nnz.value, 1),))[: nnz.value].copy()
        return sp.csr_matrix((v, i, p), shape=(nr.value, nc.value))

    # ---- immersed boundary (defined by the oracle itself: the reference has none; parity unpinned) ----
    def set_markers(self, X, Ud, dV, npts=4, iterations=1):
        """X, Ud: (dim, n); dV: (n,)."""
        X = np.ascontiguousarray(X, dtype=np.float64).reshape(self.dim, -1)
        Ud = np.ascontiguousarray(Ud, dtype=np.float64).reshape(self.dim, -1)
        dV = np.ascontiguousarray(dV, dtype=np.float64).ravel()
        self.nm = X.shape[1]
        assert Ud.shape == X.shape and dV.shape == (self.nm,)
        lib().orc_set_markers(self._h, self.nm, X.ctypes.data, Ud.ctypes.data, dV.ctypes.data, int(npts))
        lib().orc_set_ibm_iterations(self._h, int(iterations))

    def ibm_interpolate(self, v):
        v = np.ascontiguousarray(v, dtype=np.float64).reshape(self.dim, *self.cell_shape)
        Um = np.zeros((self.dim, self.nm))
        lib().orc_ibm_interpolate(self._h, v.ctypes.data, Um.ctypes.data)
        return Um

    def ibm_spread(self, Fm):
        Fm = np.ascontiguousarray(Fm, dtype=np.float64).reshape(self.dim, self.nm)
        f = np.zeros((self.dim,) + self.cell_shape)
        lib().orc_ibm_spread(self._h, Fm.ctypes.data, f.ctypes.data)
        return f

    def marker_forces(self):
        F, Um = np.zeros((self.dim, self.nm)), np.zeros((self.dim, self.nm))
        lib().orc_get_marker_forces(self._h, F.ctypes.data, Um.ctypes.data)
        return F, Um

    # split / join a solution-sized vector into (v, [U_d], p) views
    def split(self, x: np.ndarray):
        nv = self.dim * self.ncell
        v = x[:nv].reshape(self.dim, *self.cell_shape)
        U, off = [], nv
        for d in range(self.dim):
            U.append(x[off : off + self.nface[d]].reshape(self.face_shape[d]))
            off += self.nface[d]
        p = x[off:].reshape(self.cell_shape)
        return v, U, p


def uniform_faces(n: int, lo: float, hi: float) -> np.ndarray:
    """Face coordinates of MeshCartSetUniformCoordinates (cart.c:458-465)."""
    return lo + (hi - lo) * np.arange(n + 1, dtype=np.float64) / n
