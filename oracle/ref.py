"""ctypes binding of oracle/_ref/libfluca_ref_ns.so -- TEST INFRASTRUCTURE ONLY.

That library is the REFERENCE's own Navier-Stokes discretisation (cartdiscret.c, cnlinear.c, cnlinearcart2d.c, cnlinearcart3d.c,
abfpc.c of thecasterian/fluca, compiled from /root/reference by `make -C oracle ref`) running on a single-rank model of the PETSc
API subset it uses (oracle/ref_model/).  It exists to check this repository's oracle against the reference's code: operators,
right-hand side, ABF application, and the state after K steps with the linear solves taken to convergence.  It is built only where
the reference tree is present; `available()` says whether it is here.  Nothing in the product imports this module."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional, Sequence

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_ref", "libfluca_ref_ns.so")
REFERENCE_TREE = "/root/reference/fluca"
_lib = None

BC_CB = C.CFUNCTYPE(C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double))
EXACT, ABF_ONCE, GMRES_ABF = 0, 1, 2  # SNESSolve modes of ref_driver.c


def available(build: bool = True) -> bool:
    """True if the library exists (it is built on demand where the reference tree is present)."""
    if build and os.path.isdir(REFERENCE_TREE):
        subprocess.run(["make", "-C", HERE, "ref"], check=True, stdout=subprocess.DEVNULL)
    return os.path.exists(LIB)


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB)
        L.ref_create.restype = C.c_void_p
        L.ref_create.argtypes = [C.c_int, C.POINTER(C.c_int), C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int), C.c_double, C.c_double, C.c_double, BC_CB]
        L.ref_sizes.argtypes = [C.c_void_p, C.POINTER(C.c_long), C.POINTER(C.c_long)]
        L.ref_set_state.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_double]
        L.ref_get_state.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_double)]
        L.ref_step.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_void_p]
        L.ref_last_rhs.argtypes = [C.c_void_p, C.c_void_p]
        L.ref_form_function.argtypes = [C.c_void_p, C.c_void_p]
        L.ref_abf_apply.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.ref_apply_jacobian.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_matrix.restype = C.c_long
        L.ref_matrix.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_destroy.argtypes = [C.c_void_p]
        L.ref_set_inner_solvers.argtypes = [C.c_int, C.c_double]
        L.ref_inner_iterations.argtypes = [C.c_void_p, C.POINTER(C.c_long), C.POINTER(C.c_long)]
        L.ref_last_error.restype = C.c_char_p
        _lib = L
    return _lib


def set_inner_solvers(iterative: bool, rtol: float = 1e-5) -> None:
    """The KSPs inside the reference's PCABF: False = the model's exact solves (default, what the parity checks use), True =
    GMRES(30) + ILU(0) at rtol 1e-5, serial PETSc's defaults (ref_model/petsc_model_ksp.c).  Process-wide."""
    lib().ref_set_inner_solvers(1 if iterative else 0, float(rtol))


class Reference:
    """One NS object of type cnlinear (the reference's) on an n[0] x n[1] (x n[2]) Cartesian mesh with face coordinates xf.
    bcs: sequence of 2*dim objects with .type, .velocity(dim, t, x) and .pressure(dim, t, x), as oracle.BC."""

    def __init__(self, n: Sequence[int], xf: Sequence[np.ndarray], rho: float, mu: float, dt: float, bcs):
        self.dim = len(n)
        self.n = tuple(int(x) for x in n)
        self._xf = [np.ascontiguousarray(x, dtype=np.float64) for x in xf]
        self._bcs = list(bcs)
        self._error: Optional[BaseException] = None

        def cb(bnd, kind, dim, t, x, val):
            try:
                xs = [x[d] for d in range(dim)]
                if kind == 0:
                    out = self._bcs[bnd].velocity(dim, t, xs)
                    for d in range(dim):
                        val[d] = float(out[d])
                else:
                    val[0] = float(self._bcs[bnd].pressure(dim, t, xs))
                return 0
            except BaseException as exc:  # noqa: BLE001 - reported after the C call returns
                self._error = exc
                return 1

        self._cb = BC_CB(cb)
        nn = (C.c_int * 3)(*(list(self.n) + [1] * (3 - self.dim)))
        types = (C.c_int * 6)(*([int(b.type) for b in self._bcs] + [0] * (6 - 2 * self.dim)))
        ptr = [x.ctypes.data for x in self._xf] + [None] * (3 - self.dim)
        self._h = lib().ref_create(self.dim, nn, ptr[0], ptr[1], ptr[2], types, rho, mu, dt, self._cb)
        if not self._h:
            raise RuntimeError("ref_create failed: " + lib().ref_last_error().decode())
        nc, nf = C.c_long(), (C.c_long * 3)()
        lib().ref_sizes(self._h, C.byref(nc), nf)
        self.ncell, self.nface = nc.value, [nf[d] for d in range(self.dim)]
        self.nsol = self.ncell * (self.dim + 1) + sum(self.nface)
        per = [int(self._bcs[2 * d].type) == 3 for d in range(self.dim)]
        nz = self.n[2] if self.dim == 3 else 1
        self.cell_shape = (nz, self.n[1], self.n[0])
        self.face_shape = []
        for d in range(self.dim):
            s = [nz, self.n[1], self.n[0]]
            s[2 - d] += 0 if per[d] else 1
            self.face_shape.append(tuple(s))

    def _check(self, rc):
        if self._error is not None:
            e, self._error = self._error, None
            raise e
        if rc:
            raise RuntimeError(lib().ref_last_error().decode())

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                lib().ref_destroy(self._h)
                self._h = None
        except Exception:
            pass

    # ---- solution-sized vectors: v, U_x, U_y, (U_z), p back to back (the oracle's layout) ----
    def join(self, v, U, p):
        return np.concatenate([np.asarray(v, dtype=np.float64).ravel()] + [np.asarray(u, dtype=np.float64).ravel() for u in U] + [np.asarray(p, dtype=np.float64).ravel()])

    def split(self, x):
        nv = self.dim * self.ncell
        v = x[:nv].reshape((self.dim,) + self.cell_shape)
        U, off = [], nv
        for d in range(self.dim):
            U.append(x[off : off + self.nface[d]].reshape(self.face_shape[d]))
            off += self.nface[d]
        return v, U, x[off:].reshape(self.cell_shape)

    def set_state(self, v, U, p, phalf=None, step=0, t=0.0):
        x = np.ascontiguousarray(self.join(v, U, p))
        ph = np.ascontiguousarray(phalf, dtype=np.float64).ravel() if phalf is not None else None
        self._check(lib().ref_set_state(self._h, x.ctypes.data, ph.ctypes.data if ph is not None else None, int(step), float(t)))

    def get_state(self):
        x, ph = np.empty(self.nsol), np.empty(self.ncell)
        step, t = C.c_int(), C.c_double()
        self._check(lib().ref_get_state(self._h, x.ctypes.data, ph.ctypes.data, C.byref(step), C.byref(t)))
        v, U, p = self.split(x)
        return dict(v=v, U=U, p=p, phalf=ph.reshape(self.cell_shape), step=step.value, t=t.value)

    def step(self, mode=EXACT, schur_ainv=0, upper_ainv=0, rtol=1e-13, maxit=200):
        """NSStep; returns (iterations, residual history of the GMRES mode)."""
        its, nh = C.c_int(), C.c_int()
        hist = np.zeros(256)
        self._check(lib().ref_step(self._h, int(mode), int(schur_ainv), int(upper_ainv), float(rtol), int(maxit), C.byref(its), C.byref(nh), hist.ctypes.data))
        return its.value, hist[: nh.value].copy()

    def inner_iterations(self):
        """(momentum, Schur) Krylov iterations of the iterative inner KSPs since creation"""
        a, b = C.c_long(0), C.c_long(0)
        self._check(lib().ref_inner_iterations(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def last_rhs(self):
        b = np.empty(self.nsol)
        self._check(lib().ref_last_rhs(self._h, b.ctypes.data))
        return b

    def form_function(self):
        b = np.empty(self.nsol)
        self._check(lib().ref_form_function(self._h, b.ctypes.data))
        return b

    def abf_apply(self, b, schur_ainv=0, upper_ainv=0):
        b = np.ascontiguousarray(b, dtype=np.float64)
        x = np.empty(self.nsol)
        self._check(lib().ref_abf_apply(self._h, int(schur_ainv), int(upper_ainv), b.ctypes.data, x.ctypes.data))
        return x

    def apply_jacobian(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.empty(self.nsol)
        self._check(lib().ref_apply_jacobian(self._h, x.ctypes.data, y.ctypes.data))
        return y

    def matrix(self, name: str):
        """One of A G negT I negR D L Gst as scipy CSR in the oracle's numbering (explicit zeros kept)."""
        import scipy.sparse as sp

        nnz = lib().ref_matrix(self._h, name.encode(), None, None, None)
        if nnz < 0:
            raise KeyError(name)
        r, c, v = np.empty(nnz, dtype=np.int32), np.empty(nnz, dtype=np.int32), np.empty(nnz)
        got = lib().ref_matrix(self._h, name.encode(), r.ctypes.data, c.ctypes.data, v.ctypes.data)
        if got != nnz:
            raise RuntimeError(f"ref_matrix({name}) = {got}")
        nU = sum(self.nface)
        size = {"v": self.dim * self.ncell, "U": nU, "p": self.ncell}
        shape = {"A": ("v", "v"), "L": ("v", "v"), "G": ("v", "p"), "negT": ("U", "v"), "I": ("U", "U"), "negR": ("U", "p"), "Gst": ("U", "p"), "D": ("p", "U")}[name]
        return sp.csr_matrix((v, (r, c)), shape=(size[shape[0]], size[shape[1]]))
