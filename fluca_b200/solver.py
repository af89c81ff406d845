"""Thin numpy front-end of one rank's solver handle (include/fluca_b200.h).

This is plumbing above the C ABI: it allocates host arrays in the ABI's compact layout and
forwards.  All arithmetic happens in the CUDA library.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np

from . import _lib

BC_NONE, BC_VELOCITY, BC_PRESSURE_OUTLET, BC_PERIODIC, BC_SYMMETRY = range(5)
MODE_COUPLED, MODE_FRACTIONAL = 0, 1


def slab_partition(nz: int, nranks: int):
    """Equal-as-possible z-slabs (DMStag's default ownership, -cart_ranks_z P): [(k0, nzl)]."""
    base, rem = divmod(nz, nranks)
    out, k0 = [], 0
    for r in range(nranks):
        n = base + (1 if r < rem else 0)
        out.append((k0, n))
        k0 += n
    return out


class Comm:
    """Owner of a fluca_b200_comm handle (consumed by Solver)."""

    def __init__(self, L, handle, keep=()):
        self.L, self.handle, self._keep = L, handle, list(keep)

    @staticmethod
    def nccl(L, unique_id: bytes, rank: int, nranks: int) -> "Comm":
        h = C.c_void_p()
        buf = C.create_string_buffer(unique_id, len(unique_id))
        _lib.check(L, L.fluca_b200_comm_create_nccl(buf, len(unique_id), rank, nranks, C.byref(h)))
        return Comm(L, h)

    @staticmethod
    def unique_id(L) -> bytes:
        buf = C.create_string_buffer(256)
        n = C.c_int()
        _lib.check(L, L.fluca_b200_comm_unique_id(buf, 256, C.byref(n)))
        return buf.raw[: n.value]

    @staticmethod
    def callbacks(L, rank, nranks, halo, allsum, allgather=None) -> "Comm":
        hcb = _lib.HALO_FN(halo)
        acb = _lib.ALLSUM_FN(allsum)
        gcb = _lib.ALLGATHER_FN(allgather) if allgather else C.cast(None, _lib.ALLGATHER_FN)
        h = C.c_void_p()
        _lib.check(L, L.fluca_b200_comm_create_callbacks(rank, nranks, hcb, acb, gcb, None, C.byref(h)))
        return Comm(L, h, keep=(hcb, acb, gcb))


class Solver:
    def __init__(
        self,
        n: Sequence[int],
        xf: Sequence[np.ndarray],
        bc_types: Sequence[int],
        rho: float,
        mu: float,
        dt: float,
        *,
        mode: int = MODE_COUPLED,
        k0: int = 0,
        nzl: Optional[int] = None,
        comm: Optional[Comm] = None,
        library=None,
        **options,
    ):
        self.L = library if library is not None else _lib.load()
        self.dim = len(n)
        assert self.dim in (2, 3)
        self.n = tuple(int(a) for a in n) + (1,) * (3 - self.dim)
        self.bc_types = [int(b) for b in bc_types] + [BC_NONE] * (6 - len(bc_types))
        self.periodic = [self.bc_types[2 * d] == BC_PERIODIC for d in range(3)]
        self._xf = [np.ascontiguousarray(a, dtype=np.float64) for a in xf]
        self.k0 = int(k0) if self.dim == 3 else 0
        self.nzl = int(nzl if nzl is not None else self.n[2]) if self.dim == 3 else 1
        self.nzg = self.n[2]
        self.last_z = self.dim == 3 and (self.k0 + self.nzl == self.nzg) and not self.periodic[2]
        d = _lib.Desc()
        d.dim = self.dim
        for a in range(3):
            d.n[a] = self.n[a]
        for a in range(self.dim):
            assert self._xf[a].shape == (self.n[a] + 1,)
            d.xf[a] = self._xf[a].ctypes.data_as(C.POINTER(C.c_double))
        for b in range(6):
            d.bc_type[b] = self.bc_types[b]
        d.rho, d.mu, d.dt = rho, mu, dt
        d.k0, d.nzl = self.k0, self.nzl
        d.mode = mode
        for k, v in options.items():
            if not hasattr(d, k):
                raise TypeError(f"unknown solver option {k!r}")
            setattr(d, k, v)
        self._comm = comm
        h = C.c_void_p()
        _lib.check(self.L, self.L.fluca_b200_create(C.byref(d), comm.handle if comm else None, C.byref(h)))
        self._h = h
        self.rho, self.mu, self.dt, self.mode = rho, mu, dt, mode
        self.ainv_types = (0, 0)
        nx, ny = self.n[0], self.n[1]
        self.cell_shape = (self.nzl, ny, nx)
        self.face_shape = []
        for a in range(self.dim):
            s = [self.nzl, ny, nx]
            if a == 0 and not self.periodic[0]:
                s[2] += 1
            if a == 1 and not self.periodic[1]:
                s[1] += 1
            if a == 2 and self.last_z:
                s[0] += 1
            self.face_shape.append(tuple(s))
        self.bc_shape = [(self.nzl, ny), (self.nzl, ny), (self.nzl, nx), (self.nzl, nx), (ny, nx), (ny, nx)]

    # ------------------------------------------------------------------ helpers
    def close(self):
        if getattr(self, "_h", None):
            self.L.fluca_b200_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ptrs(self, arrs):
        p = (C.c_void_p * 3)()
        for a in range(self.dim):
            p[a] = arrs[a].ctypes.data if arrs[a] is not None else None
        return p

    def _cells(self, x, ncomp=None):
        shp = self.cell_shape if ncomp is None else (ncomp,) + self.cell_shape
        return np.ascontiguousarray(x, dtype=np.float64).reshape(shp)

    def _faces(self, U):
        return [np.ascontiguousarray(U[a], dtype=np.float64).reshape(self.face_shape[a]) for a in range(self.dim)]

    def new_cells(self, ncomp=None):
        return np.zeros(self.cell_shape if ncomp is None else (ncomp,) + self.cell_shape)

    def new_faces(self):
        return [np.zeros(s) for s in self.face_shape]

    # ------------------------------------------------------------------ state
    def set_state(self, v=None, U=None, p=None, phalf=None):
        v = self._cells(v, self.dim) if v is not None else None
        Us = self._faces(U) if U is not None else None
        p = self._cells(p) if p is not None else None
        ph = self._cells(phalf) if phalf is not None else None
        _lib.check(
            self.L,
            self.L.fluca_b200_set_state(
                self._h,
                v.ctypes.data if v is not None else None,
                self._ptrs(Us) if Us is not None else None,
                p.ctypes.data if p is not None else None,
                ph.ctypes.data if ph is not None else None,
            ),
        )

    def get_state(self):
        v, U, p, ph = self.new_cells(self.dim), self.new_faces(), self.new_cells(), self.new_cells()
        _lib.check(self.L, self.L.fluca_b200_get_state(self._h, v.ctypes.data, self._ptrs(U), p.ctypes.data, ph.ctypes.data))
        return dict(v=v, U=U, p=p, phalf=ph)

    def stage_state(self):
        """Enqueue an asynchronous device -> pinned-host copy of the current state (returns at once)."""
        _lib.check(self.L, self.L.fluca_b200_stage_state(self._h))

    def staged_state(self, copy: bool = True):
        """Wait for the staged copy; dict like get_state().  copy=False returns views of the library's pinned buffers
        (valid until the next stage_state)."""
        v, p, ph = C.c_void_p(), C.c_void_p(), C.c_void_p()
        U = (C.c_void_p * 3)()
        _lib.check(self.L, self.L.fluca_b200_staged_state(self._h, C.byref(v), U, C.byref(p), C.byref(ph)))

        def arr(ptr, shape):
            a = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_double)), shape=shape)
            return a.copy() if copy else a

        return dict(v=arr(v, (self.dim,) + self.cell_shape), U=[arr(U[d], self.face_shape[d]) for d in range(self.dim)], p=arr(p, self.cell_shape), phalf=arr(ph, self.cell_shape))

    def set_boundary_velocity(self, b, slot, values):
        a = np.ascontiguousarray(values, dtype=np.float64).reshape((self.dim,) + self.bc_shape[b])
        _lib.check(self.L, self.L.fluca_b200_set_boundary_velocity(self._h, b, slot, a.ctypes.data))

    def set_boundary_pressure(self, b, slot, values):
        a = np.ascontiguousarray(values, dtype=np.float64).reshape(self.bc_shape[b])
        _lib.check(self.L, self.L.fluca_b200_set_boundary_pressure(self._h, b, slot, a.ctypes.data))

    # ------------------------------------------------------------------ stepping
    def step(self, t: float, step_index: int) -> _lib.Stats:
        st = _lib.Stats()
        _lib.check(self.L, self.L.fluca_b200_step(self._h, float(t), int(step_index), C.byref(st)))
        return st

    def prepare_step(self, t: float, step_index: int):
        _lib.check(self.L, self.L.fluca_b200_prepare_step(self._h, float(t), int(step_index)))

    def get_rhs(self):
        rm, ri, rc = self.new_cells(self.dim), self.new_faces(), self.new_cells()
        _lib.check(self.L, self.L.fluca_b200_get_rhs(self._h, rm.ctypes.data, self._ptrs(ri), rc.ctypes.data))
        return rm, ri, rc

    def apply_momentum(self, x):
        x = self._cells(x, self.dim)
        y = self.new_cells(self.dim)
        _lib.check(self.L, self.L.fluca_b200_apply_momentum(self._h, x.ctypes.data, y.ctypes.data))
        return y

    def apply_schur(self, p):
        p = self._cells(p)
        y = self.new_cells()
        _lib.check(self.L, self.L.fluca_b200_apply_schur(self._h, p.ctypes.data, y.ctypes.data))
        return y

    def apply_vcycle(self, r):
        r = self._cells(r)
        z = self.new_cells()
        _lib.check(self.L, self.L.fluca_b200_apply_vcycle(self._h, r.ctypes.data, z.ctypes.data))
        return z

    def apply_coupled(self, xv, xU, xp):
        xv, xU, xp = self._cells(xv, self.dim), self._faces(xU), self._cells(xp)
        yv, yU, yp = self.new_cells(self.dim), self.new_faces(), self.new_cells()
        _lib.check(self.L, self.L.fluca_b200_apply_coupled(self._h, xv.ctypes.data, self._ptrs(xU), xp.ctypes.data, yv.ctypes.data, self._ptrs(yU), yp.ctypes.data))
        return yv, yU, yp

    def apply_abf(self, bv, bU, bp):
        bv, bU, bp = self._cells(bv, self.dim), self._faces(bU), self._cells(bp)
        xv, xU, xp = self.new_cells(self.dim), self.new_faces(), self.new_cells()
        st = _lib.Stats()
        _lib.check(self.L, self.L.fluca_b200_apply_abf(self._h, bv.ctypes.data, self._ptrs(bU), bp.ctypes.data, xv.ctypes.data, self._ptrs(xU), xp.ctypes.data, C.byref(st)))
        return xv, xU, xp, st

    def set_abf_ainv_types(self, schur_type: int = 0, upper_type: int = 0):
        """PCABFSetSchurComplementAinvType / PCABFSetUpperTriangularAinvType (flucans.h:106-107): 0 ID, 1 DIAG, 2 ROWSUM."""
        _lib.check(self.L, self.L.fluca_b200_set_abf_ainv_types(self._h, int(schur_type), int(upper_type)))
        self.ainv_types = (int(schur_type), int(upper_type))

    def set_inner_monitor(self, fn=None):
        """fn(which, it, rnorm) per inner residual norm (which: 0 momentum, 1 Schur; it = 0: initial residual of a solve) -- the
        analogue of -ns_abf_momentum_ksp_monitor / -ns_abf_schur_ksp_monitor; None switches it off."""
        self._monitor_cb = _lib.INNER_MONITOR_FN(lambda ctx, which, it, rnorm: fn(which, it, rnorm)) if fn else _lib.INNER_MONITOR_FN()
        _lib.check(self.L, self.L.fluca_b200_set_inner_monitor(self._h, self._monitor_cb, None))

    # ------------------------------------------------------------------ immersed boundary
    def set_markers(self, X, Ud, dV, delta_points: int = 4, iterations: int = 1):
        """X, Ud: (dim, n) marker positions / prescribed velocities; dV: (n,) volume weights (replicated on every rank)."""
        X = np.ascontiguousarray(X, dtype=np.float64).reshape(self.dim, -1)
        self.nm = X.shape[1]
        Ud = np.ascontiguousarray(Ud, dtype=np.float64).reshape(self.dim, self.nm)
        dV = np.ascontiguousarray(dV, dtype=np.float64).reshape(self.nm)
        _lib.check(self.L, self.L.fluca_b200_set_markers(self._h, self.nm, X.ctypes.data, Ud.ctypes.data, dV.ctypes.data, int(delta_points)))
        _lib.check(self.L, self.L.fluca_b200_set_ibm_iterations(self._h, int(iterations)))

    def marker_forces(self):
        F, Um = np.zeros((self.dim, self.nm)), np.zeros((self.dim, self.nm))
        _lib.check(self.L, self.L.fluca_b200_get_marker_forces(self._h, F.ctypes.data, Um.ctypes.data))
        return F, Um

    def ibm_interpolate(self, v):
        v = self._cells(v, self.dim)
        Um = np.zeros((self.dim, self.nm))
        _lib.check(self.L, self.L.fluca_b200_ibm_interpolate(self._h, v.ctypes.data, Um.ctypes.data))
        return Um

    def ibm_spread(self, Fm):
        Fm = np.ascontiguousarray(Fm, dtype=np.float64).reshape(self.dim, self.nm)
        f = self.new_cells(self.dim)
        _lib.check(self.L, self.L.fluca_b200_ibm_spread(self._h, Fm.ctypes.data, f.ctypes.data))
        return f

    # ------------------------------------------------------------------ device-resident helpers
    def snapshot_save(self):
        _lib.check(self.L, self.L.fluca_b200_snapshot_save(self._h))

    def snapshot_restore(self):
        _lib.check(self.L, self.L.fluca_b200_snapshot_restore(self._h))

    def stream(self) -> int:
        p = C.c_void_p()
        _lib.check(self.L, self.L.fluca_b200_stream(self._h, C.byref(p)))
        return p.value or 0

    def launch_count(self) -> int:
        return int(self.L.fluca_b200_launch_count(self._h))

    def model_bytes(self, stats) -> float:
        return float(self.L.fluca_b200_step_model_bytes(self._h, C.byref(stats)))

    def kernel_timing(self, enable: bool):
        _lib.check(self.L, self.L.fluca_b200_kernel_timing(self._h, 1 if enable else 0))

    def ibm_info(self):
        """(markers this rank works on, shared with the lower slab, shared with the upper slab, neighbour exchange in use)"""
        info = (C.c_long * 4)()
        _lib.check(self.L, self.L.fluca_b200_ibm_info(self._h, info))
        return int(info[0]), int(info[1]), int(info[2]), bool(info[3])

    def model_bytes_split(self, stats):
        """SURVEY.md 8d model bytes of one step with these iteration counts, per kernel class name."""
        n = len(_lib.KT_NAMES)
        by = (C.c_double * n)()
        _lib.check(self.L, self.L.fluca_b200_step_model_bytes_split(self._h, C.byref(stats), by))
        return {_lib.KT_NAMES[i]: by[i] for i in range(n)}

    def kernel_times(self, reset: bool = True):
        n = len(_lib.KT_NAMES)
        ms, cnt = (C.c_double * n)(), (C.c_long * n)()
        _lib.check(self.L, self.L.fluca_b200_kernel_times(self._h, ms, cnt, 1 if reset else 0))
        return {_lib.KT_NAMES[i]: (ms[i], cnt[i]) for i in range(n)}

    def time_kernel(self, name: str, reps: int = 10):
        ms, by = C.c_double(), C.c_double()
        _lib.check(self.L, self.L.fluca_b200_time_kernel(self._h, name.encode(), reps, C.byref(ms), C.byref(by)))
        return ms.value, by.value
