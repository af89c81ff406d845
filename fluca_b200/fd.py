"""Host-side mirror of the reference's FlucaFD interface (fluca/include/flucafd.h) over the C ABI's stencil layer
(include/fluca_b200.h "FlucaFD", csrc/fd.cu): same constructors, argument order and error behaviour, so that the tests read
like fluca/tests/fd/ex*.c.  One rank; the device apply generated from these stencils is the next step."""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _lib

# DMStagStencilLocation values used by FlucaFD (petscdmstag.h) and their option names (DMStagStencilLocations[])
DMSTAG_BACK_DOWN_LEFT, DMSTAG_BACK_DOWN, DMSTAG_BACK_LEFT, DMSTAG_BACK, DMSTAG_DOWN_LEFT, DMSTAG_DOWN, DMSTAG_LEFT, DMSTAG_ELEMENT = 1, 2, 4, 5, 10, 11, 13, 14
LOCATION_NAMES = {1: "BACK_DOWN_LEFT", 2: "BACK_DOWN", 4: "BACK_LEFT", 5: "BACK", 10: "DOWN_LEFT", 11: "DOWN", 13: "LEFT", 14: "ELEMENT"}
LOCATION_BY_NAME = {v.lower(): k for k, v in LOCATION_NAMES.items()}
FLUCAFD_X, FLUCAFD_Y, FLUCAFD_Z = 0, 1, 2
FLUCAFD_BC_NONE, FLUCAFD_BC_DIRICHLET, FLUCAFD_BC_NEUMANN = 0, 1, 2
BC_BY_NAME = {"none": 0, "dirichlet": 1, "neumann": 2}
FLUCAFD_CONSTANT = -7
BOUNDARY_NAMES = ["left", "right", "down", "up", "back", "front"]


class FlucaFDError(RuntimeError):
    pass


def _check(L, rc):
    if rc != 0:
        raise FlucaFDError((L.fluca_b200_fd_last_error() or b"").decode())


class FDGrid:
    """The DMStag a FlucaFD is set on: DMStagCreate{1,2,3}d + DMStagSetUniformCoordinatesProduct (or explicit coordinates)."""

    def __init__(self, n: Sequence[int], faces: Sequence[np.ndarray], periodic: Optional[Sequence[bool]] = None, stencil_width: int = 1, centres=None, library=None):
        self.L = library if library is not None else _lib.load()
        self.dim = len(n)
        self.n = tuple(int(a) for a in n)
        self.periodic = tuple(bool(p) for p in (periodic or [False] * self.dim))
        self._xf = [np.ascontiguousarray(f, dtype=np.float64) for f in faces]
        self._xc = [np.ascontiguousarray(c, dtype=np.float64) for c in centres] if centres is not None else None
        nn = (C.c_int * 3)(*(list(self.n) + [1] * (3 - self.dim)))
        per = (C.c_int * 3)(*([int(p) for p in self.periodic] + [0] * (3 - self.dim)))
        xf, xc = (C.c_void_p * 3)(), (C.c_void_p * 3)()
        for d in range(self.dim):
            assert self._xf[d].shape == (self.n[d] + 1,)
            xf[d] = self._xf[d].ctypes.data
            if self._xc is not None:
                xc[d] = self._xc[d].ctypes.data
        h = C.c_void_p()
        _check(self.L, self.L.fluca_b200_fd_grid_create(self.dim, nn, xf, xc if self._xc is not None else None, per, int(stencil_width), C.byref(h)))
        self._h = h

    @staticmethod
    def uniform(n, lo, hi, periodic=None, stencil_width=1, library=None) -> "FDGrid":
        return FDGrid(n, [np.linspace(a, b, m + 1) for m, a, b in zip(n, lo, hi)], periodic, stencil_width, library=library)

    def field_shape(self, loc: int):
        ext = []
        for d in range(self.dim):
            face = LOCATION_NAMES[loc].lower().split("_").count(("left", "down", "back")[d]) > 0
            ext.append(self.n[d] + (1 if face and not self.periodic[d] else 0))
        return tuple(reversed(ext))

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                self.L.fluca_b200_fd_grid_destroy(self._h)
                self._h = None
        except Exception:
            pass


class FlucaFD:
    def __init__(self, grid: FDGrid, handle, keep=()):
        self.grid, self.L, self._h, self._keep = grid, grid.L, handle, list(keep)

    # ---- flucafd.h:59-66
    def SetLocations(self, input_loc, input_c, output_loc, output_c):
        _check(self.L, self.L.fluca_b200_fd_set_locations(self._h, input_loc, input_c, output_loc, output_c))

    def SetBoundaryCondition(self, boundary: int, bc_type: int, value: float = 0.0):
        _check(self.L, self.L.fluca_b200_fd_set_boundary_condition(self._h, boundary, bc_type, float(value)))

    def SetUp(self):
        _check(self.L, self.L.fluca_b200_fd_setup(self._h))
        return self

    def GetStencil(self, i: int, j: int = 0, k: int = 0) -> List[Tuple[Tuple[int, int, int, int, int], float]]:
        """FlucaFDGetStencil (flucafd.h:74): [((i, j, k, loc, c), v), ...]"""
        n = C.c_int()
        col = (_lib.FDCol * _lib.FD_MAX_STENCIL)()
        v = (C.c_double * _lib.FD_MAX_STENCIL)()
        _check(self.L, self.L.fluca_b200_fd_get_stencil(self._h, i, j, k, C.byref(n), col, v))
        return [((col[q].i, col[q].j, col[q].k, col[q].loc, col[q].c), v[q]) for q in range(n.value)]

    def ApplyInputs(self) -> List[Tuple[int, int]]:
        """(location, component) of every input field the composed operator reads, in the order Apply expects them"""
        n, loc, c = C.c_int(), (C.c_int * 4)(), (C.c_int * 4)()
        _check(self.L, self.L.fluca_b200_fd_apply_inputs(self._h, C.byref(n), loc, c))
        return [(loc[q], c[q]) for q in range(n.value)]

    def Apply(self, inputs, output_loc: int) -> np.ndarray:
        """FlucaFDApply (flucafd.h:75) on the device.  inputs: {(loc, c): field} or a list in ApplyInputs() order."""
        order = self.ApplyInputs()
        if isinstance(inputs, dict):
            inputs = [inputs[k] for k in order]
        arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in inputs]
        for a, (loc, _) in zip(arrs, order):
            assert a.shape == self.grid.field_shape(loc), (a.shape, self.grid.field_shape(loc))
        ptr = (C.c_void_p * max(len(arrs), 1))(*[a.ctypes.data for a in arrs])
        out = np.zeros(self.grid.field_shape(output_loc))
        _check(self.L, self.L.fluca_b200_fd_apply(self._h, len(arrs), ptr, out.ctypes.data))
        return out

    def GetOperator(self):
        """FlucaFDGetOperator (fdapply.c:123-180) as scipy-style CSR pieces over the output points: (rowptr, cols, vals) with
        cols = [(i, j, k, loc, c)] of the interior stencil points (boundary and constant terms are not part of the matrix)."""
        nr, nz = C.c_long(), C.c_long()
        _check(self.L, self.L.fluca_b200_fd_get_operator(self._h, C.byref(nr), C.byref(nz), None, None, None))
        rowptr = (C.c_long * (nr.value + 1))()
        cols = (_lib.FDCol * max(nz.value, 1))()
        vals = (C.c_double * max(nz.value, 1))()
        _check(self.L, self.L.fluca_b200_fd_get_operator(self._h, C.byref(nr), C.byref(nz), rowptr, cols, vals))
        return np.array(rowptr[: nr.value + 1]), [(cols[q].i, cols[q].j, cols[q].k, cols[q].loc, cols[q].c) for q in range(nz.value)], np.array(vals[: nz.value])

    def ApplyDevice(self, dev_inputs: Sequence[int], dev_output: int):
        """device pointers (ints) in ApplyInputs() order; asynchronous on Stream(); Sync() waits"""
        ptr = (C.c_void_p * max(len(dev_inputs), 1))(*dev_inputs)
        _check(self.L, self.L.fluca_b200_fd_apply_device(self._h, len(dev_inputs), ptr, dev_output))

    def Stream(self) -> int:
        st = C.c_void_p()
        _check(self.L, self.L.fluca_b200_fd_stream(self._h, C.byref(st)))
        return st.value or 0

    def Sync(self):
        _check(self.L, self.L.fluca_b200_fd_sync(self._h))

    def Destroy(self):
        if getattr(self, "_h", None):
            self.L.fluca_b200_fd_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.Destroy()
        except Exception:
            pass


def _new(grid, fn, *args, keep=()):
    h = C.c_void_p()
    _check(grid.L, fn(*args, C.byref(h)))
    return FlucaFD(grid, h, keep)


def FlucaFDDerivativeCreate(grid: FDGrid, direction, deriv_order, accu_order, input_loc, input_c, output_loc, output_c) -> FlucaFD:  # flucafd.h:82
    return _new(grid, grid.L.fluca_b200_fd_derivative_create, grid._h, direction, deriv_order, accu_order, input_loc, input_c, output_loc, output_c, keep=[grid])


def FlucaFDSumCreate(ops: Sequence[FlucaFD]) -> FlucaFD:  # flucafd.h:97
    arr = (C.c_void_p * len(ops))(*[o._h for o in ops])
    return _new(ops[0].grid, ops[0].L.fluca_b200_fd_sum_create, len(ops), arr, keep=list(ops))


def FlucaFDScaleCreateConstant(operand: FlucaFD, constant: float) -> FlucaFD:  # flucafd.h:90
    return _new(operand.grid, operand.L.fluca_b200_fd_scale_create_constant, operand._h, float(constant), keep=[operand])


def FlucaFDScaleCreateVector(operand: FlucaFD, field: np.ndarray, vec_loc: int, vec_c: int = 0) -> FlucaFD:  # flucafd.h:91
    f = np.ascontiguousarray(field, dtype=np.float64)
    assert f.shape == operand.grid.field_shape(vec_loc), (f.shape, operand.grid.field_shape(vec_loc))
    return _new(operand.grid, operand.L.fluca_b200_fd_scale_create_vector, operand._h, f.ctypes.data, vec_loc, vec_c, keep=[operand])


def FlucaFDCompositionCreate(inner: FlucaFD, outer: FlucaFD) -> FlucaFD:  # flucafd.h:86
    return _new(inner.grid, inner.L.fluca_b200_fd_composition_create, inner._h, outer._h, keep=[inner, outer])


def FlucaFDSecondOrderTVDCreate(grid: FDGrid, direction, input_c=0, output_c=0) -> FlucaFD:  # flucafd.h:105
    return _new(grid, grid.L.fluca_b200_fd_tvd_create, grid._h, direction, input_c, output_c, keep=[grid])


def FlucaFDSecondOrderTVDSetLimiter(fd: FlucaFD, name: str):
    _check(fd.L, fd.L.fluca_b200_fd_tvd_set_limiter(fd._h, name.encode()))


def FlucaFDSecondOrderTVDSetVelocity(fd: FlucaFD, face_velocity: np.ndarray):
    f = np.ascontiguousarray(face_velocity, dtype=np.float64)
    _check(fd.L, fd.L.fluca_b200_fd_tvd_set_velocity(fd._h, f.ctypes.data))


def FlucaFDSecondOrderTVDSetCurrentSolution(fd: FlucaFD, phi: np.ndarray):
    f = np.ascontiguousarray(phi, dtype=np.float64)
    assert f.shape == fd.grid.field_shape(DMSTAG_ELEMENT)
    _check(fd.L, fd.L.fluca_b200_fd_tvd_set_current_solution(fd._h, f.ctypes.data))
