"""oracle/fd_oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement (pure Python, small cases only) of the stencil layer of the reference's FlucaFD operator family
(SURVEY.md 8f rank 4), for one rank on a DMStag product grid:

    fluca/src/fd/interface/fdbasic.c:149-192      FlucaFDSetUp (grid facts taken from the DMStag)
    fluca/src/fd/interface/fdapply.c:22-121       FlucaFDGetStencilRaw / FlucaFDGetStencil / FlucaFDApply
    fluca/src/fd/utils/fdutils.c:56-100           coordinates outside the local grid, ghost corners, Gaussian elimination
    fluca/src/fd/utils/fdutils.c:102-489          stencil accumulation, off-grid point removal by boundary condition, zero removal
    fluca/src/fd/impls/derivative/derivative.c    FLUCAFDDERIVATIVE (Vandermonde weights on the product coordinates)
    fluca/src/fd/impls/sum/sum.c                  FLUCAFDSUM
    fluca/src/fd/impls/scale/scale.c              FLUCAFDSCALE
    fluca/src/fd/impls/composition/composition.c  FLUCAFDCOMPOSITION

PARITY STATUS: *pinned*.  The reference's own tests for this family print stencils with "%g" and compare them byte for byte
with stored outputs (fluca/tests/fd/ex*.c, fluca/tests/fd/output/*.out); tests/golden/fd_stencils.json holds those
arguments and outputs (extracted by tests/golden/make_fd_stencils.py) and tests/test_oracle_fd.py reproduces the printed lines.

What PETSc (un-vendored) contributes here is only the DMStag index conventions, restated from its documentation: an
element index i owns the point at its LEFT/DOWN/BACK face and its centre; a non-periodic direction has one extra face
index N; a periodic direction has `stencil_width` ghost elements on either side whose coordinates continue the grid.

Only tests/ may import this module.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

MAX_STENCIL = 32          # FLUCAFD_MAX_STENCIL_SIZE, flucafdimpl.h:8
ZERO_PIVOT_TOL = 1e-14    # flucafdimpl.h:9
COEFF_ATOL = 1e-10        # flucafdimpl.h:10
COEFF_RTOL = 1e-8         # flucafdimpl.h:11
INT_MAX = 2**31 - 1

BC_NONE, BC_DIRICHLET, BC_NEUMANN = "none", "dirichlet", "neumann"  # FlucaFDBoundaryConditionType, flucafd.h:31-36
CONSTANT = -7             # FLUCAFD_CONSTANT, flucafd.h:52
BOUNDARY_NAMES = ["left", "right", "down", "up", "back", "front"]  # fdtest.h:7

# DMStagStencilLocation: the numeric order decides the sort order of the printed stencils (fdtest.h:9-36)
LOCATIONS = ["null", "back_down_left", "back_down", "back_down_right", "back_left", "back", "back_right", "back_up_left", "back_up", "back_up_right", "down_left", "down", "down_right", "left", "element", "right", "up_left", "up", "up_right", "front_down_left", "front_down", "front_down_right", "front_left", "front", "front_right", "front_up_left", "front_up", "front_up_right"]
LOC_ID = {n: i for i, n in enumerate(LOCATIONS)}
ALLOWED = {"element", "left", "down", "back", "down_left", "back_left", "back_down", "back_down_left"}  # fdutils.c:16-34


def use_face(loc: str, d: int) -> bool:
    """FlucaFDUseFaceCoordinate_Internal, fdutils.c:36-54"""
    return ("left", "down", "back")[d] in loc.split("_")


def boundary_location(loc: str, d: int) -> str:
    """GetBoundaryStencilLocation_Private, fdutils.c:198-252: add the face of direction d to the location"""
    parts = set() if loc == "element" else set(loc.split("_"))
    parts.add(("left", "down", "back")[d])
    return "_".join(p for p in ("back", "down", "left") if p in parts)


def solve(n: int, A: List[List[float]], b: List[float]) -> List[float]:
    """FlucaFDSolveLinearSystem_Internal, fdutils.c:79-100: Gaussian elimination WITHOUT pivoting"""
    A = [row[:] for row in A]
    b = b[:]
    for k in range(n - 1):
        if not abs(A[k][k]) > ZERO_PIVOT_TOL:
            raise ZeroDivisionError("zero pivot in forward elimination")
        for i in range(k + 1, n):
            f = A[i][k] / A[k][k]
            for j in range(k, n):
                A[i][j] -= f * A[k][j]
            b[i] -= f * b[k]
    x = [0.0] * n
    for i in range(n - 1, -1, -1):
        if not abs(A[i][i]) > ZERO_PIVOT_TOL:
            raise ZeroDivisionError("zero pivot in back substitution")
        s = 0.0
        for j in range(i + 1, n):
            s += A[i][j] * x[j]
        x[i] = (b[i] - s) / A[i][i]
    return x


@dataclass(frozen=True)
class Col:
    """DMStagStencil: element index, location, component (negative: boundary value / constant marker)"""
    i: int
    j: int
    k: int
    loc: str
    c: int

    def idx(self, d):
        return (self.i, self.j, self.k)[d]

    def moved(self, d, value):
        ijk = [self.i, self.j, self.k]
        ijk[d] = value
        return Col(ijk[0], ijk[1], ijk[2], self.loc, self.c)


@dataclass
class Term:
    """FlucaFDTermLink, flucafdimpl.h: what one additive term of an operator differentiates, per direction"""
    deriv_order: List[int] = field(default_factory=lambda: [-1, -1, -1])
    accu_order: List[int] = field(default_factory=lambda: [INT_MAX, INT_MAX, INT_MAX])
    input_loc: str = "element"
    input_c: int = 0

    def key(self):
        return (tuple(self.deriv_order), tuple(self.accu_order), self.input_loc, self.input_c)

    def copy(self):
        return Term(self.deriv_order[:], self.accu_order[:], self.input_loc, self.input_c)


class Grid:
    """One rank's DMStag with uniform product coordinates (DMStagSetUniformCoordinatesProduct)."""

    def __init__(self, N: Sequence[int], lo: Sequence[float], hi: Sequence[float], periodic: Sequence[bool] = (False, False, False), stencil_width: int = 1):
        self.dim = len(N)
        self.N = list(N) + [1] * (3 - self.dim)
        self.lo, self.hi = list(lo), list(hi)
        self.periodic = list(periodic) + [False] * (3 - len(periodic))
        self.sw = stencil_width

    def ghost_corners(self, d: int, face: bool) -> Tuple[int, int, int]:
        """FlucaFDGetGhostCorners_Internal, fdutils.c:67-77, for first rank = last rank"""
        per = self.periodic[d]
        gxs = 0 if not per else -self.sw
        gxm = self.N[d] + (0 if not per else 2 * self.sw)
        gxe = 1 if (face and not per) else 0
        return gxs, gxm, gxe

    def array_coord(self, d: int, idx: int, face: bool) -> float:
        """the DMStag product coordinate array entry: faces at lo + idx h, centres at lo + (idx + 1/2) h, ghosts continue"""
        h = (self.hi[d] - self.lo[d]) / self.N[d]
        return self.lo[d] + (idx + (0.0 if face else 0.5)) * h

    def coordinate(self, d: int, idx: int, face: bool, x: int, n: int, h_prev: float, h_next: float) -> float:
        """FlucaFDGetCoordinate_Internal, fdutils.c:56-65: array inside [x, x + n), uniform continuation outside"""
        if x <= idx < x + n:
            return self.array_coord(d, idx, face)
        if idx < x:
            return self.array_coord(d, x, face) - (x - idx) * h_prev
        return self.array_coord(d, x + n - 1, face) + (idx - (x + n - 1)) * h_next

    def end_spacings(self, d: int, gxs: int, gxm: int) -> Tuple[float, float]:
        """h_prev / h_next of derivative.c:79-82 and fdutils.c:300-303"""
        first = gxs
        last = gxs + gxm - (0 if not self.periodic[d] else 1)
        return (self.array_coord(d, first + 1, True) - self.array_coord(d, first, True), self.array_coord(d, last, True) - self.array_coord(d, last - 1, True))


class FD:
    """struct _p_FlucaFD: input / output location and component, six boundary conditions, the term list"""

    def __init__(self, grid: Grid, input_loc="element", input_c=0, output_loc="element", output_c=0):
        for loc in (input_loc, output_loc):
            if loc not in ALLOWED:
                raise ValueError(f"Invalid stencil location {loc}")
        self.grid, self.dim = grid, grid.dim
        self.input_loc, self.input_c, self.output_loc, self.output_c = input_loc, input_c, output_loc, output_c
        self.bcs = [(BC_NONE, 0.0)] * 6  # LEFT, RIGHT, DOWN, UP, BACK, FRONT
        self.terms: List[Term] = []

    def set_bc(self, boundary: int, kind: str, value: float = 0.0):
        self.bcs[boundary] = (kind, value)

    # ---- fdapply.c:22-45
    def stencil_raw(self, i, j, k):
        return remove_zero(self._raw(i, j, k))

    def stencil(self, i, j, k):
        return remove_zero(remove_off_grid(self, self.stencil_raw(i, j, k)))

    def apply_point(self, i, j, k, value_at):
        """one output point of FlucaFDApply, fdapply.c:85-106; value_at(Col) reads the input field"""
        r = 0.0
        for col, v in self.stencil(i, j, k):
            if col.c >= 0:
                r += v * value_at(col)
            elif col.c == CONSTANT:
                r += v
            else:
                r += v * self.bcs[-col.c - 1][1]
        return r


def add_point(st: List[Tuple[Col, float]], col: Col, v: float):
    """FlucaFDAddStencilPoint_Internal, fdutils.c:102-124"""
    for n, (c, w) in enumerate(st):
        if c == col:
            st[n] = (c, w + v)
            return
    if len(st) >= MAX_STENCIL:
        raise ValueError("Resulting stencil is too large")
    st.append((col, v))


def remove_zero(st):
    """FlucaFDRemoveZeroStencilPoints_Internal, fdutils.c:465-489"""
    s = sum(abs(v) for _, v in st)
    return [(c, v) for c, v in st if not (abs(v) < COEFF_ATOL or abs(v / s) < COEFF_RTOL)]


def _off_grid(fd: FD, col: Col):
    """IsOffGrid_Private, fdutils.c:126-168: (direction, is_low) of the first direction in which the point is outside"""
    if col.c < 0:
        return None
    for d in range(fd.dim):
        gxs, gxm, gxe = fd.grid.ghost_corners(d, use_face(col.loc, d))
        if col.idx(d) < gxs:
            return d, True
        if col.idx(d) >= gxs + gxm + gxe:
            return d, False
    return None


def _off_grid_stencil_size(fd: FD, col: Col, d: int) -> int:
    """GetStencilSizeForOffGridPoint_Private, fdutils.c:170-196"""
    m = INT_MAX
    for t in fd.terms:
        if t.deriv_order[d] != -1 and t.accu_order[d] != INT_MAX and t.input_loc == col.loc and t.input_c == col.c:
            m = min(m, t.deriv_order[d] + t.accu_order[d])
    if m == INT_MAX:
        raise ValueError("Cannot find a term from the given stencil point")
    return max(m, 1)


def remove_off_grid(fd: FD, st):
    """FlucaFDRemoveOffGridPoints_Internal, fdutils.c:254-463: one off-grid point at a time is replaced by its
    extrapolation from on-grid points and, with a Dirichlet / Neumann condition, the boundary value marker"""
    st = list(st)
    g = fd.grid
    for _ in range(100):
        found = None
        for n, (col, _) in enumerate(st):
            og = _off_grid(fd, col)
            if og is not None:
                found = (n, og)
                break
        if found is None:
            break
        n, (d, low) = found
        off_col, off_v = st.pop(n)
        per = g.periodic[d]
        bc = fd.bcs[2 * d][0] if (low and not per) else (fd.bcs[2 * d + 1][0] if (not low and not per) else BC_NONE)
        face = use_face(off_col.loc, d)
        size = _off_grid_stencil_size(fd, off_col, d)
        gxs, gxm, gxe = g.ghost_corners(d, face)
        hp, hn = g.end_spacings(d, gxs, gxm)
        coord = lambda idx, fc=face: g.coordinate(d, idx, fc, gxs, gxm + gxe, hp, hn)  # noqa: E731
        off_x = coord(off_col.idx(d))
        if bc == BC_NONE:  # :326-356
            start = gxs if low else gxs + gxm + gxe - size
            xs = [coord(start + m) for m in range(size)]
            w = solve(size, [[(xs[m] - off_x) ** r for m in range(size)] for r in range(size)], [1.0 if r == 0 else 0.0 for r in range(size)])
            for m in range(size):
                add_point(st, off_col.moved(d, start + m), off_v * w[m])
        elif bc == BC_DIRICHLET:  # :358-404
            start = gxs if low else gxs + gxm + gxe - (size - 1)
            if face:
                start += 1 if low else -1
            bnd = 0 if low else g.N[d]
            xs = [coord(bnd, True)] + [coord(start + m) for m in range(size - 1)]
            w = solve(size, [[(xs[m] - off_x) ** r for m in range(size)] for r in range(size)], [1.0 if r == 0 else 0.0 for r in range(size)])
            marker = Col(*[bnd if a == d else off_col.idx(a) for a in range(3)], boundary_location(off_col.loc, d), -(2 * d + (1 if low else 2)))
            add_point(st, marker, off_v * w[0])
            for m in range(size - 1):
                add_point(st, off_col.moved(d, start + m), off_v * w[m + 1])
        elif bc == BC_NEUMANN:  # :406-452
            start = gxs if low else gxs + gxm + gxe - (size - 1)
            bnd = 0 if low else g.N[d]
            bx = coord(bnd, True)
            xs = [off_x] + [coord(start + m) for m in range(size - 1)]
            w = solve(size, [[(xs[m] - bx) ** r for m in range(size)] for r in range(size)], [1.0 if r == 1 else 0.0 for r in range(size)])
            if abs(w[0]) < COEFF_ATOL:
                raise ValueError("Neumann BC coefficient for off-grid point is too small")
            marker = Col(*[bnd if a == d else off_col.idx(a) for a in range(3)], boundary_location(off_col.loc, d), -(2 * d + (1 if low else 2)))
            add_point(st, marker, off_v / w[0])
            for m in range(size - 1):
                add_point(st, off_col.moved(d, start + m), -off_v * w[m + 1] / w[0])
        else:
            raise ValueError("Unsupported boundary condition type")
    else:
        raise RuntimeError("Failed to remove all off-grid points")
    return remove_zero(st)


class Derivative(FD):
    """FLUCAFDDERIVATIVE, derivative.c:16-150"""

    def __init__(self, grid, direction: int, deriv_order=1, accu_order=1, input_loc="element", input_c=0, output_loc="element", output_c=0):
        super().__init__(grid, input_loc, input_c, output_loc, output_c)
        if direction >= grid.dim:
            raise ValueError("Cannot compute derivative in that direction on this DM")
        self.dir, self.deriv_order, self.accu_order = direction, deriv_order, accu_order
        fin, fout = use_face(input_loc, direction), use_face(output_loc, direction)
        valid = fin != fout
        for d in range(grid.dim):
            if d != direction and use_face(input_loc, d) != use_face(output_loc, d):
                valid = False
        if not (input_loc == output_loc or valid):
            raise ValueError("Cannot compute derivative between these locations")
        self.size = deriv_order + accu_order
        if self.size > MAX_STENCIL:
            raise ValueError("Required stencil size exceeds maximum")
        self.offset = -((self.size - 1) // 2)  # central (derivative.c:60): C integer division of a non-negative number
        if not fin and fout:
            self.offset -= 1
        self.fin, self.fout = fin, fout
        gxs, gxm, gxe = grid.ghost_corners(direction, fin)
        self.gxs, self.gxm, self.gxe = gxs, gxm, gxe
        self.v_start = gxs - (self.offset + self.size - 1)
        self.v_end = gxs + gxm + gxe - self.offset
        self.hp, self.hn = grid.end_spacings(direction, gxs, gxm)
        t = Term(input_loc=input_loc, input_c=input_c)
        t.deriv_order[direction], t.accu_order[direction] = deriv_order, accu_order
        self.terms = [t]
        self._cache: Dict[int, List[float]] = {}

    def weights(self, idx: int) -> List[float]:
        i = min(max(idx, self.v_start - 1), self.v_end)  # v_prev / v_next, derivative.c:84-91,128-130
        if i not in self._cache:
            g, d, n = self.grid, self.dir, self.size
            out_x = g.coordinate(d, i, self.fout, self.gxs, self.gxm, self.hp, self.hn)
            A = [[0.0] * n for _ in range(n)]
            for c in range(n):
                h = g.coordinate(d, i + self.offset + c, self.fin, self.gxs, self.gxm + self.gxe, self.hp, self.hn) - out_x
                for r in range(n):
                    A[r][c] = h**r
            b = [float(math.factorial(self.deriv_order)) if c == self.deriv_order else 0.0 for c in range(n)]
            self._cache[i] = solve(n, A, b)
        return self._cache[i]

    def _raw(self, i, j, k):
        ijk = (i, j, k)
        w = self.weights(ijk[self.dir])
        st = []
        for c in range(self.size):
            p = [i, j, k]
            p[self.dir] += self.offset + c
            st.append((Col(p[0], p[1], p[2], self.input_loc, self.input_c), w[c]))
        return st


def _merge_terms(dst: List[Term], src: List[Term]):
    keys = {t.key() for t in dst}
    for t in src:
        if t.key() not in keys:
            dst.append(t.copy())
            keys.add(t.key())


class Sum(FD):
    """FLUCAFDSUM, sum.c:3-52"""

    def __init__(self, operands: Sequence[FD]):
        first = operands[0]
        super().__init__(first.grid, first.output_loc, first.output_c, first.output_loc, first.output_c)
        for op in operands:
            if op.output_loc != self.output_loc or op.output_c != self.output_c:
                raise ValueError("All operands must have the same output stencil location and component")
        self.operands = list(operands)
        for op in operands:
            _merge_terms(self.terms, op.terms)

    def _raw(self, i, j, k):
        st = []
        for op in self.operands:
            for col, v in op.stencil_raw(i, j, k):
                add_point(st, col, v)
        return st


class Scale(FD):
    """FLUCAFDSCALE, scale.c:18-93: the operand's stencil times a constant or times a field sampled at the output point"""

    def __init__(self, operand: FD, constant: Optional[float] = None, vector=None, vec_loc: Optional[str] = None, vec_c: int = 0):
        super().__init__(operand.grid, operand.output_loc, operand.output_c, operand.output_loc, operand.output_c)
        self.operand, self.constant, self.vector = operand, constant, vector
        self.vec_loc, self.vec_c = (vec_loc if vec_loc is not None else operand.output_loc), vec_c
        if (constant is None) == (vector is None):
            raise ValueError("exactly one of constant / vector")
        self.terms = [t.copy() for t in operand.terms]

    def check(self):
        """the consistency checks of FlucaFDSetUp_Scale (scale.c:24-30), after options may have changed the locations"""
        if not (self.operand.output_c == self.input_c == self.output_c and self.operand.output_loc == self.input_loc == self.output_loc):
            raise ValueError("Cannot change component / location")
        if self.vector is not None and self.operand.output_loc != self.vec_loc:
            raise ValueError("Operand and vector must have the same location")
        return self

    def _raw(self, i, j, k):
        s = self.constant if self.constant is not None else self.vector(i, j, k, self.vec_loc, self.vec_c)
        return [(c, v * s) for c, v in self.operand.stencil_raw(i, j, k)]


class Composition(FD):
    """FLUCAFDCOMPOSITION, composition.c:3-76: outer(inner(.)); terms multiply (orders add, accuracies take the minimum)"""

    def __init__(self, inner: FD, outer: FD):
        super().__init__(inner.grid, inner.input_loc, inner.input_c, outer.output_loc, outer.output_c)
        if inner.output_c != outer.input_c or inner.output_loc != outer.input_loc:
            raise ValueError("Inner output must match outer input")
        self.inner, self.outer = inner, outer
        for ot in outer.terms:
            for it in inner.terms:
                t = Term(input_loc=it.input_loc, input_c=it.input_c)
                for d in range(3):
                    if it.deriv_order[d] == -1:
                        t.deriv_order[d] = ot.deriv_order[d]
                    elif ot.deriv_order[d] == -1:
                        t.deriv_order[d] = it.deriv_order[d]
                    else:
                        t.deriv_order[d] = it.deriv_order[d] + ot.deriv_order[d]
                    t.accu_order[d] = min(it.accu_order[d], ot.accu_order[d])
                _merge_terms(self.terms, [t])

    def _raw(self, i, j, k):
        st = []
        for oc, ov in self.outer.stencil_raw(i, j, k):
            if oc.c < 0:  # constant or boundary marker of the outer operator passes through
                add_point(st, oc, ov)
                continue
            for ic, iv in self.inner.stencil_raw(oc.i, oc.j, oc.k):
                add_point(st, ic, ov * iv)
        return st


# flux limiters psi(r), secondordertvdlimiter.c:3-82 (registered names: secondordertvd.c:19-36)
LIMITERS = {
    "superbee": lambda r: max(0.0, max(min(2.0 * r, 1.0), min(r, 2.0))),
    "minmod": lambda r: max(0.0, min(r, 1.0)),
    "mc": lambda r: max(0.0, min(min(2.0 * r, (1.0 + r) / 2.0), 2.0)),
    "vanleer": lambda r: (r + abs(r)) / (1.0 + abs(r)),
    "vanalbada": lambda r: 0.0 if r <= 0.0 else (r * r + r) / (r * r + 1.0),
    "barthjesperson": lambda r: 0.0 if r <= 0.0 else (1.0 + r) / 2.0 * min(1.0, min(4.0 * r / (1.0 + r), 4.0 / (1.0 + r))),
    "venkatakrishnan": lambda r: 0.0 if r <= 0.0 else (1.0 + r) / 2.0 * min(4.0 * r * (3.0 * r + 1.0) / (11.0 * r * r + 4.0 * r + 1.0), 4.0 * (r + 3.0) / (r * r + 4.0 * r + 11.0)),
    "koren": lambda r: max(0.0, min(min(2.0 * r, (1.0 + 2.0 * r) / 3.0), 2.0)),
    "upwind": lambda r: 0.0,
    "sou": lambda r: r,
    "quick": lambda r: (3.0 + r) / 4.0,
}


class SecondOrderTVD(FD):
    """FLUCAFDSECONDORDERTVD, secondordertvd.c:53-356: element values -> face value, upwind plus a limited correction that
    enters the stencil as a CONSTANT term evaluated on the current solution (deferred correction).
    velocity(i, j, k): advecting velocity at the face; phi(i, j, k): current solution at element centres (the reference reads
    its ghosted local array there; beyond a non-periodic end that array holds 0, as the stored outputs show)."""

    def __init__(self, grid, direction: int, input_c=0, output_c=0, limiter="superbee", velocity=None, phi=None):
        super().__init__(grid, "element", input_c, ("left", "down", "back")[direction], output_c)
        self.dir, self.limiter, self.velocity, self.phi = direction, LIMITERS[limiter], velocity, phi
        t = Term(input_loc="element", input_c=input_c)
        t.deriv_order[direction], t.accu_order[direction] = 0, 2  # interpolation, second order (:131-139)
        self.terms = [t]
        self._grad = None

    def _alpha(self, idx):
        """alpha_plus, alpha_minus of face idx (:86-128): the face's position between the two centres it separates"""
        g, d = self.grid, self.dir
        if (idx == 0 or idx == g.N[d]) and not g.periodic[d]:
            return 0.5, 0.5
        xf, xl, xr = g.array_coord(d, idx, True), g.array_coord(d, idx - 1, False), g.array_coord(d, idx, False)
        dx = xr - xl
        return ((xf - xl) / dx, (xr - xf) / dx) if abs(dx) > 1e-14 else (0.5, 0.5)

    def _phi(self, i, j, k):
        g = self.grid
        for d, idx in enumerate((i, j, k)[: g.dim]):
            gxs, gxm, _ = g.ghost_corners(d, False)
            if not gxs <= idx < gxs + gxm:
                return 0.0
        return self.phi(i, j, k)

    def _face_gradient(self, i, j, k):
        """ComputeFaceCenteredGradient_Private, :150-185: d phi / dx at a face from the first-order element -> face
        derivative with this operator's boundary condition types (fixed at set-up, :76-78) and current boundary values"""
        if self._grad is None:
            self._grad = Derivative(self.grid, self.dir, 1, 1, "element", self.input_c, self.output_loc, 0)
            self._grad.bcs = list(self.bcs)
        s = 0.0
        for col, v in self._grad.stencil(i, j, k):
            s += v * (self._phi(col.i, col.j, col.k) if col.c >= 0 else self.bcs[-col.c - 1][1])
        return s

    def _raw(self, i, j, k):
        g, d = self.grid, self.dir
        p = [i, j, k]
        idx = p[d]
        lo = list(p)
        lo[d] -= 1
        up = list(p)
        up[d] += 1
        vel = self.velocity(i, j, k)
        here, below = Col(p[0], p[1], p[2], "element", self.input_c), Col(lo[0], lo[1], lo[2], "element", self.input_c)
        at_prev = idx == 0 and not g.periodic[d]
        at_next = idx == g.N[d] and not g.periodic[d]
        const = Col(0, 0, 0, "element", CONSTANT)
        if vel > 0:  # upwind element is the one below the face
            if at_prev:
                return [(below, 0.5), (here, 0.5)]
            alpha = self._alpha(idx)[0]
            gfu, gfc = self._face_gradient(*lo), self._face_gradient(*p)
            psi = self.limiter(gfu / gfc if abs(gfc) > 1e-30 else 1.0)
            return [(below, 1.0), (const, alpha * psi * (self._phi(*p) - self._phi(*lo)))]
        if at_next:
            return [(here, 0.5), (below, 0.5)]
        alpha = self._alpha(idx)[1]
        gfu, gfc = self._face_gradient(*up), self._face_gradient(*p)
        psi = self.limiter(gfu / gfc if abs(gfc) > 1e-30 else 1.0)
        return [(here, 1.0), (const, alpha * psi * (self._phi(*lo) - self._phi(*p)))]


# ------------------------------------------------------------------ printing as the reference's tests do (fdtest.h, ex*.c)
def sort_key(item):
    """CompareDMStagStencil, fdtest.h:9-36: boundary markers last, then component, location, i, j, k"""
    col, _ = item
    return (1 if col.c < 0 else 0, col.c, LOC_ID[col.loc], col.i, col.j, col.k)


def fmt_g(v: float) -> str:
    """C printf("%g") as PetscPrintf emits it for real scalars: PETSc appends a '.' to a value printed without one"""
    s = "%g" % v
    if not any(ch in s for ch in ".einf"):
        s += "."
    return s


def print_stencil(st, dim: int) -> List[str]:
    out = [f"  ncols = {len(st)}"]
    for n, (col, v) in enumerate(sorted(st, key=sort_key)):
        if col.c == CONSTANT:  # ex7.c prints the deferred-correction term without a position
            out.append(f"  col[{n}]: constant, v={fmt_g(v)}")
            continue
        where = f"i={col.i}" if dim == 1 else (f"i={col.i}, j={col.j}" if dim == 2 else f"i={col.i}, j={col.j}, k={col.k}")
        comp = f"{BOUNDARY_NAMES[-col.c - 1]}_boundary" if -6 <= col.c < 0 else ("constant" if col.c == CONSTANT else str(col.c))
        out.append(f"  col[{n}]: {where}, loc={col.loc.upper()}, c={comp}, v={fmt_g(v)}")
    return out
