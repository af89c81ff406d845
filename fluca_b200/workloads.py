"""Workload definitions of the b200 NS type: the reference drivers' cases and BASELINE.json's configurations.

Product-side module (no oracle, no test imports): bench.py, the smoke test and the GPU parity tests build their NS objects
from here, the way the reference's drivers do (fluca/tests/cavity_flow/cavity_flow_2d.c:38-71).

Each case mirrors one of the reference's NS drivers:
  cavity2d  -- fluca/tests/cavity_flow/cavity_flow_2d.c:9-21,49-68 (walls + moving lid)
  cavity3d  -- fluca/tests/cavity_flow/cavity_flow_3d.c:51-77 (z in [0,0.5], BACK symmetry)
  tgv       -- fluca/tests/taylor_green_vortex/taylor_green_vortex.c:13-22,62-66 (Dirichlet-exact or periodic)
  channel*  -- inflow / pressure outlet / symmetry, the BC set of BASELINE configs 2 and 4
"""

from __future__ import annotations

import math

import numpy as np

BC_NONE, BC_VELOCITY, BC_PRESSURE_OUTLET, BC_PERIODIC, BC_SYMMETRY = range(5)


class Case:
    def __init__(self, name, n, lo, hi, rho, mu, dt, bcs, init=None, stretch=0.0):
        self.name, self.n, self.lo, self.hi = name, tuple(n), tuple(lo), tuple(hi)
        self.dim = len(n)
        self.rho, self.mu, self.dt = rho, mu, dt
        self.bcs = bcs  # list of dict(type=, velocity=callable|None, pressure=callable|None)
        self.init = init
        self.stretch = stretch

    def faces(self):
        out = []
        for d in range(self.dim):
            s = np.arange(self.n[d] + 1, dtype=np.float64) / self.n[d]
            if self.stretch and not self.bcs[2 * d]["type"] == BC_PERIODIC:
                # smooth non-uniform mapping, keeps end points
                s = s + self.stretch * np.sin(2 * np.pi * s) / (2 * np.pi)
            out.append(self.lo[d] + (self.hi[d] - self.lo[d]) * s)
        return out

    def centres(self):
        return [(f[:-1] + f[1:]) / 2.0 for f in self.faces()]

    def periodic(self):
        return [self.bcs[2 * d]["type"] == BC_PERIODIC for d in range(self.dim)]

    def shapes(self):
        n = self.n + (1,) * (3 - self.dim)
        per = self.periodic() + [False] * (3 - self.dim)
        cell = (n[2], n[1], n[0])
        face = []
        for d in range(self.dim):
            s = [n[2], n[1], n[0]]
            s[2 - d] += 0 if per[d] else 1
            face.append(tuple(s))
        return cell, face

    def initial_state(self, seed=None):
        """(v, U, p) -- zero, analytic (init callable), or seeded smooth random."""
        cell, face = self.shapes()
        xc = self.centres()
        xf = self.faces()
        v = np.zeros((self.dim,) + cell)
        U = [np.zeros(s) for s in face]
        p = np.zeros(cell)
        if self.init is not None:
            grids = _mesh(xc, self.dim)
            vv, pp = self.init(grids)
            for c in range(self.dim):
                v[c] = vv[c]
            p[...] = pp
            for d in range(self.dim):
                coords = list(xc)
                fd = xf[d] if not self.periodic()[d] else xf[d][:-1]
                coords[d] = fd
                vv, _ = self.init(_mesh(coords, self.dim))
                U[d][...] = vv[d]
        elif seed is not None:
            rng = np.random.default_rng(seed)
            grids = _mesh(xc, self.dim)
            for c in range(self.dim):
                v[c] = _smooth(rng, grids, self.lo, self.hi)
            p[...] = _smooth(rng, grids, self.lo, self.hi)
            for d in range(self.dim):
                coords = list(xc)
                coords[d] = xf[d] if not self.periodic()[d] else xf[d][:-1]
                U[d][...] = _smooth(rng, _mesh(coords, self.dim), self.lo, self.hi)
        return v, U, p


def _mesh(coords, dim):
    if dim == 2:
        Y, X = np.meshgrid(coords[1], coords[0], indexing="ij")
        return [X[None, :, :], Y[None, :, :]]
    Z, Y, X = np.meshgrid(coords[2], coords[1], coords[0], indexing="ij")
    return [X, Y, Z]


def _smooth(rng, grids, lo, hi):
    out = np.zeros(np.broadcast(*grids).shape)
    for _ in range(4):
        ph = rng.uniform(0, 2 * np.pi)
        arg = ph
        for d, g in enumerate(grids):
            arg = arg + rng.integers(1, 4) * 2 * np.pi * (g - lo[d]) / (hi[d] - lo[d])
        out = out + rng.uniform(-1, 1) * np.sin(arg)
    return out


def _const(*vals):
    """Constant boundary velocity.  Works point-wise (x = list of floats, what the oracle's callback
    passes) and vectorised (x = list of coordinate arrays, what the GPU host layer passes once per plane)."""

    def f(dim, t, x):
        shape = np.shape(x[0])
        if shape:
            return [np.full(shape, float(v)) for v in vals[:dim]]
        return vals[:dim]

    f.const = tuple(vals)
    f.vectorized = True
    f.time_independent = True
    return f


def _constp(val):
    """Constant boundary pressure, point-wise or vectorised (see _const)."""

    def f(dim, t, x):
        shape = np.shape(x[0])
        return np.full(shape, float(val)) if shape else float(val)

    f.constp = float(val)
    f.vectorized = True
    f.time_independent = True
    return f


def cavity2d(n=16, Re=100.0, dt=None):
    wall = dict(type=BC_VELOCITY, velocity=_const(0.0, 0.0), pressure=None)
    lid = dict(type=BC_VELOCITY, velocity=_const(1.0, 0.0), pressure=None)
    dt = dt if dt is not None else 0.5 / n
    return Case("cavity2d", (n, n), (0, 0), (1, 1), 1.0, 1.0 / Re, dt, [wall, wall, wall, lid])


def cavity3d(n=(8, 8, 4), Re=100.0, dt=None):
    wall = dict(type=BC_VELOCITY, velocity=_const(0.0, 0.0, 0.0), pressure=None)
    lid = dict(type=BC_VELOCITY, velocity=_const(1.0, 0.0, 0.0), pressure=None)
    sym = dict(type=BC_SYMMETRY, velocity=None, pressure=None)
    dt = dt if dt is not None else 0.5 / n[0]
    return Case("cavity3d", n, (0, 0, 0), (1, 1, 0.5), 1.0, 1.0 / Re, dt, [wall, wall, wall, lid, sym, wall])


def cavity3d_full(n=(8, 8, 8), Re=400.0, dt=None):
    wall = dict(type=BC_VELOCITY, velocity=_const(0.0, 0.0, 0.0), pressure=None)
    lid = dict(type=BC_VELOCITY, velocity=_const(1.0, 0.0, 0.0), pressure=None)
    dt = dt if dt is not None else 0.5 / n[0]
    return Case("cavity3d_full", n, (0, 0, 0), (1, 1, 1), 1.0, 1.0 / Re, dt, [wall, wall, wall, lid, wall, wall])


def tgv(n=8, periodic=False, rho=1.0, mu=1.0, dt=0.1):
    nu = mu / rho

    def vel(dim, t, x):
        e = math.exp(-2.0 * nu * t)
        return (math.sin(x[0]) * math.cos(x[1]) * e, -math.cos(x[0]) * math.sin(x[1]) * e)

    def init(g):
        X, Y = g
        return [np.sin(X) * np.cos(Y), -np.cos(X) * np.sin(Y)], rho / 4.0 * (np.cos(2 * X) + np.cos(2 * Y))

    bc = dict(type=BC_PERIODIC if periodic else BC_VELOCITY, velocity=vel, pressure=None)
    L = 2 * math.pi
    return Case("tgv_periodic" if periodic else "tgv", (n, n), (0, 0), (L, L), rho, mu, dt, [bc] * 4, init=init)


def channel2d(n=(24, 12), Re=100.0, dt=None, pout=0.0, time_dependent=False):
    """Inflow LEFT, pressure outlet RIGHT, symmetry DOWN/UP (BC set of BASELINE config 2)."""

    def inflow(dim, t, x):
        a = 1.0 + (0.1 * math.sin(3.0 * t) if time_dependent else 0.0)
        return (a * (1.0 + 0.2 * math.cos(2 * math.pi * x[1] / 4.0)), 0.0)

    def pressure(dim, t, x):
        return pout * (1.0 + (0.5 * math.sin(2.0 * t) if time_dependent else 0.0)) * (1.0 + 0.1 * x[1])

    inl = dict(type=BC_VELOCITY, velocity=inflow, pressure=None)
    out = dict(type=BC_PRESSURE_OUTLET, velocity=None, pressure=pressure)
    sym = dict(type=BC_SYMMETRY, velocity=None, pressure=None)
    dt = dt if dt is not None else 0.5 * 8.0 / n[0]
    return Case("channel2d", n, (-2, -2), (6, 2), 1.0, 1.0 / Re, dt, [inl, out, sym, sym])


def channel3d(n=(12, 8, 8), Re=300.0, dt=None, pout=0.0, periodic_z=False):
    """Inflow LEFT, outlet RIGHT, symmetry DOWN/UP, symmetry or periodic BACK/FRONT (config 4 BCs)."""

    def inflow(dim, t, x):
        return (1.0, 0.0, 0.0)

    def pressure(dim, t, x):
        return pout

    inl = dict(type=BC_VELOCITY, velocity=inflow, pressure=None)
    out = dict(type=BC_PRESSURE_OUTLET, velocity=None, pressure=pressure)
    sym = dict(type=BC_SYMMETRY, velocity=None, pressure=None)
    per = dict(type=BC_PERIODIC, velocity=None, pressure=None)
    z = per if periodic_z else sym
    dt = dt if dt is not None else 0.5 * 6.0 / n[0]
    return Case("channel3d", n, (-2, -2, -2), (4, 2, 2), 1.0, 1.0 / Re, dt, [inl, out, sym, sym, z, z])


def channel3d_z(n=(6, 6, 8), Re=300.0, dt=None, pout=0.0):
    """The same boundary set turned into z: inflow BACK, pressure outlet FRONT, symmetry elsewhere.  The outlet then sits on the last
    z-slab of a multi-rank run, and on the upper boundary where the reference's 3-D file forms operator T in its own way
    (cnlinearcart3d.c:2114)."""

    def inflow(dim, t, x):
        return (0.0, 0.0, 1.0 + 0.2 * math.cos(2 * math.pi * x[0] / 4.0))

    def pressure(dim, t, x):
        return pout * (1.0 + 0.1 * x[1])

    inl = dict(type=BC_VELOCITY, velocity=inflow, pressure=None)
    out = dict(type=BC_PRESSURE_OUTLET, velocity=None, pressure=pressure)
    sym = dict(type=BC_SYMMETRY, velocity=None, pressure=None)
    dt = dt if dt is not None else 0.5 * 6.0 / n[2]
    return Case("channel3d_z", n, (-2, -2, -2), (2, 2, 4), 1.0, 1.0 / Re, dt, [sym, sym, sym, sym, inl, out])


# ------------------------------------------------------------------ immersed-boundary marker sets
def cylinder_markers(centre, D, n, h, Ud=(0.0, 0.0), npts=4):
    """n markers equally spaced in angle on a circle (BASELINE config 2: theta_k = 2 pi k / n);
    volume weight = arc length x h."""
    th = 2 * np.pi * np.arange(n) / n
    X = np.stack([centre[0] + 0.5 * D * np.cos(th), centre[1] + 0.5 * D * np.sin(th)])
    dV = np.full(n, np.pi * D / n * h)
    return dict(X=X, Ud=np.tile(np.asarray(Ud, dtype=float)[:, None], (1, n)), dV=dV, npts=npts)


def sphere_markers(centre, D, n, h, Ud=(0.0, 0.0, 0.0), npts=4):
    """n markers on a Fibonacci-sphere lattice (BASELINE configs 4 and 5); volume weight = area / n x h."""
    k = np.arange(n) + 0.5
    z = 1.0 - 2.0 * k / n
    r = np.sqrt(np.maximum(0.0, 1.0 - z * z))
    ph = np.pi * (1.0 + 5.0**0.5) * k
    X = np.stack([centre[0] + 0.5 * D * r * np.cos(ph), centre[1] + 0.5 * D * r * np.sin(ph), centre[2] + 0.5 * D * z])
    dV = np.full(n, np.pi * D * D / n * h)
    return dict(X=X, Ud=np.tile(np.asarray(Ud, dtype=float)[:, None], (1, n)), dV=dV, npts=npts)


def multi_sphere_markers(centres, D, n_per, h, npts=4):
    """Several spheres, n_per Fibonacci markers each (BASELINE config 5)."""
    parts = [sphere_markers(c, D, n_per, h, npts=npts) for c in centres]
    return dict(X=np.concatenate([p["X"] for p in parts], axis=1), Ud=np.concatenate([p["Ud"] for p in parts], axis=1), dV=np.concatenate([p["dV"] for p in parts]), npts=npts)


# ------------------------------------------------------------------ BASELINE.json configurations
def cavity_bench_case(n, nz, Re=400.0):
    """BASELINE config 3: 3-D lid-driven cavity Re=400 on the unit cube, uniform h = 1/n, dt = 0.5 h, zero initial state.
    nz != n extends the box in z (weak-scaling runs keep an n^3 slab per GPU)."""
    c = cavity3d_full(n=(n, n, nz), Re=Re, dt=0.5 / n)
    c.hi = (1.0, 1.0, float(nz) / n)
    return c


def sphere_bench_case(n, nz, Re=300.0):
    """BASELINE config 4: flow past a sphere (D = 1 at the origin, U_inf = 1) by the immersed-boundary coupling on
    [-4,12] x [-8,8]^2 with n^3 cells (h = 16/n; 512^3 -> h = 1/32), inflow LEFT, pressure outlet RIGHT (p = 0), symmetry on
    the four side boundaries, dt = 0.5 h (CFL 0.5), initial state uniform U_inf.  nz != n extends the box in z."""
    c = channel3d(n=(n, n, nz), Re=Re, dt=0.5 * 16.0 / n)
    c.lo, c.hi = (-4.0, -8.0, -8.0), (12.0, 8.0, -8.0 + 16.0 * nz / n)
    for b in c.bcs:  # constant boundary data, evaluated once per plane
        for k in ("velocity", "pressure"):
            if b[k] is not None:
                b[k] = _const(1.0, 0.0, 0.0) if k == "velocity" else _constp(0.0)
    return c


def channel_bench_case(n=(2048, 1024, 1024), Re=300.0, periodic_z=True):
    """BASELINE config 5: multi-body channel, domain 32 x 16 x 16 at 2048 x 1024 x 1024 (h = 1/64; any n scales h), periodic x,
    no-slip walls at -y / +y, periodic (or symmetry) z, initial bulk velocity U = 1, dt = 0.5 h."""
    h = 32.0 / n[0]
    wall = dict(type=BC_VELOCITY, velocity=_const(0.0, 0.0, 0.0), pressure=None)
    per = dict(type=BC_PERIODIC, velocity=None, pressure=None)
    sym = dict(type=BC_SYMMETRY, velocity=None, pressure=None)
    z = per if periodic_z else sym
    return Case("channel_multibody", n, (0.0, 0.0, 0.0), (h * n[0], h * n[1], h * n[2]), 1.0, 1.0 / Re, 0.5 * h, [per, per, wall, wall, z, z])


def channel_sphere_centres(lo, hi, nspheres=80, D=1.0, seed=12345, gap=0.5):
    """Stratified jittered lattice of sphere centres with a minimum surface gap of gap * D between spheres and to the
    y walls; deterministic."""
    rng = np.random.default_rng(seed)
    L = [hi[d] - lo[d] for d in range(3)]
    # lattice with roughly cubic cells holding >= nspheres sites
    vol = L[0] * L[1] * L[2] / nspheres
    a = vol ** (1.0 / 3.0)
    m = [max(1, int(L[d] / a)) for d in range(3)]
    while m[0] * m[1] * m[2] < nspheres:
        m[int(np.argmax([L[d] / m[d] for d in range(3)]))] += 1
    sites = [(i, j, k) for k in range(m[2]) for j in range(m[1]) for i in range(m[0])]
    pick = rng.permutation(len(sites))[:nspheres]
    out = []
    for s in sorted(pick):
        i = sites[s]
        c = []
        for d in range(3):
            cell = L[d] / m[d]
            slack = max(0.0, 0.5 * (cell - (1.0 + gap) * D))
            c.append(lo[d] + (i[d] + 0.5) * cell + rng.uniform(-slack, slack))
        c[1] = min(max(c[1], lo[1] + (0.5 + gap) * D), hi[1] - (0.5 + gap) * D)
        out.append(tuple(c))
    return out


def uniform_inflow_state(case, U=1.0, slab=None):
    """(v, U, p): uniform stream along x, zero pressure.  slab = (k0, nzl, last_z): only that z-slab of a 3-D case (what one
    rank holds; the z-face field of the last wall rank has one plane more) -- a global state of BASELINE config 5 is 120 GB."""
    cell, face = case.shapes()
    if slab is not None:
        k0, nzl, last = slab
        cell = (nzl,) + cell[1:]
        face = [(nzl + (1 if (d == 2 and last) else 0),) + f[1:] for d, f in enumerate(face)]
    v = np.zeros((3,) + cell)
    v[0] = U
    Uf = [np.zeros(s) for s in face]
    Uf[0][...] = U
    return v, Uf, np.zeros(cell)


# ------------------------------------------------------------------ NS object of a case, the way the reference drivers build it
def make_ns(case, library=None, mode="coupled", comm=None, **opts):
    """cavity_flow_2d.c:38-71 with -ns_type b200."""
    import fluca_b200 as fb

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


def set_initial(ns, state):
    import fluca_b200 as fb

    v, U, p = state
    fb.NSSetSolutionSubVector(ns, fb.NS_FIELD_VELOCITY, v)
    fb.NSSetSolutionSubVector(ns, fb.NS_FIELD_FACE_NORMAL_VELOCITY, U)
    fb.NSSetSolutionSubVector(ns, fb.NS_FIELD_PRESSURE, p)


def set_initial_slab(ns, state):
    """Multi-rank form of set_initial: every rank writes its own z-slab of a global state."""
    import fluca_b200 as fb

    s = fb.NSB200GetSolver(ns)
    v, U, p = state
    k0, nzl = s.k0, s.nzl
    s.set_state(v=v[:, k0 : k0 + nzl], U=[U[0][k0 : k0 + nzl], U[1][k0 : k0 + nzl], U[2][k0 : k0 + nzl + (1 if s.last_z else 0)]], p=p[k0 : k0 + nzl], phalf=p[k0 : k0 + nzl])
