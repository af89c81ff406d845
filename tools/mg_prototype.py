"""Design-time prototype (numpy): pick the V-cycle components for the pressure Poisson PCG.

Cell-centred finite-volume Laplacian with homogeneous Neumann walls on a uniform grid (the
Schur complement S = -D*Gst of a wall-bounded case, flux form).  Compares prolongation
(piecewise constant vs trilinear) and smoother (damped Jacobi vs Chebyshev) by the PCG
iteration count to reach 1e-8.  Not used at run time; results are quoted in DESIGN.md.
"""
import sys
import numpy as np


def apply_A(p):
    q = np.zeros_like(p)
    for ax in range(p.ndim):
        d = np.diff(p, axis=ax)
        sl_lo = [slice(None)] * p.ndim
        sl_hi = [slice(None)] * p.ndim
        sl_lo[ax] = slice(0, -1)
        sl_hi[ax] = slice(1, None)
        q[tuple(sl_lo)] -= d
        q[tuple(sl_hi)] += d
    return q


def diag_A(shape):
    dg = np.zeros(shape)
    for ax in range(len(shape)):
        c = np.full(shape[ax], 2.0)
        c[0] = c[-1] = 1.0
        sh = [1] * len(shape)
        sh[ax] = shape[ax]
        dg = dg + c.reshape(sh)
    return dg


def restrict_sum(r):
    for ax in range(r.ndim):
        sl0 = [slice(None)] * r.ndim
        sl1 = [slice(None)] * r.ndim
        sl0[ax] = slice(0, None, 2)
        sl1[ax] = slice(1, None, 2)
        r = r[tuple(sl0)] + r[tuple(sl1)]
    return r


def prolong_const(e):
    for ax in range(e.ndim):
        e = np.repeat(e, 2, axis=ax)
    return e


def prolong_linear(e):
    for ax in range(e.ndim):
        n = e.shape[ax]
        lo = np.take(e, np.r_[0, np.arange(n - 1)], axis=ax)   # clamped neighbour below
        hi = np.take(e, np.r_[np.arange(1, n), n - 1], axis=ax)  # clamped neighbour above
        even = 0.75 * e + 0.25 * lo
        odd = 0.75 * e + 0.25 * hi
        out = np.stack([even, odd], axis=ax + 1)
        sh = list(e.shape)
        sh[ax] *= 2
        e = out.reshape(sh)
    return e


def smooth(x, b, dg, kind, nsweeps, scale):
    if kind == "jacobi":
        w = 0.8 if x.ndim == 2 else 6.0 / 7.0
        for _ in range(nsweeps):
            x = x + w * (b - scale * apply_A(x)) / (scale * dg)
        return x
    # Chebyshev on D^-1 A, eigenvalues in [lmax/alpha, lmax], lmax <= 2
    lmax, alpha = 2.0, 4.0 if kind == "cheb4" else 8.0
    lmin = lmax / alpha
    theta, delta = 0.5 * (lmax + lmin), 0.5 * (lmax - lmin)
    sigma = theta / delta
    rho = 1.0 / sigma
    r = (b - scale * apply_A(x)) / (scale * dg)
    d = r / theta
    for k in range(nsweeps):
        x = x + d
        if k == nsweeps - 1:
            break
        r = (b - scale * apply_A(x)) / (scale * dg)
        rho_new = 1.0 / (2.0 * sigma - rho)
        d = rho_new * rho * d + 2.0 * rho_new / delta * r
        rho = rho_new
    return x


def vcycle(b, kind, prol, nu, scale=1.0, over=1.0):
    shape = b.shape
    dg = diag_A(shape)
    if min(shape) <= 2 or any(s % 2 for s in shape):
        x = np.zeros_like(b)
        return smooth(x, b - b.mean(), dg, "jacobi", 30, scale)
    x = smooth(np.zeros_like(b), b, dg, kind, nu, scale)
    r = b - scale * apply_A(x)
    rc = restrict_sum(r)
    # coarse operator: rediscretised flux-form Laplacian; coarse h = 2h: face area x2^(d-1), distance x2
    cscale = scale * (2.0 ** (b.ndim - 1)) / 2.0
    ec = vcycle(rc, kind, prol, nu, cscale, over)
    e = prolong_const(ec) if prol == "const" else prolong_linear(ec)
    x = x + over * e
    return smooth(x, b, dg, kind, nu, scale)


def pcg(b, M, tol=1e-8, flexible=False, maxit=200):
    x = np.zeros_like(b)
    r = b.copy()
    z = M(r)
    p = z.copy()
    rz = np.vdot(r, z)
    n0 = np.linalg.norm(r)
    for it in range(1, maxit + 1):
        q = apply_A(p)
        a = rz / np.vdot(p, q)
        x += a * p
        rold = r.copy()
        r -= a * q
        if np.linalg.norm(r) <= tol * n0:
            return it
        z = M(r)
        rz_new = np.vdot(r, z)
        beta = (np.vdot(r - rold, z) if flexible else rz_new) / rz
        rz = rz_new
        p = z + beta * p
    return maxit


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    dim = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    rng = np.random.default_rng(0)
    b = rng.standard_normal((n,) * dim)
    b -= b.mean()
    print(f"grid {n}^{dim}")
    for kind, nu in (("jacobi", 2), ("jacobi", 3), ("cheb4", 2), ("cheb4", 3), ("cheb8", 3), ("cheb8", 4)):
        for prol, over in (("const", 1.0), ("const", 2.0), ("linear", 1.0)):
            M = lambda r: vcycle(r, kind, prol, nu, 1.0, over)
            it = pcg(b, M)
            itf = pcg(b, M, flexible=True)
            print(f"  smoother {kind:6s} nu={nu} prolong {prol:6s} over={over}: PCG its {it:3d}  flexible {itf:3d}")
