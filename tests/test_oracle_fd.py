"""The FlucaFD oracle (oracle/fd_oracle.py) against the reference's own golden outputs: every test of
fluca/tests/fd/ex*.c that stores an expected output (tests/golden/fd_stencils.json, extracted by
tests/golden/make_fd_stencils.py) is rebuilt from its command-line arguments and must print the same lines -- the
reference's harness compares these files byte for byte."""
import json
import os

import pytest

from oracle import fd_oracle as F

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "fd_stencils.json")))
BOUNDARIES = ["left", "right", "down", "up", "back", "front"]


def options(args):
    o, k = {}, 0
    while k < len(args):
        assert args[k].startswith("-")
        o[args[k][1:]] = args[k + 1]
        k += 2
    return o


def grid_from(o, dim, n=8, lo=0.0, hi=1.0):
    """DMSetFromOptions of the DMStag the programs create: -stag_grid_{x,y,z}, -stag_boundary_type_{x,y,z}, -stag_stencil_width"""
    N = [int(o.get(f"stag_grid_{a}", n)) for a in "xyz"[:dim]]
    per = [o.get(f"stag_boundary_type_{a}", "none") == "periodic" for a in "xyz"[:dim]]
    return F.Grid(N, [lo] * dim, [hi] * dim, per, int(o.get("stag_stencil_width", 1)))


def fd_options(fd, o, prefix=""):
    """FlucaFDSetFromOptions, fdopts.c:83-99: locations, components and the six boundary condition types"""
    fd.input_loc = o.get(f"{prefix}flucafd_input_loc", fd.input_loc)
    fd.output_loc = o.get(f"{prefix}flucafd_output_loc", fd.output_loc)
    fd.input_c = int(o.get(f"{prefix}flucafd_input_c", fd.input_c))
    fd.output_c = int(o.get(f"{prefix}flucafd_output_c", fd.output_c))
    for b, name in enumerate(BOUNDARIES):
        key = f"{prefix}flucafd_{name}_bc_type"
        if key in o:
            fd.set_bc(b, o[key])
    return fd


def derivative(grid, o, prefix, direction, input_loc="element", output_loc="element"):
    return F.Derivative(
        grid,
        {"x": 0, "y": 1, "z": 2}.get(o.get(f"{prefix}flucafd_dir", ""), direction),
        int(o.get(f"{prefix}flucafd_deriv_order", 1)),
        int(o.get(f"{prefix}flucafd_accu_order", 1)),
        o.get(f"{prefix}flucafd_input_loc", input_loc),
        int(o.get(f"{prefix}flucafd_input_c", 0)),
        o.get(f"{prefix}flucafd_output_loc", output_loc),
        int(o.get(f"{prefix}flucafd_output_c", 0)),
    )


def run_ex1(o):
    """fluca/tests/fd/ex1.c:21-46: 1-D, 8 elements on [0, 1], one derivative operator, stencil at -i (default M / 2)"""
    g = grid_from(o, 1)
    fd = fd_options(derivative(g, o, "", 0), o)
    i = int(o.get("i", g.N[0] // 2))
    return [f"Stencil at i={i}:"] + F.print_stencil(fd.stencil(i, 0, 0), 1)


def run_ex2(o):
    """fluca/tests/fd/ex2.c:21-69: 3-D 8^3 on the unit cube, the sum of one derivative per direction"""
    g = grid_from(o, 3)
    ops = [fd_options(derivative(g, o, f"{a}_", d), o, f"{a}_") for d, a in enumerate("xyz")]
    s = fd_options(F.Sum(ops), o, "sum_")
    idx = [int(o.get(a, g.N[d] // 2)) for d, a in enumerate("ijk")]
    return [f"Sum stencil at (i,j,k)=({idx[0]},{idx[1]},{idx[2]}):"] + F.print_stencil(s.stencil(*idx), 3)


def run_ex3(o):
    """fluca/tests/fd/ex3.c:22-72: 1-D derivative scaled by a constant (default 1, -scale_flucafd_constant) or by the field of
    FillScaleVector (ex3.c:75-95): 2 i at the LEFT points, 2 i + 1 at the element centres"""
    g = grid_from(o, 1)
    deriv = fd_options(derivative(g, o, "deriv_", 0), o, "deriv_")
    if o.get("const", "true") == "true":
        sc = F.Scale(deriv, constant=float(o.get("scale_flucafd_constant", 1.0)))
    else:
        sc = F.Scale(deriv, vector=lambda i, j, k, loc, c: float(2 * i if loc == "left" else 2 * i + 1), vec_loc=o.get("scale_flucafd_vec_loc", deriv.output_loc))
    fd_options(sc, o, "scale_").check()
    i = int(o.get("i", g.N[0] // 2))
    return [f"Scaled stencil at i={i}:"] + F.print_stencil(sc.stencil(i, 0, 0), 1)


def run_ex4(o):
    """fluca/tests/fd/ex4.c:20-75: 2-D 8 x 8 on the unit square, outer(inner(.)) of two derivative operators"""
    g = grid_from(o, 2)
    inner = fd_options(derivative(g, o, "inner_", 0), o, "inner_")
    outer = fd_options(derivative(g, o, "outer_", 0), o, "outer_")
    comp = fd_options(F.Composition(inner, outer), o, "comp_")
    idx = [int(o.get(a, g.N[d] // 2)) for d, a in enumerate("ij")]
    header = f"Sum stencil at (i,j)=({idx[0]},{idx[1]}):"  # sic: ex4.c prints the header of ex2
    return [header] + F.print_stencil(comp.stencil(idx[0], idx[1], 0), 2)


def run_ex7(o):
    """fluca/tests/fd/ex7.c:22-118: 1-D TVD interpolation of phi = sin(pi x / 2) to the faces, velocity +1 everywhere,
    Dirichlet (0, 1) by default; boundary values follow the chosen types (pi / 2 and 0 for Neumann)"""
    import math

    g = grid_from({}, 1)  # ex7 does not call DMSetFromOptions
    h = 1.0 / g.N[0]
    tvd = F.SecondOrderTVD(g, 0, limiter=o.get("flucafd_limiter", "superbee"), velocity=lambda i, j, k: 1.0, phi=lambda i, j, k: math.sin(math.pi * ((i + 0.5) * h) / 2.0))
    tvd.set_bc(0, "dirichlet", 0.0)
    tvd.set_bc(1, "dirichlet", 1.0)
    fd_options(tvd, o)
    values = {0: {"dirichlet": 0.0, "neumann": math.pi / 2.0}, 1: {"dirichlet": 1.0, "neumann": 0.0}}
    for b in (0, 1):
        kind = tvd.bcs[b][0]
        tvd.set_bc(b, kind, values[b].get(kind, tvd.bcs[b][1]))
    i = int(o.get("i", g.N[0] // 2))
    return [f"Stencil at i={i}:"] + F.print_stencil(tvd.stencil(i, 0, 0), 1)


RUNNERS = {"ex1": run_ex1, "ex2": run_ex2, "ex3": run_ex3, "ex4": run_ex4, "ex7": run_ex7}
NAMES = sorted(k for k, v in GOLD.items() if v["program"] in RUNNERS)


@pytest.mark.parametrize("name", NAMES)
def test_oracle_prints_the_reference_golden_output(name):
    case = GOLD[name]
    got = RUNNERS[case["program"]](options(case["args"]))
    assert got == case["output"], "\n".join(["", "got:"] + got + ["expected:"] + case["output"])


def test_every_stored_golden_is_accounted_for():
    """all 52 reference tests that store an output are reproduced (ex5 / ex6 store none)"""
    progs = {v["program"] for v in GOLD.values()}
    assert progs == set(RUNNERS) == {"ex1", "ex2", "ex3", "ex4", "ex7"} and len(GOLD) == len(NAMES) == 52


def test_apply_is_exact_on_polynomials_up_to_the_boundaries():
    """FlucaFDApply (fdapply.c:47-121) point by point: a second derivative of accuracy 2 with Dirichlet data differentiates a
    cubic exactly at every element, the one-sided boundary closures included; with Neumann data at the right end too."""
    g = F.Grid([8], [0.0], [1.0])
    f = lambda x: 1.0 + 2.0 * x - 3.0 * x * x + 0.5 * x**3  # noqa: E731
    d2 = lambda x: -6.0 + 3.0 * x  # noqa: E731
    xc = lambda i: (i + 0.5) / 8.0  # noqa: E731
    for right in ("dirichlet", "neumann"):
        fd = F.Derivative(g, 0, 2, 2)
        fd.set_bc(0, "dirichlet", f(0.0))
        fd.set_bc(1, right, f(1.0) if right == "dirichlet" else 2.0 - 6.0 + 1.5)
        for i in range(8):
            got = fd.apply_point(i, 0, 0, lambda col: f(xc(col.i)))
            assert abs(got - d2(xc(i))) < 1e-9, (right, i, got)


def test_limiters_are_second_order_tvd_where_the_reference_says_so():
    """Sweby region: 0 <= psi <= min(2 r, 2), psi(1) = 1, psi(r <= 0) = 0 for the TVD limiters; upwind / sou / quick are
    the three the reference itself flags as not second-order TVD (secondordertvd.c:141-145)."""
    for name, psi in F.LIMITERS.items():
        if name in ("upwind", "sou", "quick"):
            continue
        assert abs(psi(1.0) - 1.0) < 1e-14, name
        for r in [-2.0, -0.5, 0.0, 0.1, 0.5, 0.9, 1.5, 3.0, 10.0]:
            v = psi(r)
            assert -1e-14 <= v <= min(2.0 * max(r, 0.0), 2.0) + 1e-14, (name, r, v)
