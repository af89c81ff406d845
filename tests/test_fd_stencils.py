"""FlucaFD through the product's C ABI (include/fluca_b200.h "FlucaFD", csrc/fd.cu) against the
reference's own golden outputs: every test of fluca/tests/fd/ex*.c with a stored output (tests/golden/fd_stencils.json) is
rebuilt from its command line with the FlucaFD-named host API (fluca_b200/fd.py) and must print the same lines the
reference's harness compares byte for byte.  The stencil layer is host code, so these run on the CUDA product library
itself without a GPU (and on the host-emulation build of the same sources)."""
import json
import math
import os

import numpy as np
import pytest

import fluca_b200 as fb
from fluca_b200 import fd as FD
from tests import parity

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "fd_stencils.json")))


def options(args):
    return {args[k][1:]: args[k + 1] for k in range(0, len(args), 2)}


def fmt_g(v):
    """printf("%g") as PetscPrintf emits it: a '.' is appended when the number was printed without one"""
    s = "%g" % v
    return s if any(ch in s for ch in ".einf") else s + "."


def print_stencil(st, dim):
    """SortStencil + the printing loop of ex*.c (fdtest.h:9-44): boundary / constant markers last, then c, loc, i, j, k"""
    st = sorted(st, key=lambda e: (1 if e[0][4] < 0 else 0, e[0][4], e[0][3], e[0][0], e[0][1], e[0][2]))
    out = [f"  ncols = {len(st)}"]
    for n, ((i, j, k, loc, c), v) in enumerate(st):
        if c == FD.FLUCAFD_CONSTANT:
            out.append(f"  col[{n}]: constant, v={fmt_g(v)}")
            continue
        where = f"i={i}" if dim == 1 else (f"i={i}, j={j}" if dim == 2 else f"i={i}, j={j}, k={k}")
        comp = f"{FD.BOUNDARY_NAMES[-c - 1]}_boundary" if c < 0 else str(c)
        out.append(f"  col[{n}]: {where}, loc={FD.LOCATION_NAMES[loc]}, c={comp}, v={fmt_g(v)}")
    return out


def grid_from(lib, o, dim):
    n = [int(o.get(f"stag_grid_{a}", 8)) for a in "xyz"[:dim]]
    per = [o.get(f"stag_boundary_type_{a}", "none") == "periodic" for a in "xyz"[:dim]]
    return FD.FDGrid.uniform(n, [0.0] * dim, [1.0] * dim, per, int(o.get("stag_stencil_width", 1)), library=lib)


def set_from_options(fd, o, prefix, locs):
    """FlucaFDSetFromOptions (fdopts.c:83-99); locs = the (input_loc, input_c, output_loc, output_c) the object was created with"""
    il, ic, ol, oc = locs
    il = FD.LOCATION_BY_NAME[o[f"{prefix}flucafd_input_loc"]] if f"{prefix}flucafd_input_loc" in o else il
    ol = FD.LOCATION_BY_NAME[o[f"{prefix}flucafd_output_loc"]] if f"{prefix}flucafd_output_loc" in o else ol
    fd.SetLocations(il, int(o.get(f"{prefix}flucafd_input_c", ic)), ol, int(o.get(f"{prefix}flucafd_output_c", oc)))
    for b, name in enumerate(FD.BOUNDARY_NAMES):
        if f"{prefix}flucafd_{name}_bc_type" in o:
            fd.SetBoundaryCondition(b, FD.BC_BY_NAME[o[f"{prefix}flucafd_{name}_bc_type"]])
    return (il, ic, ol, oc)


def derivative(g, o, prefix, direction):
    d = {"x": 0, "y": 1, "z": 2}.get(o.get(f"{prefix}flucafd_dir", ""), direction)
    fd = FD.FlucaFDDerivativeCreate(g, d, int(o.get(f"{prefix}flucafd_deriv_order", 1)), int(o.get(f"{prefix}flucafd_accu_order", 1)), FD.DMSTAG_ELEMENT, 0, FD.DMSTAG_ELEMENT, 0)
    locs = set_from_options(fd, o, prefix, (FD.DMSTAG_ELEMENT, 0, FD.DMSTAG_ELEMENT, 0))
    return fd.SetUp(), locs


def run_ex1(lib, o):
    g = grid_from(lib, o, 1)
    fd, _ = derivative(g, o, "", 0)
    i = int(o.get("i", g.n[0] // 2))
    return [f"Stencil at i={i}:"] + print_stencil(fd.GetStencil(i), 1)


def run_ex2(lib, o):
    g = grid_from(lib, o, 3)
    ops, locs = zip(*[derivative(g, o, f"{a}_", d) for d, a in enumerate("xyz")])
    s = FD.FlucaFDSumCreate(ops)
    set_from_options(s, o, "sum_", (locs[0][2], locs[0][3], locs[0][2], locs[0][3]))
    s.SetUp()
    idx = [int(o.get(a, g.n[d] // 2)) for d, a in enumerate("ijk")]
    return [f"Sum stencil at (i,j,k)=({idx[0]},{idx[1]},{idx[2]}):"] + print_stencil(s.GetStencil(*idx), 3)


def run_ex3(lib, o):
    g = grid_from(lib, o, 1)
    deriv, locs = derivative(g, o, "deriv_", 0)
    out_loc, out_c = locs[2], locs[3]
    if o.get("const", "true") == "true":
        sc = FD.FlucaFDScaleCreateConstant(deriv, float(o.get("scale_flucafd_constant", 1.0)))
    else:
        vec_loc = FD.LOCATION_BY_NAME[o["scale_flucafd_vec_loc"]] if "scale_flucafd_vec_loc" in o else out_loc
        n = g.field_shape(vec_loc)[0]
        field = np.array([2.0 * i if vec_loc == FD.DMSTAG_LEFT else 2.0 * i + 1.0 for i in range(n)])  # FillScaleVector, ex3.c:75-95
        sc = FD.FlucaFDScaleCreateVector(deriv, field, vec_loc, 0)
    set_from_options(sc, o, "scale_", (out_loc, out_c, out_loc, out_c))
    sc.SetUp()
    i = int(o.get("i", g.n[0] // 2))
    return [f"Scaled stencil at i={i}:"] + print_stencil(sc.GetStencil(i), 1)


def run_ex4(lib, o):
    g = grid_from(lib, o, 2)
    inner, li = derivative(g, o, "inner_", 0)
    outer, lo = derivative(g, o, "outer_", 0)
    comp = FD.FlucaFDCompositionCreate(inner, outer)
    set_from_options(comp, o, "comp_", (li[0], li[1], lo[2], lo[3]))
    comp.SetUp()
    idx = [int(o.get(a, g.n[d] // 2)) for d, a in enumerate("ij")]
    header = f"Sum stencil at (i,j)=({idx[0]},{idx[1]}):"  # sic: ex4.c prints the header of ex2
    return [header] + print_stencil(comp.GetStencil(idx[0], idx[1]), 2)


def run_ex7(lib, o):
    g = grid_from(lib, {}, 1)  # ex7 does not call DMSetFromOptions
    n = g.n[0]
    tvd = FD.FlucaFDSecondOrderTVDCreate(g, FD.FLUCAFD_X, 0, 0)
    types = [FD.FLUCAFD_BC_DIRICHLET, FD.FLUCAFD_BC_DIRICHLET]
    for b, name in enumerate(("left", "right")):
        if f"flucafd_{name}_bc_type" in o:
            types[b] = FD.BC_BY_NAME[o[f"flucafd_{name}_bc_type"]]
    if "flucafd_limiter" in o:
        FD.FlucaFDSecondOrderTVDSetLimiter(tvd, o["flucafd_limiter"])
    tvd.SetBoundaryCondition(0, types[0], 0.0)
    tvd.SetBoundaryCondition(1, types[1], 1.0)
    tvd.SetUp()
    values = [{FD.FLUCAFD_BC_DIRICHLET: 0.0, FD.FLUCAFD_BC_NEUMANN: math.pi / 2.0}, {FD.FLUCAFD_BC_DIRICHLET: 1.0, FD.FLUCAFD_BC_NEUMANN: 0.0}]
    for b in (0, 1):  # boundary values follow the chosen types, set after FlucaFDSetUp (ex7.c:84-110)
        tvd.SetBoundaryCondition(b, types[b], values[b].get(types[b], 0.0))
    FD.FlucaFDSecondOrderTVDSetVelocity(tvd, np.ones(n + 1))
    FD.FlucaFDSecondOrderTVDSetCurrentSolution(tvd, np.sin(np.pi * (np.arange(n) + 0.5) / n / 2.0))
    i = int(o.get("i", n // 2))
    return [f"Stencil at i={i}:"] + print_stencil(tvd.GetStencil(i), 1)


RUNNERS = {"ex1": run_ex1, "ex2": run_ex2, "ex3": run_ex3, "ex4": run_ex4, "ex7": run_ex7}
NAMES = sorted(GOLD)


@pytest.fixture(scope="module", params=["product", "hostemu"])
def lib(request):
    if request.param == "product":
        L = fb._lib.load()  # the CUDA build: this layer needs no device
        assert L.fluca_b200_is_host_emulation() == 0
        return L
    return parity.hostemu_library()


@pytest.mark.parametrize("name", NAMES)
def test_c_abi_prints_the_reference_golden_output(lib, name):
    case = GOLD[name]
    got = RUNNERS[case["program"]](lib, options(case["args"]))
    assert got == case["output"], "\n".join(["", "got:"] + got + ["expected:"] + case["output"])


def test_all_52_stored_outputs_are_covered():
    assert len(NAMES) == 52 and {v["program"] for v in GOLD.values()} == set(RUNNERS)


def test_errors_follow_the_reference(lib):
    g = FD.FDGrid.uniform([8, 8], [0.0, 0.0], [1.0, 1.0], library=lib)
    d = FD.FlucaFDDerivativeCreate(g, FD.FLUCAFD_X, 1, 2, FD.DMSTAG_ELEMENT, 0, FD.DMSTAG_ELEMENT, 0)
    with pytest.raises(FD.FlucaFDError, match="not set up"):  # operands are set up before their parents (sum.c:120-123)
        FD.FlucaFDSumCreate([d])
    with pytest.raises(FD.FlucaFDError, match="not setup"):  # fdapply.c:64
        d.GetStencil(4, 4)
    bad = FD.FlucaFDDerivativeCreate(g, FD.FLUCAFD_Z, 1, 2, FD.DMSTAG_ELEMENT, 0, FD.DMSTAG_ELEMENT, 0)
    with pytest.raises(FD.FlucaFDError, match="direction"):  # derivative.c:23
        bad.SetUp()
    cross = FD.FlucaFDDerivativeCreate(g, FD.FLUCAFD_X, 1, 2, FD.DMSTAG_ELEMENT, 0, FD.DMSTAG_DOWN, 0)
    with pytest.raises(FD.FlucaFDError, match="locations"):  # derivative.c:27-37: only the derivative direction may change face-ness
        cross.SetUp()
    with pytest.raises(FD.FlucaFDError, match="Invalid stencil location"):  # fdutils.c:16-34
        FD.FlucaFDDerivativeCreate(g, FD.FLUCAFD_X, 1, 2, 15, 0, FD.DMSTAG_ELEMENT, 0)  # DMSTAG_RIGHT
    d.SetUp()
    up = FD.FlucaFDDerivativeCreate(g, FD.FLUCAFD_Y, 1, 2, FD.DMSTAG_ELEMENT, 0, FD.DMSTAG_DOWN, 0).SetUp()
    with pytest.raises(FD.FlucaFDError, match="same output"):  # sum.c:17-18
        FD.FlucaFDSumCreate([d, up]).SetUp()
    t = FD.FlucaFDSecondOrderTVDCreate(g, FD.FLUCAFD_X)
    with pytest.raises(FD.FlucaFDError, match="Unknown limiter"):
        FD.FlucaFDSecondOrderTVDSetLimiter(t, "nosuch")
