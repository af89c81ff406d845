"""Pin the oracle's stencil formulas against the reference's own fd golden outputs.

tests/golden/fd_coefficients.json is extracted (tests/golden/make_fd_coefficients.py) from
/root/reference/fluca/tests/fd/output/*.out -- the byte-exact expected outputs of the reference's
registered ctest cases, on a uniform 8-cell grid (h = 1/8).  SURVEY.md section 4 maps each golden
to the hand-derived NS formula of fluca/src/ns/utils/cartdiscret.c it cross-pins.  The goldens
print with %g (6 significant digits), hence rel=2e-6.
"""
import json
import os

import numpy as np
import pytest

from oracle import oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "fd_coefficients.json")) as f:
    GOLD = json.load(f)

H = 1.0 / 8.0
XF = np.arange(9) * H
XC = (XF[:-1] + XF[1:]) / 2


def gold(case, key="i"):
    return {(c["index"][key], c["loc"], c["c"]): c["v"] for c in GOLD[case]["cols"]}


def approx(v):
    return pytest.approx(v, rel=2e-6)


def test_second_derivative_central():  # cartdiscret.c:210
    off, w = O.formula("d2_central", [XC[3], XF[4], XC[4], XF[5], XC[5]])
    g = gold("ex1_second_deriv")
    for o, wt in zip(off, w):
        assert wt == approx(g[(4 + o, "ELEMENT", "0")])
    assert w == approx([64.0, -128.0, 64.0])


def test_second_derivative_left_dirichlet():  # cartdiscret.c:167 + cnlinearcart2d.c:494-498
    off, w = O.formula("d2_fwd_dirichlet", [XF[0], XC[0], XC[1], XC[2]])
    g = gold("ex1_second_deriv_left_bc_dirichlet")
    for o, wt in zip(off, w):
        assert wt == approx(g[(0 + o, "ELEMENT", "0")])
    h1, h2, h3 = XC[0] - XF[0], XC[1] - XC[0], XC[2] - XC[0]
    bcw = 2.0 * (h2 + h3) / (h1 * (h1 + h2) * (h1 + h3))
    assert bcw == approx(g[(0, "LEFT", "left_boundary")])


def test_second_derivative_right_dirichlet():  # cartdiscret.c:262 ; golden is scaled by 1.5
    off, w = O.formula("d2_bwd_dirichlet", [XC[5], XC[6], XC[7], XF[8]])
    g = gold("ex3_second_deriv_right_bc_dirichlet_scale_const")
    for o, wt in zip(off, w):
        assert 1.5 * wt == approx(g[(7 + o, "ELEMENT", "0")])
    h1, h2, h3 = XF[8] - XC[7], XC[7] - XC[6], XC[7] - XC[5]
    assert 1.5 * 2.0 * (h2 + h3) / (h1 * (h1 + h2) * (h1 + h3)) == approx(g[(8, "LEFT", "right_boundary")])


def test_second_derivative_up_neumann():  # cartdiscret.c:286 (symmetry / outlet walls)
    off, w = O.formula("d2_bwd_neumann", [XC[6], XF[7], XC[7], XF[8]])
    cols = GOLD["ex2_all_second_deriv_up_bc_neumann"]["cols"]
    centre = [c["v"] for c in cols if c["index"] == {"i": 4, "j": 7, "k": 4}][0]
    below = [c["v"] for c in cols if c["index"] == {"i": 4, "j": 6, "k": 4}][0]
    assert dict(zip(off, w))[-1] == approx(below)
    _, wc = O.formula("d2_central", [XC[3], XF[4], XC[4], XF[5], XC[5]])
    assert 2 * wc[1] + dict(zip(off, w))[0] == approx(centre)  # -(128 + 128 + 64)


def test_first_derivative_central():  # cartdiscret.c:64, cell-centred G
    off, w = O.formula("d1_central", [XC[3], XC[5]])
    g = gold("ex1_first_deriv")
    for o, wt in zip(off, w):
        assert wt == approx(g[(4 + o, "ELEMENT", "0")])


def test_face_normal_first_derivative():  # cartdiscret.c:444, Gst
    off, w = O.formula("fn_central", [XC[3], XC[4]])
    g = gold("ex1_first_deriv_input_loc_elem_output_loc_left")
    for o, wt in zip(off, w):
        assert wt == approx(g[(4 + o, "ELEMENT", "0")])
    gp = gold("ex1_first_deriv_input_loc_elem_output_loc_left_left_bc_periodic")
    off, w = O.formula("fn_central", [XC[0] - H, XC[0]])
    for o, wt in zip(off, w):
        assert wt == approx(gp[(0 + o, "ELEMENT", "0")])


def test_divergence_coefficients():  # cnlinearcart3d.c:2314-2408
    cols = GOLD["ex2_all_first_deriv_input_loc_face_output_loc_elem"]["cols"]
    vals = sorted(c["v"] for c in cols)
    assert vals == approx([-8.0] * 3 + [8.0] * 3)
    # the oracle's D on the same 8^3 grid
    import tests.cases as cases

    case = cases.cavity3d_full(n=(8, 8, 8))
    o = cases.make_oracle(case)
    D = o.matrix("D")
    row = D.getrow(4 + 8 * (4 + 8 * 4))
    assert sorted(row.data) == approx([-8.0] * 3 + [8.0] * 3)


def test_unused_forward_nocond_second_derivative():  # cartdiscret.c:139 (unused by NS, still restated)
    off, w = O.formula("d2_fwd_nocond", [XC[0], XC[1], XC[2], XC[3]])
    g = gold("ex1_second_deriv_left_bc_none")
    for o, wt in zip(off, w):
        assert wt == approx(g[(0 + o, "ELEMENT", "0")])


def test_boundary_formulas_uniform_values():
    """Closed-form values quoted in SURVEY.md Appendix A for uniform h."""
    h = H
    _, w = O.formula("d1_fwd_nocond", [XC[0], XC[1], XC[2]])
    assert w == approx([-1.5 / h, 2.0 / h, -0.5 / h])
    _, w = O.formula("d2_fwd_dirichlet", [XF[0], XC[0], XC[1], XC[2]])
    assert w == approx([-5 / h**2, 2 / h**2, -0.2 / h**2])
    _, w = O.formula("lin_fwd_extrap", [XF[0], XC[0], XC[1]])
    assert w == approx([9.0 / 8.0, -1.0 / 8.0])
    _, w = O.formula("lin_interp", [XC[0], XF[1], XC[1]])
    assert w == approx([0.5, 0.5])
    # exactness: first/second derivative formulas differentiate quadratics exactly on a non-uniform grid
    xs = np.array([0.0, 0.07, 0.19, 0.36, 0.5])
    f = lambda x: 1.0 + 2.0 * x + 3.0 * x * x
    off, w = O.formula("d1_fwd_nocond", xs[1:4])
    assert sum(wt * f(xs[1 + o]) for o, wt in zip(off, w)) == pytest.approx(2.0 + 6.0 * xs[1], rel=1e-12)
    off, w = O.formula("d1_bwd_nocond", xs[1:4])
    assert sum(wt * f(xs[3 + o]) for o, wt in zip(off, w)) == pytest.approx(2.0 + 6.0 * xs[3], rel=1e-12)
    off, w = O.formula("d2_fwd_dirichlet", xs[0:4])
    h1, h2, h3 = xs[1] - xs[0], xs[2] - xs[1], xs[3] - xs[1]
    bcw = 2.0 * (h2 + h3) / (h1 * (h1 + h2) * (h1 + h3))
    assert sum(wt * f(xs[1 + o]) for o, wt in zip(off, w)) + bcw * f(xs[0]) == pytest.approx(6.0, rel=1e-10)
