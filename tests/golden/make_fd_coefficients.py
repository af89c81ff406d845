"""Extract the stencil coefficients that the reference's own fd golden outputs pin.

Run in the build container (needs /root/reference); writes tests/golden/fd_coefficients.json.
The .out files are the byte-exact expected outputs of the reference's registered ctest cases
(fluca/tests/fd/CMakeLists.txt:15-25, runner fluca/cmake/RunTest.cmake:31-59).  SURVEY.md section 4
lists which of them cross-pin the hand-derived NS coefficients of fluca/src/ns/utils/cartdiscret.c.
All cases use a uniform 8-cell grid on [0,1] (h = 1/8), see fluca/tests/fd/ex1.c.
"""
import json
import os
import re

REF = "/root/reference/fluca/tests/fd/output"
CASES = [
    "ex1_second_deriv",
    "ex1_second_deriv_left_bc_dirichlet",
    "ex3_second_deriv_right_bc_dirichlet_scale_const",
    "ex2_all_second_deriv_up_bc_neumann",
    "ex1_first_deriv",
    "ex1_first_deriv_input_loc_elem_output_loc_left",
    "ex2_all_first_deriv_input_loc_face_output_loc_elem",
    "ex1_second_deriv_left_bc_none",
    "ex1_second_deriv_right_bc_neumann",
    "ex1_second_deriv_right_bc_none",
    "ex1_first_deriv_input_loc_elem_output_loc_left_left_bc_periodic",
    "ex2_all_second_deriv",
    "ex2_all_first_deriv",
]

LINE = re.compile(r"col\[(\d+)\]:\s*(.*?),\s*loc=(\w+),\s*c=([\w-]+),\s*v=([-+0-9.eE]+)")
IDX = re.compile(r"([ijk])=(-?\d+)")


def parse(path):
    out = {"header": None, "cols": []}
    with open(path) as f:
        for ln in f:
            ln = ln.rstrip("\n")
            m = LINE.search(ln)
            if m:
                idx = {k: int(v) for k, v in IDX.findall(m.group(2))}
                out["cols"].append({"index": idx, "loc": m.group(3), "c": m.group(4), "v": float(m.group(5))})
            elif out["header"] is None and ln.strip():
                out["header"] = ln.strip()
    return out


def main():
    res = {"_source": "thecasterian/fluca fluca/tests/fd/output/*.out", "_grid": {"cells": 8, "lo": 0.0, "hi": 1.0}}
    for c in CASES:
        p = os.path.join(REF, c + ".out")
        if os.path.exists(p):
            res[c] = parse(p)
    here = os.path.dirname(os.path.abspath(__file__))
    with open(os.path.join(here, "fd_coefficients.json"), "w") as f:
        json.dump(res, f, indent=1, sort_keys=True)
    print("wrote", len(res) - 2, "cases")


if __name__ == "__main__":
    main()
