"""Generates tests/golden/ns_steps.npz: small Navier-Stokes step fixtures produced by the CPU oracle (oracle/), inputs included.

The reference stores no NS golden output and cannot be built in this image, so -- as SURVEY.md 8c
prescribes -- the oracle itself produces the fixtures: seeded inputs, exact (tight-tolerance, mode A = coupled solve and mode
B = one ABF application) outputs after two steps, plus one operator-level tier (right-hand side of the first step).  They
pin the oracle against accidental change and give the product tests a comparison that needs no oracle at run time.
PARITY STATUS: these are the ORACLE's outputs (incl. the IBM case, which has no reference); the fixtures computed by the reference's
own compiled sources are tests/golden/ns_reference.npz (make_ns_reference_golden.py).  See oracle/fluca_oracle.h.

    python tests/golden/make_ns_golden.py          # rewrites tests/golden/ns_steps.npz
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import oracle as O  # noqa: E402
from tests import cases  # noqa: E402

TIGHT = dict(outer_rtol=1e-13, mom_rtol=1e-13, schur_rtol=1e-13)


def fixtures():
    """name -> (case, seed, markers or None, (Schur, upper) PCABFAinvType); kept tiny: the file is a few hundred kB"""
    c3 = cases.channel3d(n=(12, 8, 8), pout=0.1, dt=0.05)
    ID, DIAG, ROWSUM = 0, 1, 2  # PCABFAinvType, flucans.h:99-103
    return {
        "cavity2d_8": (cases.cavity2d(n=8), 3, None, (ID, ID)),
        "cavity3d_sym_6x6x4": (cases.cavity3d(n=(6, 6, 4)), 5, None, (ID, ID)),
        "channel3d_outlet_8x6x6": (cases.channel3d(n=(8, 6, 6), pout=0.2, dt=0.05), 11, None, (ID, ID)),
        "tgv_periodic_8": (cases.tgv(n=8, periodic=True, dt=0.05), None, None, (ID, ID)),
        "sphere_ibm_12x8x8": (c3, 31, cases.sphere_markers((0.1, 0.0, 0.05), 1.2, 60, 0.5), (ID, ID)),
        # PCABF variants (abfpc.c:81-94, 151-168)
        "cavity2d_8_abf_diag": (cases.cavity2d(n=8), 3, None, (DIAG, DIAG)),
        "cavity3d_sym_6x6x4_abf_rowsum_diag": (cases.cavity3d(n=(6, 6, 4)), None, None, (ROWSUM, DIAG)),
        "channel3d_outlet_8x6x6_abf_diag": (cases.channel3d(n=(8, 6, 6), pout=0.2, dt=0.05), 11, None, (DIAG, DIAG)),
    }


def run(case, seed, markers, mode, nsteps=2, ainv=(0, 0)):
    orc = cases.make_oracle(case)
    state = case.initial_state(seed=seed)
    orc.set_state(*state)
    if markers is not None:
        orc.set_markers(markers["X"], markers["Ud"], markers["dV"], markers.get("npts", 4))
    rhs = orc.prepare_step().copy()
    orc.set_state(*state)  # prepare_step does not advance, but keep the two uses independent
    its = []
    for _ in range(nsteps):
        info = orc.step(O.default_options(mode=mode, schur_ainv=ainv[0], upper_ainv=ainv[1], **TIGHT))
        its.append(info.outer_its)
    out = orc.get_state()
    return state, rhs, out, its


def main():
    data = {}
    for name, (case, seed, markers, ainv) in fixtures().items():
        for mode, tag in ((0, "coupled"), (1, "fractional")):
            state, rhs, out, its = run(case, seed, markers, mode, ainv=ainv)
            k = f"{name}/{tag}"
            if tag == "coupled":
                data[f"{name}/in_v"], data[f"{name}/in_p"] = state[0], state[2]
                for d, u in enumerate(state[1]):
                    data[f"{name}/in_U{d}"] = u
                data[f"{name}/rhs"] = rhs
            data[f"{k}/v"], data[f"{k}/p"], data[f"{k}/phalf"] = out["v"], out["p"], out["phalf"]
            for d, u in enumerate(out["U"]):
                data[f"{k}/U{d}"] = u
            data[f"{k}/outer_its"] = np.array(its)
    path = os.path.join(ROOT, "tests", "golden", "ns_steps.npz")
    np.savez_compressed(path, **data)
    print(path, os.path.getsize(path), "bytes,", len(data), "arrays")


if __name__ == "__main__":
    main()
