"""Seeded test cases shared by the oracle tests and the GPU parity tests: the product's workload definitions
(fluca_b200/workloads.py) plus the factories of the CPU oracle for them."""
from __future__ import annotations

from fluca_b200.workloads import *  # noqa: F401,F403
from fluca_b200.workloads import _const, _constp, _mesh, _smooth  # noqa: F401


def make_oracle(case):
    from oracle import oracle as O

    bcs = [O.BC(b["type"], velocity=b["velocity"], pressure=b["pressure"]) for b in case.bcs]
    return O.Oracle(case.n, case.faces(), case.rho, case.mu, case.dt, bcs)


def make_oracle_fast(case):
    """Same, but constant boundary data goes through the oracle's built-in C callbacks (timing runs)."""
    from oracle import oracle as O

    bcs = []
    for b in case.bcs:
        cv = getattr(b["velocity"], "const", None) if b["velocity"] is not None else None
        cp = getattr(b["pressure"], "constp", None) if b["pressure"] is not None else None
        bcs.append(O.BC(b["type"], velocity=b["velocity"], pressure=b["pressure"], const_velocity=cv, const_pressure=cp))
    return O.Oracle(case.n, case.faces(), case.rho, case.mu, case.dt, bcs)
