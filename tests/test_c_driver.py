"""The C ABI driven from plain C (tests/c/cavity_flow_2d_b200.c: the case of the reference's cavity_flow_2d.c, no Python and
no PETSc on the product side) against the oracle.  CPU: linked with the host-emulation test double; GPU: with the CUDA library."""
import os
import subprocess

import numpy as np
import pytest

from oracle import oracle as O
from tests import cases, parity

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "c", "cavity_flow_2d_b200.c")


def _build(tmp_path, libpath):
    exe = str(tmp_path / "cavity_flow_2d_b200")
    d, f = os.path.split(libpath)
    cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    subprocess.run([cc, "-std=c99", "-O1", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), SRC, "-o", exe, "-L", d, f"-l:{f}", f"-Wl,-rpath,{d}", "-lm"], check=True)
    return exe


def _run_and_compare(exe, n, steps, mode, ainv):
    r = subprocess.run([exe, str(n), "100", str(steps), str(mode), str(ainv[0]), str(ainv[1]), "1e-13"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    lines = r.stdout.strip().splitlines()
    res = [ln for ln in lines if ln.startswith("RESULT")][0].split()
    got = np.array([float(x) for x in res[3:]])
    monitor = [ln for ln in lines if " NS dt " in ln]
    assert len(monitor) == steps
    case = cases.cavity2d(n=n, Re=100.0)
    orc = cases.make_oracle(case)
    orc.set_state(*case.initial_state())
    opt = O.default_options(mode=mode, schur_ainv=ainv[0], upper_ainv=ainv[1], **parity.ORC_TIGHT)
    infos = [orc.step(opt) for _ in range(steps)]
    st = orc.get_state()
    ref = np.concatenate([[st["v"][0].sum(), st["v"][1].sum(), st["p"].sum()], st["v"][0][0, :, n // 2]])
    assert np.abs(got[3:] - ref[3:]).max() <= 1e-10 * np.abs(ref[3:]).max()  # u on the vertical centre line
    assert abs(got[0] - ref[0]) <= 1e-9 * np.abs(st["v"][0]).sum() and abs(got[1] - ref[1]) <= 1e-9 * np.abs(st["v"][1]).sum()
    assert abs(got[2] - ref[2]) <= 1e-8 * np.abs(st["p"]).sum()
    if mode == 0:  # the monitor lines carry the outer iteration counts of the oracle
        assert [int(ln.split("outer")[1].split()[0]) for ln in monitor] == [i.outer_its for i in infos]
    maxdiv = float([ln for ln in lines if ln.startswith("MAXDIV")][0].split()[1])
    assert maxdiv < 1e-9


@pytest.mark.parametrize("mode,ainv", [(0, (0, 0)), (1, (0, 0)), (0, (1, 1))], ids=["coupled", "fractional", "coupled_abf_diag"])
def test_c_driver_host_emulation(tmp_path, mode, ainv):
    parity.hostemu_library()
    _run_and_compare(_build(tmp_path, parity.HOSTEMU), 16, 3, mode, ainv)


@pytest.mark.gpu
@pytest.mark.parametrize("mode,ainv", [(0, (0, 0)), (1, (2, 2))], ids=["coupled", "fractional_abf_rowsum"])
def test_c_driver_cuda(tmp_path, mode, ainv):
    import fluca_b200 as fb

    assert fb._lib.load().fluca_b200_is_host_emulation() == 0
    _run_and_compare(_build(tmp_path, fb._lib.PRODUCT_LIB), 32, 3, mode, ainv)
