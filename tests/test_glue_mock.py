"""glue/nsb200.c COMPILED, LINKED AND RUN -- not against PETSc (absent from this image) but against the single-rank functional model
of the PETSc / Fluca API subset it uses (tests/petsc_stub/petsc_fluca_mock.c: DMStag's element-wise storage with partial elements and
ghost elements, VecNest state counters, the order of operations of NSSetUp / NSStep / NSViewSolution / NSLoadSolution).

tests/c/ns_b200_glue_driver.c is an application in the style of the reference's drivers (cavity_flow_2d.c, cavity_flow_3d.c,
taylor_green_vortex.c): -ns_type b200, boundary callbacks, initial condition written into ns->sol, NSStep, solution read back through
the DMStag API.  Here its output is compared with the oracle: K steps to 1e-10, NSFormFunction, a user edit of ns->sol between steps,
write / destroy / load / continue, lazy and staged downloads, the boundary-plane caches, the base class with and without the
matrix-free hooks of glue/patches/.  CPU: the library underneath is the host-emulation build; -m gpu: the CUDA library."""
import os
import subprocess

import numpy as np
import pytest

from oracle import oracle as O
from tests import cases, parity

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TIGHT = ["-ns_b200_outer_rtol=1e-13", "-ns_b200_momentum_rtol=1e-13", "-ns_b200_schur_rtol=1e-13"]
_exe = {}


def _build(tmp_path_factory, libpath, matrixfree=False):
    key = (libpath, matrixfree)
    if key not in _exe:
        exe = str(tmp_path_factory.mktemp("glue") / "ns_b200_glue_driver")
        d, f = os.path.split(libpath)
        cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
        src = [os.path.join(ROOT, "tests", "c", "ns_b200_glue_driver.c"), os.path.join(ROOT, "tests", "petsc_stub", "petsc_fluca_mock.c"), os.path.join(ROOT, "glue", "nsb200.c")]
        flags = ["-std=gnu11", "-O1", "-g", "-Wall", "-Wextra", "-Werror", "-Wno-unused-parameter"] + (["-DFLUCA_NS_HAS_MATRIXFREE"] if matrixfree else [])
        subprocess.run([cc] + flags + ["-I", os.path.join(ROOT, "tests", "petsc_stub"), "-I", os.path.join(ROOT, "include")] + src + ["-o", exe, "-L", d, f"-l:{f}", f"-Wl,-rpath,{d}", "-lm"], check=True)
        _exe[key] = exe
    return _exe[key]


@pytest.fixture(scope="module")
def exe(tmp_path_factory):
    parity.hostemu_library()
    return _build(tmp_path_factory, parity.HOSTEMU)


@pytest.fixture(scope="module")
def exe_matrixfree(tmp_path_factory):
    parity.hostemu_library()
    return _build(tmp_path_factory, parity.HOSTEMU, matrixfree=True)


CASES = {
    "cavity2d": lambda n, st: cases.cavity2d(n=n[0] if n else 16),
    "cavity3d": lambda n, st: cases.cavity3d(n=n or (8, 8, 4)),
    "cavity3d_full": lambda n, st: cases.cavity3d_full(n=n or (8, 8, 8)),
    "tgv": lambda n, st: cases.tgv(n=n[0] if n else 8),
    "tgv_periodic": lambda n, st: cases.tgv(n=n[0] if n else 8, periodic=True),
    "channel2d": lambda n, st, pout=0.0: cases.channel2d(n=n or (24, 12), pout=pout),
    "channel2d_t": lambda n, st, pout=0.0: cases.channel2d(n=n or (24, 12), pout=pout, time_dependent=True),
    "channel3d": lambda n, st, pout=0.0: cases.channel3d(n=n or (12, 8, 8), pout=pout),
    "channel3d_pz": lambda n, st, pout=0.0: cases.channel3d(n=n or (12, 8, 8), pout=pout, periodic_z=True),
}


def make_case(name, n=None, stretch=0.0, pout=0.0):
    c = CASES[name](n, stretch, pout) if name.startswith("channel") else CASES[name](n, stretch)
    c.stretch = stretch
    return c


def run(exe, name, tmp_path, n=None, stretch=0.0, pout=0.0, steps=2, steps2=1, scenario="plain", init="zero", opts=(), extra=()):
    out = str(tmp_path / f"{name}_{scenario}.bin")
    cmd = [exe, f"case={name}", f"out={out}", f"steps={steps}", f"steps2={steps2}", f"scenario={scenario}", f"init={init}", f"stretch={stretch}", f"pout={pout}"]
    if n:
        cmd.append("n=" + ",".join(str(x) for x in n))
    r = subprocess.run(cmd + list(extra) + list(opts), capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    raw = np.fromfile(out)
    dim, nstate, nf = int(raw[0]), int(raw[4]), int(raw[5])
    marks = raw[6 + nstate + nf :]
    raw = raw[: 6 + nstate + nf]
    case = make_case(name, n, stretch, pout)
    cell, face = case.shapes()
    assert dim == case.dim and tuple(int(x) for x in raw[1 : 1 + dim]) == case.n

    def split(x, with_phalf):
        o, res = 0, {}
        ncell = int(np.prod(cell))
        res["v"] = x[o : o + dim * ncell].reshape((dim,) + cell)
        o += dim * ncell
        res["U"] = []
        for d in range(dim):
            k = int(np.prod(face[d]))
            res["U"].append(x[o : o + k].reshape(face[d]))
            o += k
        res["p"] = x[o : o + ncell].reshape(cell)
        o += ncell
        if with_phalf:
            res["phalf"] = x[o : o + ncell].reshape(cell)
            o += ncell
        assert o == len(x)
        return res

    lines = r.stdout.splitlines()
    info = {ln.split()[0]: ln.split()[1:] for ln in lines if ln.split() and ln.split()[0] in ("STEP", "BCCALLS", "G2L", "LIVE")}
    assert info["LIVE"] == ["0"], "the glue leaked PETSc objects or memory: " + r.stdout  # NSDestroy_B200 frees everything it made
    return dict(case=case, state=split(raw[6 : 6 + nstate], True), f=split(raw[6 + nstate : 6 + nstate + nf], False) if nf else None, stdout=r.stdout, info=info, marks=marks)


def smooth_state(case):
    """The closed forms of f_smooth_* in tests/c/ns_b200_glue_driver.c on this case's cell centres and face centres."""
    xc, xf, per = case.centres(), case.faces(), case.periodic()

    def grids(coords):
        g = cases._mesh(coords, case.dim)
        return g if case.dim == 3 else g + [np.zeros_like(g[0])]

    X = grids(xc)
    cell, face = case.shapes()
    v = np.stack([(0.5 * np.sin(1.3 * X[0] + 0.7 * X[1] + 0.5 * X[2] + c)).reshape(cell) for c in range(case.dim)])
    p = np.cos(0.9 * X[0] - 1.1 * X[1] + 0.3 * X[2]).reshape(cell)
    U = []
    for d in range(case.dim):
        coords = list(xc)
        coords[d] = xf[d][:-1] if per[d] else xf[d]
        F = grids(coords)
        U.append((0.5 * np.sin(0.8 * F[0] + 1.2 * F[1] + 0.6 * F[2] + 2.0 + d)).reshape(face[d]))
    return v, U, p


def oracle_steps(case, steps, mode=0, ainv=(0, 0), state=None):
    orc = cases.make_oracle(case)
    orc.set_state(*(state if state is not None else case.initial_state()))
    opt = O.default_options(mode=mode, schur_ainv=ainv[0], upper_ainv=ainv[1], **parity.ORC_TIGHT)
    infos = [orc.step(opt) for _ in range(steps)]
    return orc, opt, infos


def assert_state(got, ref, tol=1e-10):
    err = dict(v=parity.rel(got["v"], ref["v"]), U=parity.relU(got["U"], ref["U"]), p=parity.rel(got["p"], ref["p"]), phalf=parity.rel(got["phalf"], ref["phalf"]))
    assert not any(np.isnan(x).any() for x in [got["v"], got["p"], got["phalf"]] + got["U"]), "an entry of a partial DMStag element was read"
    assert err["v"] <= tol and err["U"] <= tol and err["p"] <= 10 * tol and err["phalf"] <= 10 * tol, err
    return err


PLAIN = [
    ("cavity2d", None, 0.0, 0.0, 0, "zero"),
    ("cavity2d", (12, 12), 0.2, 0.0, 1, "smooth"),
    ("cavity3d", None, 0.0, 0.0, 0, "zero"),
    ("cavity3d_full", (6, 5, 4), 0.15, 0.0, 0, "smooth"),
    ("tgv", None, 0.0, 0.0, 0, "tgv"),
    ("tgv_periodic", None, 0.0, 0.0, 1, "tgv"),
    ("channel2d_t", None, 0.0, 0.3, 0, "smooth"),
    ("channel3d", (8, 6, 5), 0.1, 0.2, 0, "zero"),
    ("channel3d_pz", None, 0.0, 0.0, 0, "smooth"),
]


@pytest.mark.parametrize("name,n,stretch,pout,mode,init", PLAIN, ids=[f"{c[0]}{'_stretched' if c[2] else ''}_{'fractional' if c[4] else 'coupled'}" for c in PLAIN])
def test_steps_through_the_glue_equal_the_oracle(exe, tmp_path, name, n, stretch, pout, mode, init):
    """ns->sol after K steps of NS type b200 = the oracle's state: boundary callbacks evaluated by the glue at the right points and
    times (time-dependent inflow / outlet pressure, the analytic Taylor-Green walls), initial condition taken from the host Vec,
    DMStag <-> compact layout both ways, on uniform and stretched meshes, with periodic directions and partial elements."""
    res = run(exe, name, tmp_path, n=n, stretch=stretch, pout=pout, steps=3, init=init, opts=TIGHT + [f"-ns_b200_mode={mode}"])
    case = res["case"]
    state = smooth_state(case) if init == "smooth" else None
    orc, _, _ = oracle_steps(case, 3, mode=mode, state=state)
    assert_state(res["state"], orc.get_state())
    assert res["info"]["STEP"][0] == "3"


def test_formfunction_returns_b_and_leaves_the_run_alone(exe, tmp_path):
    """NSFormFunction (nsbasic.c:316-323) fills f with the right-hand side b of the coming step, as the reference's type does
    (cnlinearcart2d.c:2071-2171: the b(x) of SNESSetPicard); calling it between two steps changes nothing."""
    res = run(exe, "channel2d_t", tmp_path, pout=0.3, steps=2, steps2=1, scenario="formfunction", opts=TIGHT)
    case = res["case"]
    orc, opt, _ = oracle_steps(case, 2)
    b = orc.prepare_step(opt)
    bv, bU, bp = orc.split(b)
    f = res["f"]
    assert parity.rel(f["v"], bv) <= 1e-11 and parity.relU(f["U"], bU) <= 1e-11 and np.abs(f["p"] - bp).max() <= 1e-10 * max(1.0, np.abs(bv).max())
    orc2, _, _ = oracle_steps(case, 3)
    assert_state(res["state"], orc2.get_state())


@pytest.mark.parametrize("nestbump", [1, 0])
def test_user_edit_of_the_host_solution_is_uploaded(exe, tmp_path, nestbump):
    """A user writes ns->sol between two steps (SURVEY 8b "host/device state coherence"): the type notices through the object state of
    the VecNest, whether or not restoring an untouched sub-vector bumps that state."""
    res = run(exe, "cavity2d", tmp_path, n=(12, 12), steps=2, steps2=2, scenario="edit", opts=TIGHT, extra=[f"nestbump={nestbump}"])
    case = res["case"]
    orc, opt, _ = oracle_steps(case, 2)
    st = orc.get_state()
    v, U, p = smooth_state(case)
    orc.set_state(v, U, p, phalf=st["phalf"], step=st["step"], t=st["t"])
    for _ in range(2):
        orc.step(opt)
    assert_state(res["state"], orc.get_state())


def test_write_destroy_load_continue(exe, tmp_path):
    """NSViewSolution -> NSDestroy -> new NS -> NSLoadSolution -> NSStep: p-half round-trips under the name "PressureHalfStep"
    (cnlinear.c:54,146-162), step and time come back, and the continued run equals the uninterrupted one."""
    res = run(exe, "channel3d", tmp_path, n=(8, 6, 5), pout=0.2, steps=2, steps2=2, scenario="restart", opts=TIGHT)
    orc, _, _ = oracle_steps(res["case"], 4)
    assert_state(res["state"], orc.get_state())
    assert res["info"]["STEP"][0] == "4"


def test_lazy_download_refreshes_for_the_first_observer(exe, tmp_path):
    """-ns_b200_sync_interval 0: nothing is downloaded during the time loop, the first NSViewSolution refreshes ns->sol."""
    res = run(exe, "cavity3d", tmp_path, steps=3, opts=TIGHT + ["-ns_b200_sync_interval=0"])
    orc, _, _ = oracle_steps(res["case"], 3)
    assert_state(res["state"], orc.get_state())
    res1 = run(exe, "cavity3d", tmp_path, steps=3, opts=TIGHT)
    assert int(res["info"]["G2L"][0]) == int(res1["info"]["G2L"][0])  # uploads: one set at the first step in both runs


def test_staged_solution_is_the_state_at_the_staging_point(exe, tmp_path):
    """NSB200StageSolution before the last step, NSB200SyncSolution after it: ns->sol holds the state of the staging point."""
    res = run(exe, "cavity2d", tmp_path, steps=3, scenario="stage", opts=TIGHT + ["-ns_b200_sync_interval=0"])
    orc, _, _ = oracle_steps(res["case"], 2)
    assert_state(res["state"], orc.get_state())


def test_boundary_planes_are_cached(exe, tmp_path):
    """-ns_b200_bc_time_independent evaluates the callbacks once; the answer does not change (constant lid)."""
    a = run(exe, "cavity2d", tmp_path, steps=3, opts=TIGHT)
    b = run(exe, "cavity2d", tmp_path, steps=3, opts=TIGHT + ["-ns_b200_bc_time_independent"])
    n = 16
    assert int(a["info"]["BCCALLS"][0]) == 3 * 4 * 2 * n and int(b["info"]["BCCALLS"][0]) == 4 * 2 * n
    for k in ("v", "p", "phalf"):
        assert np.array_equal(a["state"][k], b["state"][k])


def test_options_monitor_and_abf_variants(exe, tmp_path):
    """-ns_pc_abf_*_ainv_type reach the device-side factors; -ns_ksp_monitor prints PETSc's residual lines; NSView reports both."""
    res = run(exe, "cavity2d", tmp_path, n=(12, 12), steps=2, opts=TIGHT + ["-ns_pc_abf_schur_ainv_type=DIAG", "-ns_pc_abf_upper_ainv_type=rowsum", "-ns_ksp_monitor"])
    orc, _, infos = oracle_steps(res["case"], 2, ainv=(1, 2))
    assert_state(res["state"], orc.get_state())
    assert "Schur complement A inverse type DIAG, upper triangular A inverse type ROWSUM" in res["stdout"]
    mon = [ln for ln in res["stdout"].splitlines() if "KSP Residual norm" in ln]
    assert len(mon) == sum(i.nhist for i in infos)
    first = [float(ln.split()[-1]) for ln in mon[: infos[0].nhist]]
    for a, b in zip(first, [infos[0].hist[i] for i in range(infos[0].nhist)]):
        assert a == pytest.approx(b, rel=1e-6, abs=1e-12 * infos[0].hist[0])


def test_base_class_with_the_matrix_free_hooks(exe_matrixfree, tmp_path):
    """The same glue against a base class carrying glue/patches/0001 (no J, no sol -> sol0 copy): same answer."""
    res = run(exe_matrixfree, "channel3d_pz", tmp_path, steps=2, init="smooth", opts=TIGHT)
    orc, _, _ = oracle_steps(res["case"], 2, state=smooth_state(res["case"]))
    assert_state(res["state"], orc.get_state())


def test_immersed_boundary_through_the_glue(exe, tmp_path):
    """NSB200SetMarkers / NSB200GetMarkerForces (the b200 type's own entry points: the reference has no IBM): a sphere of markers in
    the inflow / outlet channel, two coupled steps; fields and marker forces against the oracle's definition of the coupling."""
    nm = 150
    res = run(exe, "channel3d", tmp_path, n=(12, 8, 8), pout=0.1, steps=2, init="smooth", opts=TIGHT, extra=[f"markers={nm}"])
    case = res["case"]
    mk = cases.sphere_markers((0.1, 0.0, 0.05), 1.2, nm, 4.0 / case.n[1])
    orc = cases.make_oracle(case)
    orc.set_state(*smooth_state(case))
    orc.set_markers(mk["X"], mk["Ud"], mk["dV"], 4)
    opt = O.default_options(mode=0, **parity.ORC_TIGHT)
    for _ in range(2):
        orc.step(opt)
    assert_state(res["state"], orc.get_state())
    F, Um = res["marks"][: 3 * nm].reshape(3, nm), res["marks"][3 * nm :].reshape(3, nm)
    Fo, Uo = orc.marker_forces()
    assert parity.rel(Um, Uo) <= 1e-10 and parity.rel(F, Fo) <= 1e-8


def test_the_3d_quirks_of_the_reference_are_options(exe, tmp_path):
    """-ns_b200_no_t_outlet_quirk / -ns_b200_no_bcg_quirk select the 2-D file's form of the two places where cnlinearcart3d.c differs
    (operator T at an upper outlet, scaling of the outlet-gradient BC vector); the default is what the reference's 3-D file does."""
    res = run(exe, "channel3d", tmp_path, n=(8, 6, 5), pout=0.3, steps=2, init="smooth", opts=TIGHT + ["-ns_b200_no_t_outlet_quirk", "-ns_b200_no_bcg_quirk"])
    case = res["case"]
    try:
        O.set_t_outlet_quirk(False)
        orc = cases.make_oracle(case)
    finally:
        O.set_t_outlet_quirk(True)
    orc.set_state(*smooth_state(case))
    opt = O.default_options(mode=0, quirk_bcg_scale=0, **parity.ORC_TIGHT)
    for _ in range(2):
        orc.step(opt)
    assert_state(res["state"], orc.get_state())
    dflt = run(exe, "channel3d", tmp_path, n=(8, 6, 5), pout=0.3, steps=2, init="smooth", opts=TIGHT)
    assert parity.rel(dflt["state"]["v"], res["state"]["v"]) > 1e-4  # the quirks do change the answer


def test_unknown_option_value_is_an_error(exe, tmp_path):
    r = subprocess.run([exe, "case=cavity2d", f"out={tmp_path / 'x.bin'}", "-ns_pc_abf_schur_ainv_type=nonsense"], capture_output=True, text=True)
    assert r.returncode != 0 and "unknown value" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("name,n,mode", [("cavity3d", (32, 16, 8), 0), ("channel3d_pz", (32, 8, 8), 0), ("tgv", (16, 16), 1)], ids=["cavity3d_tiles", "channel3d_periodic_z", "tgv_fractional"])
def test_steps_through_the_glue_cuda(tmp_path_factory, tmp_path, name, n, mode):
    import fluca_b200 as fb

    assert fb._lib.load().fluca_b200_is_host_emulation() == 0
    exe = _build(tmp_path_factory, fb._lib.PRODUCT_LIB)
    init = "tgv" if name == "tgv" else "smooth"
    res = run(exe, name, tmp_path, n=n, steps=2, init=init, opts=TIGHT + [f"-ns_b200_mode={mode}"])
    case = res["case"]
    orc, _, _ = oracle_steps(case, 2, mode=mode, state=None if name == "tgv" else smooth_state(case))
    assert_state(res["state"], orc.get_state())
