"""NS type "b200" INSIDE THE REFERENCE'S OWN PROGRAMS.

oracle/_ref/ holds the reference as a program (`make -C oracle ref_app`, where /root/reference is present): its sys, mesh and NS
libraries -- flucainit.c, meshbasic.c, cart.c, nsbasic.c, nsopts.c, nssol.c, nsreg.c, cnlinear*.c, abfpc.c, ... everything but the
CGNS viewer and FlucaFD -- and its own NS test drivers cavity_flow_2d.c, cavity_flow_3d.c and taylor_green_vortex.c, all UNMODIFIED
and compiled from where they lie, on the single-rank PETSc model of oracle/ref_model/.  glue/nsb200.c is built against the
reference's own headers as the out-of-tree plugin of INTEGRATION.md 2(b) and loaded the way PETSc loads it:

    cavity_flow_2d -dll_append libfluca_nsb200.so -ns_type b200 ...

PetscInitialize opens the library and calls PetscDLLibraryRegister_fluca_nsb200 -> NSRegister("b200", NSCreate_B200)
(nsreg.c:5-11); NSSetFromOptions (nsopts.c:179-181) switches the type the driver had hard-coded; NSSetUp / NSSolve / NSStep /
NSMonitorSolution -> NSViewSolution of the reference's base class drive it.  The solution every step writes through the viewer
(Velocity, FaceNormalVelocity, Pressure and the type's PressureHalfStep, in DMStag's own ordering) is compared with the same
program run with the reference's own type cnlinear.  CPU: the plugin links the host-emulation build; -m gpu: the CUDA library.
What is modelled is PETSc only (exact linear solves in place of GMRES + ILU); no reference source is changed or copied."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFDIR = os.path.join(ROOT, "oracle", "_ref")
TIGHT = ["-ns_b200_outer_rtol", "1e-13", "-ns_b200_momentum_rtol", "1e-13", "-ns_b200_schur_rtol", "1e-13"]


def _ready(kind):
    if os.path.isdir("/root/reference/fluca"):
        subprocess.run(["make", "-C", os.path.join(ROOT, "tests", "hostemu")], check=True, stdout=subprocess.DEVNULL)
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "ref_app"], check=True, stdout=subprocess.DEVNULL)
    need = [os.path.join(REFDIR, p) for p in ("cavity_flow_2d", "cavity_flow_3d", "taylor_green_vortex", "libfluca_ref_full.so", os.path.join(kind, "libfluca_nsb200.so"))]
    return all(os.path.exists(p) for p in need)


def load_dump(path):
    """name -> values, as VecView of the model's stand-in for the CGNS viewer writes them (petsc_model_app.c)"""
    out, b, i = {}, open(path, "rb").read(), 0
    while i < len(b):
        j = b.index(b"\n", i)
        k = b.index(b"\n", j + 1)
        if b[i:j] == b"@sequence":  # step and time of the mesh when the file was written
            i = k + 1
            continue
        n = int(b[j + 1 : k])
        out[b[i:j].decode()] = np.frombuffer(b[k + 1 : k + 1 + 8 * n], dtype=np.float64).copy()
        i = k + 1 + 8 * n + 1
    return out


def run(prog, args, tmp_path, tag, plugin=None):
    dump = str(tmp_path / f"{prog}_{tag}.bin")
    cmd = [os.path.join(REFDIR, prog)] + list(args) + ["-ns_monitor", "-ns_monitor_solution", f"flucacgns:{dump}"]
    if plugin:
        cmd += ["-dll_append", os.path.join(REFDIR, plugin, "libfluca_nsb200.so"), "-ns_type", "b200"] + TIGHT
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "WARNING! There are options you set that were not used" not in r.stderr, r.stderr
    return load_dump(dump), r.stdout


def compare(prog, args, tmp_path, plugin, ref_extra=(), b200_extra=(), tol=1e-10):
    a, out_a = run(prog, list(args) + list(ref_extra), tmp_path, "cnlinear")
    b, out_b = run(prog, list(args) + list(b200_extra), tmp_path, "b200", plugin=plugin)
    assert set(a) == set(b) == {"Velocity", "FaceNormalVelocity", "Pressure", "PressureHalfStep"}
    for k in a:
        scale = max(np.abs(a[k]).max(), 1e-300)
        assert a[k].shape == b[k].shape and np.abs(a[k] - b[k]).max() <= (10 * tol if k.startswith("Pressure") else tol) * scale, (prog, k, np.abs(a[k] - b[k]).max() / scale)
    mon = [ln for ln in out_a.splitlines() if " NS dt " in ln]
    assert mon and mon == [ln for ln in out_b.splitlines() if " NS dt " in ln]  # NSMonitorDefault: same steps, same times
    return out_a, out_b


CASES = [
    ("cavity_flow_2d", ["-cart_grid_x", "12", "-cart_grid_y", "12", "-ns_time_step_size", "0.04", "-ns_max_steps", "4"], (), ()),
    ("cavity_flow_2d", ["-cart_grid_x", "10", "-cart_grid_y", "8", "-Re", "400", "-ns_time_step_size", "0.05", "-ns_max_steps", "3"], ("-ns_ksp_type", "preonly"), ("-ns_b200_mode", "1")),
    ("cavity_flow_2d", ["-cart_grid_x", "8", "-cart_grid_y", "8", "-ns_time_step_size", "0.06", "-ns_max_steps", "3", "-ns_pc_abf_schur_ainv_type", "DIAG", "-ns_pc_abf_upper_ainv_type", "ROWSUM"], ("-ns_ksp_type", "preonly"), ("-ns_b200_mode", "1")),
    ("cavity_flow_3d", ["-cart_grid_x", "6", "-cart_grid_y", "6", "-cart_grid_z", "4", "-ns_time_step_size", "0.08", "-ns_max_steps", "3"], (), ()),
    ("taylor_green_vortex", ["-nsteps", "5", "-t_final", "0.25"], (), ()),
    ("taylor_green_vortex", ["-nsteps", "4", "-t_final", "0.2", "-periodic", "-cart_grid_x", "12", "-cart_grid_y", "12"], (), ()),
]
IDS = ["cavity2d_12", "cavity2d_10x8_preonly_vs_fractional", "cavity2d_abf_diag_rowsum", "cavity3d_6x6x4", "tgv_dirichlet", "tgv_periodic_12"]


@pytest.mark.skipif(not _ready("hostemu"), reason="oracle/_ref (the reference's programs on the PETSc model) is not built here")
@pytest.mark.parametrize("prog,args,ref_extra,b200_extra", CASES, ids=IDS)
def test_b200_inside_the_references_own_programs(tmp_path, prog, args, ref_extra, b200_extra):
    out_a, out_b = compare(prog, args, tmp_path, "hostemu", ref_extra, b200_extra)
    if prog == "taylor_green_vortex":  # the driver's own verdict: the error against the analytic solution, printed with %g
        err_a = [float(ln.split(":")[1]) for ln in out_a.splitlines() if "error" in ln]
        err_b = [float(ln.split(":")[1]) for ln in out_b.splitlines() if "error" in ln]
        assert len(err_a) == 2 and err_a == pytest.approx(err_b, rel=1e-5)


@pytest.mark.skipif(not _ready("hostemu"), reason="oracle/_ref is not built here")
def test_the_type_is_unknown_without_the_plugin(tmp_path):
    """-ns_type b200 without -dll_append: the reference's NSSetType does not know the type (nsbasic.c:55-79) -- nothing of it is linked
    into the reference's programs."""
    r = subprocess.run([os.path.join(REFDIR, "cavity_flow_2d"), "-ns_type", "b200", "-cart_grid_x", "8", "-cart_grid_y", "8", "-ns_max_steps", "1"], capture_output=True, text=True)
    assert r.returncode != 0 and "b200" in r.stderr


@pytest.mark.gpu
@pytest.mark.skipif(not _ready("cuda"), reason="oracle/_ref is not built here")
@pytest.mark.parametrize(
    "prog,args,ref_extra,b200_extra",
    [
        # 32 x 8 cells per plane reach the TMA tile kernels; the model's inner solves are dense, hence few planes and one ABF application
        ("cavity_flow_3d", ["-cart_grid_x", "32", "-cart_grid_y", "8", "-cart_grid_z", "3", "-ns_time_step_size", "0.02", "-ns_max_steps", "2"], ("-ns_ksp_type", "preonly"), ("-ns_b200_mode", "1")),
        ("taylor_green_vortex", ["-nsteps", "3", "-t_final", "0.15", "-cart_grid_x", "16", "-cart_grid_y", "16"], (), ()),
    ],
    ids=["cavity3d_tiles_fractional", "tgv_16_coupled"],
)
def test_b200_cuda_inside_the_references_own_programs(tmp_path, prog, args, ref_extra, b200_extra):
    compare(prog, args, tmp_path, "cuda", ref_extra, b200_extra)


@pytest.mark.skipif(not os.path.isdir("/root/reference/fluca"), reason="the reference tree is not on this machine")
def test_matrix_free_patch_compiled_into_the_references_base_class(tmp_path):
    """SURVEY 8f rank 2, as a running program: glue/patches/0001-ns-matrix-free-type-hooks.patch applied to a COPY of the reference's
    sources (under the test's temporary directory), the patched reference built on the PETSc model, the glue built with
    -DFLUCA_NS_HAS_MATRIXFREE.  The patched NSSetUp builds no Jacobian / null space / SNES for b200 and NSStep skips the host copy
    sol -> sol0; the answer is the unpatched reference's cnlinear answer, and cnlinear itself is unchanged by the patch."""
    import shutil

    tree = tmp_path / "patched"
    (tree / "fluca").mkdir(parents=True)
    for sub in ("include", "src", "tests", "app"):
        shutil.copytree(os.path.join("/root/reference/fluca", sub), tree / "fluca" / sub)
    r = subprocess.run(["patch", "-p1", "-i", os.path.join(ROOT, "glue", "patches", "0001-ns-matrix-free-type-hooks.patch")], cwd=tree, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    out = tmp_path / "out"
    subprocess.run(["make", "-C", os.path.join(ROOT, "tests", "hostemu")], check=True, stdout=subprocess.DEVNULL)
    r = subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "ref_app", f"REF={tree}/fluca", f"OUTDIR={out}", "PLUGINDEFS=-DFLUCA_NS_HAS_MATRIXFREE", f"PLUGINS={out}/hostemu/libfluca_nsb200.so"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    args = ["-cart_grid_x", "6", "-cart_grid_y", "6", "-cart_grid_z", "4", "-ns_time_step_size", "0.08", "-ns_max_steps", "3", "-ns_monitor"]

    trace = {}

    def go(exe_dir, extra, tag):
        dump = str(tmp_path / f"{tag}.bin")
        rr = subprocess.run([os.path.join(exe_dir, "cavity_flow_3d")] + args + ["-ns_monitor_solution", f"flucacgns:{dump}"] + extra, capture_output=True, text=True, timeout=600, env=dict(os.environ, PETSC_MODEL_TRACE="1"))
        assert rr.returncode == 0, rr.stdout + rr.stderr
        trace[tag] = rr.stderr
        return load_dump(dump)

    ref = go(REFDIR, [], "unpatched_cnlinear")
    cn = go(str(out), [], "patched_cnlinear")
    b2 = go(str(out), ["-dll_append", str(out / "hostemu" / "libfluca_nsb200.so"), "-ns_type", "b200"] + TIGHT, "patched_b200")
    for k in ref:
        assert np.array_equal(ref[k], cn[k]), k  # the patch does not touch a type that leaves the flag unset
        scale = np.abs(ref[k]).max()
        assert np.abs(ref[k] - b2[k]).max() <= (1e-9 if k.startswith("Pressure") else 1e-10) * scale, k
    # what the patch is for: no MatNest Jacobian (nsbasic.c:203-208) and no host copy sol -> sol0 per step (nsbasic.c:281-282) for b200
    assert trace["patched_cnlinear"].count("MatCreateNest") == 1 and trace["patched_cnlinear"].count("VecCopy of a nest") >= 3  # one per step (+ the model's solver)
    assert "MatCreateNest" not in trace["patched_b200"] and "VecCopy of a nest" not in trace["patched_b200"]


@pytest.mark.skipif(not _ready("hostemu"), reason="oracle/_ref is not built here")
def test_b200_at_its_default_tolerances_lands_on_the_references_converged_answer(tmp_path):
    """cavity_flow_3d as a user would run it: -ns_type b200 with NO tolerance options (the reference's defaults, 1e-5, with the inner
    solves relaxed by the inexact-Krylov rule) against the reference's cnlinear with every linear solve taken to convergence."""
    args = ["-cart_grid_x", "8", "-cart_grid_y", "8", "-cart_grid_z", "5", "-ns_time_step_size", "0.06", "-ns_max_steps", "4", "-ns_monitor"]
    dumps = {}
    for tag, extra in (("cnlinear", []), ("b200", ["-dll_append", os.path.join(REFDIR, "hostemu", "libfluca_nsb200.so"), "-ns_type", "b200"])):
        dump = str(tmp_path / f"{tag}.bin")
        r = subprocess.run([os.path.join(REFDIR, "cavity_flow_3d")] + args + ["-ns_monitor_solution", f"flucacgns:{dump}"] + extra, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout + r.stderr
        dumps[tag] = load_dump(dump)
    for k in ("Velocity", "FaceNormalVelocity"):
        a, b = dumps["cnlinear"][k], dumps["b200"][k]
        assert np.linalg.norm(a - b) <= 1e-4 * np.linalg.norm(a), (k, np.linalg.norm(a - b) / np.linalg.norm(a))
    a, b = dumps["cnlinear"]["Pressure"], dumps["b200"]["Pressure"]
    assert np.linalg.norm(a - b) <= 2e-3 * np.linalg.norm(a), np.linalg.norm(a - b) / np.linalg.norm(a)


@pytest.mark.skipif(not (_ready("hostemu") and os.path.exists(os.path.join(REFDIR, "ref_restart_app"))), reason="oracle/_ref is not built here")
@pytest.mark.parametrize("writer,reader", [("cnlinear", "b200"), ("b200", "cnlinear"), ("b200", "b200")])
def test_restart_files_are_interchangeable_between_cnlinear_and_b200(tmp_path, writer, reader):
    """tests/c/ref_restart_app.c (written against the reference's public API, linked with the reference's own libraries): two steps
    with one NS type, NSViewSolution; a second process with the other type, NSLoadSolution, two more steps.  The reference's
    NSLoadSolution restores step and time from the file and the type's extra state travels as "PressureHalfStep"
    (cnlinear.c:54,146-162; NSViewSolution_B200 / NSLoadSolution_B200), so the continued run equals four uninterrupted cnlinear steps."""
    exe = os.path.join(REFDIR, "ref_restart_app")
    plug = ["-dll_append", os.path.join(REFDIR, "hostemu", "libfluca_nsb200.so"), "-ns_type", "b200"] + TIGHT
    common = ["-ns_time_step_size", "0.06"]

    def go(mode, path, steps, kind):
        r = subprocess.run([exe, mode, path] + common + ["-ns_max_steps", str(steps)] + (plug if kind == "b200" else []), capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout + r.stderr
        return r.stdout

    full, half = str(tmp_path / "full.bin"), str(tmp_path / "half.bin")
    go("write", full, 4, "cnlinear")
    go("write", half, 2, writer)
    out = go("read", half, 4, reader)
    assert "finished at step 4 time 0.24" in out  # step and time came back from the file (nssol.c:200-202)
    a, b = load_dump(full), load_dump(half + ".out")
    for k in ("Velocity", "FaceNormalVelocity", "Pressure", "PressureHalfStep"):
        assert np.abs(a[k] - b[k]).max() <= (1e-9 if k.startswith("Pressure") else 1e-10) * np.abs(a[k]).max(), (k, writer, reader)


@pytest.mark.skipif(not (_ready("hostemu") and os.path.exists(os.path.join(REFDIR, "ref_restart_app"))), reason="oracle/_ref is not built here")
def test_non_uniform_mesh_through_the_references_coordinate_api(tmp_path):
    """A stretched mesh set through MeshCartGetCoordinateArrays / MeshCartRestoreCoordinateArrays (cart.c:467-502) of the reference's
    own Mesh implementation: the glue reads the face coordinates back through MeshCartGetCoordinateArraysRead and b200 reproduces
    cnlinear on it."""
    exe = os.path.join(REFDIR, "ref_restart_app")
    plug = ["-dll_append", os.path.join(REFDIR, "hostemu", "libfluca_nsb200.so"), "-ns_type", "b200"] + TIGHT
    dumps = {}
    for tag, extra in (("cnlinear", ["-stretch", "0.3"]), ("b200", ["-stretch", "0.3"] + plug), ("uniform", [])):
        path = str(tmp_path / f"{tag}.bin")
        r = subprocess.run([exe, "write", path, "-ns_time_step_size", "0.06", "-ns_max_steps", "3"] + extra, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout + r.stderr
        dumps[tag] = load_dump(path)
    for k in dumps["cnlinear"]:
        a, b = dumps["cnlinear"][k], dumps["b200"][k]
        assert np.abs(a - b).max() <= (1e-9 if k.startswith("Pressure") else 1e-10) * np.abs(a).max(), k
    assert np.abs(dumps["cnlinear"]["Velocity"] - dumps["uniform"]["Velocity"]).max() > 1e-2  # the stretching did reach the solver


ITER_CASES = [
    ("cavity_flow_3d", ["-cart_grid_x", "8", "-cart_grid_y", "7", "-cart_grid_z", "5", "-ns_time_step_size", "0.06", "-ns_max_steps", "2"]),
    ("cavity_flow_2d", ["-cart_grid_x", "16", "-cart_grid_y", "14", "-ns_time_step_size", "0.03", "-ns_max_steps", "3"]),
    ("taylor_green_vortex", ["-nsteps", "3", "-t_final", "0.15", "-periodic", "-cart_grid_x", "12", "-cart_grid_y", "12"]),
]


@pytest.mark.skipif(not _ready("hostemu"), reason="oracle/_ref is not built here")
@pytest.mark.parametrize("prog,args", ITER_CASES, ids=[c[0] for c in ITER_CASES])
def test_the_models_iterative_solvers_converge_to_its_exact_solves(tmp_path, prog, args):
    """`-model_solvers iterative` (oracle/ref_model/petsc_model_ksp.c: outer right-preconditioned GMRES(30) + the reference's PCABF
    with GMRES(30) + ILU(0) inside, what a serial PETSc run does by default) is the configuration bench.py times as the CPU figure
    of the reference's own sources.  Taken to convergence it must land on the exact solves the parity checks use; at PETSc's default
    tolerances (1e-5 everywhere, nssol.c:22-24) it stays within the accuracy those tolerances buy."""
    exact, _ = run(prog, args, tmp_path, "exact")
    tight, _ = run(prog, args + ["-model_solvers", "iterative", "-ns_ksp_rtol", "1e-12", "-ns_abf_momentum_ksp_rtol", "1e-13", "-ns_abf_schur_ksp_rtol", "1e-13"], tmp_path, "tight")
    loose, _ = run(prog, args + ["-model_solvers", "iterative"], tmp_path, "loose")
    for k in ("Velocity", "FaceNormalVelocity", "Pressure"):
        n = np.linalg.norm(exact[k])
        assert np.linalg.norm(tight[k] - exact[k]) <= 1e-8 * n, (k, np.linalg.norm(tight[k] - exact[k]) / n)
        assert np.linalg.norm(loose[k] - exact[k]) <= (5e-2 if k == "Pressure" else 2e-3) * n, (k, np.linalg.norm(loose[k] - exact[k]) / n)


@pytest.mark.skipif(not _ready("hostemu"), reason="oracle/_ref is not built here")
def test_b200_and_the_reference_both_at_default_tolerances(tmp_path):
    """What a user sees when switching types with no other option: the reference's cnlinear with iterative solvers at PETSc's
    defaults against -ns_type b200 at its defaults, in the reference's own cavity_flow_3d.  Neither is converged; they differ by what
    1e-5 tolerances leave open, not by more."""
    args = ["-cart_grid_x", "12", "-cart_grid_y", "12", "-cart_grid_z", "6", "-ns_time_step_size", "0.04", "-ns_max_steps", "4"]
    ref, _ = run("cavity_flow_3d", args + ["-model_solvers", "iterative"], tmp_path, "ref_defaults")
    dump = str(tmp_path / "b200_defaults.bin")
    r = subprocess.run([os.path.join(REFDIR, "cavity_flow_3d")] + args + ["-ns_monitor_solution", f"flucacgns:{dump}", "-dll_append", os.path.join(REFDIR, "hostemu", "libfluca_nsb200.so"), "-ns_type", "b200"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    b200 = load_dump(dump)
    for k, tol in (("Velocity", 2e-3), ("FaceNormalVelocity", 2e-3), ("Pressure", 5e-2)):
        n = np.linalg.norm(ref[k])
        assert np.linalg.norm(ref[k] - b200[k]) <= tol * n, (k, np.linalg.norm(ref[k] - b200[k]) / n)


def _program(prog, args, tmp_path, tag, b200):
    dump = str(tmp_path / f"{prog}_{tag}.bin")
    cmd = [os.path.join(REFDIR, prog)] + list(args) + ["-ns_monitor", "-ns_monitor_solution", f"flucacgns:{dump}"]
    if b200:
        cmd += ["-dll_append", os.path.join(REFDIR, "hostemu", "libfluca_nsb200.so"), "-ns_type", "b200"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    return r, (load_dump(dump) if os.path.exists(dump) else None)


@pytest.mark.skipif(not _ready("hostemu"), reason="oracle/_ref is not built here")
def test_one_command_line_for_both_types(tmp_path):
    """A command line written for the reference's type keeps its meaning when only the type changes: the tolerances go in under the
    reference's own option names (-ns_ksp_rtol for the KSP of the SNES, nssol.c:13-25; -ns_abf_momentum_ksp_rtol and
    -ns_abf_schur_ksp_rtol for the KSPs of PCABF, abfpc.c:33-46), b200 reads them (glue/nsb200.c NSSetFromOptions_B200), nothing is
    left unused, and both runs land on the same answer."""
    args = ["-cart_grid_x", "12", "-cart_grid_y", "10", "-cart_grid_z", "6", "-ns_time_step_size", "0.05", "-ns_max_steps", "3", "-model_solvers", "iterative",
            "-ns_ksp_rtol", "1e-12", "-ns_abf_momentum_ksp_rtol", "1e-13", "-ns_abf_schur_ksp_rtol", "1e-13"]
    out = {}
    for tag, b200 in (("cnlinear", False), ("b200", True)):
        r, out[tag] = _program("cavity_flow_3d", args, tmp_path, tag, b200)
        assert r.returncode == 0, r.stdout + r.stderr
        assert "options you set that were not used" not in r.stderr, r.stderr
    for k in ("Velocity", "FaceNormalVelocity", "Pressure", "PressureHalfStep"):
        a, b = out["cnlinear"][k], out["b200"][k]
        assert np.linalg.norm(a - b) <= 1e-8 * np.linalg.norm(a), (k, np.linalg.norm(a - b) / np.linalg.norm(a))


@pytest.mark.skipif(not _ready("hostemu"), reason="oracle/_ref is not built here")
def test_a_step_that_does_not_converge_meets_the_references_failure_policy(tmp_path):
    """-ns_ksp_max_it 1 with a tolerance one iteration cannot reach: the library reports DIVERGED, the glue sets ns->reason =
    NS_DIVERGED_NONLINEAR_SOLVE as NSCheckDiverged does (nsbasic.c:425-436) and the reference's NSStep applies its policy
    (nsbasic.c:288-297): by default the program stops with PETSC_ERR_NOT_CONVERGED; with -ns_error_if_step_failed 0 NSSolve ends
    quietly with the step counter and the time where they were."""
    args = ["-cart_grid_x", "10", "-cart_grid_y", "10", "-ns_time_step_size", "0.05", "-ns_max_steps", "3", "-ns_ksp_rtol", "1e-12", "-ns_ksp_max_it", "1"]
    r, _ = _program("cavity_flow_2d", args, tmp_path, "stop", True)
    assert r.returncode != 0 and "NSStep has failed due to" in r.stderr and "DIVERGED_NONLINEAR_SOLVE" in r.stderr, r.stdout + r.stderr
    r, _ = _program("cavity_flow_2d", args + ["-ns_error_if_step_failed", "0"], tmp_path, "quiet", True)
    assert r.returncode == 0, r.stdout + r.stderr
    mon = [ln for ln in r.stdout.splitlines() if " NS dt " in ln]
    assert mon and all(ln.startswith("0 NS dt 0.05 time 0") for ln in mon), mon  # NSSolve's monitors before and after: the failed step did not advance


@pytest.mark.skipif(not (_ready("hostemu") and os.path.exists(os.path.join(REFDIR, "fluca_app"))), reason="oracle/_ref is not built here")
@pytest.mark.parametrize("first,second", [("b200", "cnlinear"), ("cnlinear", "b200")])
def test_the_references_application_restarts_across_types(tmp_path, first, second):
    """The reference's APPLICATION (fluca/app/main.c, unmodified: lid-driven cavity, -ns_load_solution_from_file) on the PETSc model:
    four steps with cnlinear in one run against two steps with one type, the solution file NSMonitorSolution wrote, and a second
    run of the application that loads it (-ns_load_solution_from_file -> NSLoadSolution, nssol.c:176-204: step, time, PressureHalfStep)
    and continues to step 4 with the OTHER type."""
    app = os.path.join(REFDIR, "fluca_app")
    grid = ["-cart_grid_x", "12", "-cart_grid_y", "10", "-ns_monitor"]
    plug = ["-dll_append", os.path.join(REFDIR, "hostemu", "libfluca_nsb200.so"), "-ns_type", "b200"] + TIGHT

    def go(args, kind):
        r = subprocess.run([app] + grid + args + (plug if kind == "b200" else []), capture_output=True, text=True, timeout=600)
        assert r.returncode == 0 and "options you set that were not used" not in r.stderr, r.stdout + r.stderr
        return [ln for ln in r.stdout.splitlines() if " NS dt " in ln]

    full, half, cont = (str(tmp_path / f"{n}.bin") for n in ("full", "half", "cont"))
    go(["-ns_max_steps", "4", "-ns_monitor_solution", f"flucacgns:{full}"], "cnlinear")
    go(["-ns_max_steps", "2", "-ns_monitor_solution", f"flucacgns:{half}"], first)
    mon = go(["-ns_max_steps", "4", "-ns_load_solution_from_file", half, "-ns_monitor_solution", f"flucacgns:{cont}"], second)
    assert mon[0].startswith("2 NS dt 0.002 time 0.004") and mon[-1].startswith("4 NS"), mon  # picked up at step 2
    a, b = load_dump(full), load_dump(cont)
    for k in ("Velocity", "FaceNormalVelocity", "Pressure", "PressureHalfStep"):
        assert np.abs(a[k] - b[k]).max() <= 1e-10 * np.abs(a[k]).max(), (k, np.abs(a[k] - b[k]).max() / np.abs(a[k]).max())


@pytest.mark.skipif(not _ready("hostemu"), reason="oracle/_ref is not built here")
def test_inner_ksp_monitors_under_the_references_option_names(tmp_path):
    """-ns_abf_momentum_ksp_monitor / -ns_abf_schur_ksp_monitor (the KSPs of PCABF, abfpc.c:33-46) and -ns_ksp_monitor on a b200 run
    inside the reference's program: KSPMonitorResidual's lines, one block per inner solve, as many iterations as the statistics count."""
    args = ["-cart_grid_x", "8", "-cart_grid_y", "8", "-ns_time_step_size", "0.05", "-ns_max_steps", "2", "-ns_ksp_monitor", "-ns_abf_momentum_ksp_monitor", "-ns_abf_schur_ksp_monitor"]
    r, _ = _program("cavity_flow_2d", args, tmp_path, "mon", True)
    assert r.returncode == 0 and "options you set that were not used" not in r.stderr, r.stdout + r.stderr
    out = r.stdout.splitlines()
    blocks = {"momentum": 0, "schur": 0}
    its = {"momentum": 0, "schur": 0}
    cur = None
    for ln in out:
        if ln.startswith("    Residual norms for ns_abf_"):
            cur = "momentum" if "momentum" in ln else "schur"
            blocks[cur] += 1
        elif ln.startswith("    ") and " KSP Residual norm " in ln and cur:
            k = int(ln.split()[0])
            its[cur] += k > 0
        elif ln.startswith("  Residual norms for ns_ solve."):
            cur = None
    assert blocks["momentum"] >= 2 and blocks["schur"] >= 2
    summary = [ln for ln in out if "ns_abf_momentum_ solves:" in ln]  # the per-step totals the outer monitor prints
    assert len(summary) == 2
    tot_m = sum(int(ln.split("ns_abf_momentum_ solves:")[1].split()[0]) for ln in summary)
    tot_s = sum(int(ln.split("ns_abf_schur_ solves:")[1].split()[0]) for ln in summary)
    assert (its["momentum"], its["schur"]) == (tot_m, tot_s), (its, tot_m, tot_s)
    # without the options: silent
    r, _ = _program("cavity_flow_2d", args[:8], tmp_path, "quiet", True)
    assert "KSP Residual norm" not in r.stdout


@pytest.mark.skipif(not _ready("hostemu"), reason="oracle/_ref is not built here")
def test_log_events_of_the_type_show_next_to_the_references(tmp_path):
    """The glue registers PETSc log events for what it adds around the device work (NSB200HostToDevice / DeviceToHost / BoundaryData /
    DeviceStep); with the model's stand-in for -log_view (PETSC_MODEL_TIMING) they are listed next to the reference's own NSSetUp /
    NSStep events: one upload of the initial state, one download per step at the default sync interval, none with
    -ns_b200_sync_interval 0 until a viewer asks."""
    def events(extra):
        cmd = [os.path.join(REFDIR, "cavity_flow_2d"), "-cart_grid_x", "12", "-cart_grid_y", "12", "-ns_time_step_size", "0.04", "-ns_max_steps", "4",
               "-dll_append", os.path.join(REFDIR, "hostemu", "libfluca_nsb200.so"), "-ns_type", "b200"] + extra
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, PETSC_MODEL_TIMING="1"))
        assert r.returncode == 0, r.stdout + r.stderr
        return {ln.split()[4]: int(ln.split()[5]) for ln in r.stderr.splitlines() if ln.startswith("[PETSc model timing] event")}

    ev = events([])
    assert ev["NSStep"] == 4 and ev["NSB200DeviceStep"] == 4 and ev["NSB200HostToDevice"] == 1 and ev["NSB200DeviceToHost"] == 4 and ev["NSB200BoundaryData"] == 4, ev
    ev = events(["-ns_b200_sync_interval", "0"])
    assert ev["NSB200DeviceStep"] == 4 and "NSB200DeviceToHost" not in ev, ev  # nothing observes ns->sol: it stays on the device
