"""The reference-side binding glue/nsb200.c (NS type "b200") cannot be compiled here for real -- PETSc, MPI and
the Fluca headers it includes are absent from this image -- so it is compiled with `gcc -fsyntax-only` against a
declaration-only stub of the API subset it uses (tests/petsc_stub/), with include/fluca_b200.h as the real header:
every call into the C ABI is type-checked against the declarations the CUDA library exports."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_glue_compiles_against_the_api_stub():
    cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    r = subprocess.run([cc, "-std=gnu11", "-fsyntax-only", "-Wall", "-Wextra", "-Werror", "-Wno-unused-parameter", "-I", os.path.join(ROOT, "tests", "petsc_stub"), "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "glue", "nsb200.c")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_glue_compiles_with_the_matrix_free_hooks():
    """the same file against a base class carrying glue/patches/0001-ns-matrix-free-type-hooks.patch"""
    cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    r = subprocess.run([cc, "-std=gnu11", "-fsyntax-only", "-Wall", "-Wextra", "-Werror", "-Wno-unused-parameter", "-DFLUCA_NS_HAS_MATRIXFREE", "-I", os.path.join(ROOT, "tests", "petsc_stub"), "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "glue", "nsb200.c")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_matrix_free_patch_applies_to_the_reference(tmp_path):
    """glue/patches/0001-...: a real unified diff against nsbasic.c:153-299 / nsimpl.h (checked where the reference is present)"""
    import pytest
    import shutil

    ref = "/root/reference/fluca"
    if not os.path.isdir(ref):
        pytest.skip("the reference tree is not on this machine")
    for rel in ("src/ns/interface/nsbasic.c", "include/fluca/private/nsimpl.h"):
        dst = tmp_path / "fluca" / rel
        dst.parent.mkdir(parents=True, exist_ok=True)
        shutil.copy(os.path.join(ref, rel), dst)
    r = subprocess.run(["patch", "-p1", "--dry-run", "-i", os.path.join(ROOT, "glue", "patches", "0001-ns-matrix-free-type-hooks.patch")], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr


def test_glue_compiles_against_the_references_own_headers():
    """Where the reference tree is present: glue/nsb200.c against the reference's REAL Fluca headers (fluca/include: nsimpl.h with
    struct _p_NS, flucans.h, flucameshcart.h, flucansbc.h, flucaviewer.h) -- only the PETSc headers underneath come from the model of
    oracle/ref_model/.  Every Fluca function the glue calls is thereby checked against the declaration the reference itself compiles with."""
    import pytest

    ref = "/root/reference/fluca/include"
    if not os.path.isdir(ref):
        pytest.skip("the reference tree is not on this machine")
    cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    for extra in ([], ["-DFLUCA_NS_HAS_MATRIXFREE_CHECK_ONLY"]):
        r = subprocess.run([cc, "-std=gnu11", "-fsyntax-only", "-Wall", "-Wextra", "-Werror", "-Wno-unused-parameter", "-Wno-unused-variable", "-I", os.path.join(ROOT, "oracle", "ref_model", "include"), "-I", ref, "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "glue", "nsb200.c")] + extra, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr


def test_glue_fills_every_ns_op_and_registers_the_type():
    src = open(os.path.join(ROOT, "glue", "nsb200.c")).read()
    for op in ("setfromoptions", "setup", "step", "formjacobian", "formfunction", "destroy", "view", "viewsolution", "loadsolution"):  # nsimpl.h:21-31
        assert re.search(rf"ns->ops->{op}\s*=\s*NS\w+_B200;", src), op
    assert 'NSRegister(NSB200, NSCreate_B200)' in src
    # every C-ABI symbol the glue calls is declared in the public header
    hdr = open(os.path.join(ROOT, "include", "fluca_b200.h")).read()
    for sym in set(re.findall(r"\b(fluca_b200_\w+)\s*\(", src)):
        assert re.search(rf"\b{sym}\s*\(", hdr), sym
