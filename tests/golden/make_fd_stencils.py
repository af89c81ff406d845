"""Extracts tests/golden/fd_stencils.json from the reference's own FlucaFD tests: for every `test:` entry of
fluca/tests/fd/ex*.c (the /*TEST ... TEST*/ block PETSc's harness reads) the program, its arguments, and the stored expected
output fluca/tests/fd/output/<program>_<suffix>.out, which the reference compares byte for byte.  These are golden VECTORS
(arguments + printed stencils), not sources.

    python tests/golden/make_fd_stencils.py        # needs /root/reference; rewrites tests/golden/fd_stencils.json
"""
import json
import os
import re

REF = "/root/reference/fluca/tests/fd"
HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    out = {}
    for prog in sorted(f[:-2] for f in os.listdir(REF) if re.fullmatch(r"ex\d+\.c", f)):
        src = open(os.path.join(REF, prog + ".c")).read()
        block = src[src.index("/*TEST") : src.index("TEST*/")]
        for m in re.finditer(r"suffix:\s*(\S+)(.*?)(?=\n\s*test:|\Z)", block, flags=re.S):
            suffix, body = m.group(1), m.group(2)
            a = re.search(r"args:\s*(.*)", body)
            path = os.path.join(REF, "output", f"{prog}_{suffix}.out")
            if not os.path.exists(path):
                continue
            out[f"{prog}_{suffix}"] = {"program": prog, "args": a.group(1).split() if a else [], "output": open(path).read().splitlines()}
    path = os.path.join(HERE, "fd_stencils.json")
    json.dump(out, open(path, "w"), indent=1, sort_keys=True)
    print(path, len(out), "tests")


if __name__ == "__main__":
    main()
