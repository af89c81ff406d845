"""Writes profiles/<name>: per-kernel counts of the SASS mnemonics that show how each kernel moves data (UTMALDG = TMA tensor
loads, SYNCS = mbarrier, LDG/STG = direct global access, fp64 arithmetic, shuffles, atomics), from cuobjdump -sass of the
product library.  Runs without a GPU.    python tools/sass_summary.py [profiles/r04_sass_summary.md]"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ("UTMALDG", "SYNCS", "LDG", "STG", "DFMA", "DADD", "DMUL", "BAR", "SHFL", "ATOM", "RED")


def main():
    out_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r04_sass_summary.md")
    lib = os.path.join(ROOT, "fluca_b200", "csrc", "libfluca_b200.so")
    txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    rows = []
    for part in re.split(r"\n\s*Function : ", txt)[1:]:
        name = part.split("\n", 1)[0].strip()
        cnt = {k: len(re.findall(r"\b" + k + r"\b", part)) for k in KEYS}
        rows.append((name, len(re.findall(r"/\*[0-9a-f]{4}\*/", part)), cnt))
    names = subprocess.run(["c++filt"], input="\n".join(r[0] for r in rows), capture_output=True, text=True).stdout.splitlines()
    out = ["# SASS evidence (cuobjdump -sass fluca_b200/csrc/libfluca_b200.so, sm_100a)", "",
           "Counts of the mnemonics that show how each kernel moves data: `UTMALDG` = TMA tensor load (cp.async.bulk.tensor), `SYNCS` = mbarrier",
           "arrive / try_wait, `LDG` / `STG` = direct global loads / stores, `DFMA` / `DADD` / `DMUL` = fp64 arithmetic, `SHFL` = warp shuffles",
           "(reductions), `ATOM+RED` = atomics.  Regenerate with `python tools/sass_summary.py`.", "",
           f"Kernels: {len(rows)}; UTMALDG total: {sum(r[2]['UTMALDG'] for r in rows)}.", "",
           "| kernel | SASS lines | UTMALDG | SYNCS | LDG | STG | DFMA | DADD | DMUL | BAR | SHFL | ATOM+RED |", "|---|---|---|---|---|---|---|---|---|---|---|---|"]
    order = sorted(range(len(rows)), key=lambda i: (-rows[i][2]["UTMALDG"], -rows[i][1]))
    for i in order:
        _, n, c = rows[i]
        d = re.sub(r"\(anonymous namespace\)::", "", names[i])
        d = d.split("(fluca::")[0][:140]
        out.append(f"| `{d}` | {n} | {c['UTMALDG']} | {c['SYNCS']} | {c['LDG']} | {c['STG']} | {c['DFMA']} | {c['DADD']} | {c['DMUL']} | {c['BAR']} | {c['SHFL']} | {c['ATOM'] + c['RED']} |")
    open(out_path, "w").write("\n".join(out) + "\n")
    print(out_path, len(rows), "kernels")


if __name__ == "__main__":
    main()
