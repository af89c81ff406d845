"""tools/summarize_capture.py TAG OUT.md -- turn gpurun_out/<TAG>_launches_sphere256.csv, gpurun_out/<TAG>_prof_256.ncu-rep (one
--set full capture holding several kernels) and, if present, gpurun_out/<TAG>_prof_512_aapply.ncu-rep (one launch of the
momentum tile kernel at 512^3) into the markdown summary committed under profiles/, and refresh profiles/traffic.json (DRAM bytes
per cell of the captured momentum / Poisson launches, which bench.py scales to the launch it times)."""
import collections
import csv
import io
import json
import os
import re
import subprocess
import sys

TAG, OUT = sys.argv[1], sys.argv[2]
F = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_sector_hit_rate.pct", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"]


def short(name):
    name = re.sub(r"fluca::|<unnamed>::|\(anonymous namespace\)::", "", name)
    name = re.sub(r"\(Solver &[^)]*\)", "()", name)
    return re.sub(r"\(.*$", "", name)[:100]


def raw(path):
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(io.StringIO(txt)))
    return rr[0], rr[1], rr[2:]


out = [f"# ncu captures {TAG}: sphere workload (BASELINE config 4 code path: IBM, outlet, flexible GMRES) on one B200", "",
       f"commands (tools/gpu_{TAG}.sh), each after the same command had exited 0 without ncu: launch list `ncu --metrics gpu__time_duration.sum --clock-control none -c 3000`",
       "of `python bench.py --n 256 --steps 1 --warmup 1 ...`; `ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:...`",
       "at 256^3; and one launch of the momentum tile kernel at 512^3 (`-c 1`).  Per-launch times under ncu are cold-cache and serialised: compare shares.", ""]
lst = f"gpurun_out/{TAG}_launches_sphere256.csv"
if os.path.exists(lst):
    rows = list(csv.reader([ln for ln in open(lst) if not ln.startswith("==")]))
    h = rows[0]
    idx = {n: i for i, n in enumerate(h)}
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
    tot = n = 0
    for r in rows[1:]:
        if len(r) < len(h) or r[idx["Metric Name"]] != "gpu__time_duration.sum":
            continue
        v = float(r[idx["Metric Value"]].replace(",", ""))
        u = r[idx["Metric Unit"]]
        v = v / 1000 if u == "ns" else (v * 1000 if u == "ms" else v)
        k = short(r[idx["Kernel Name"]])
        agg[k][0] += 1
        agg[k][1] += v
        agg[k][2] = max(agg[k][2], v)
        tot += v
        n += 1
    out += [f"## Launch list at 256^3: {n} launches, {tot / 1000:.1f} ms of kernel time (2 steps: 1 warm-up + 1 timed)", "", "| share | total ms | launches | avg us | max us | kernel |", "|---|---|---|---|---|---|"]
    for k, (c, t, m) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:28]:
        out.append(f"| {t / tot * 100:.1f} % | {t / 1000:.2f} | {c} | {t / c:.1f} | {m:.1f} | `{k}` |")
    out.append("")
traffic = {}
rep = f"gpurun_out/{TAG}_prof_256.ncu-rep"
if os.path.exists(rep):
    hdr, units, rows = raw(rep)
    ti = hdr.index("gpu__time_duration.sum")
    best = {}
    for r in rows:  # the longest instance of every kernel = its finest-level / full-grid launch
        k = short(r[hdr.index("Kernel Name")])
        t = float(r[ti].replace(",", ""))
        if k not in best or t > best[k][0]:
            best[k] = (t, r)
    out += ["## Full captures at 256^3 (the longest launch of every captured kernel = its full-grid instance)", ""]
    for k, (t, r) in sorted(best.items(), key=lambda kv: -kv[1][0]):
        out += [f"### `{k}`  (grid {r[hdr.index('launch__grid_size')]})", "", "| metric | value | unit |", "|---|---|---|"]
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                out.append(f"| {w} | {r[i]} | {units[i]} |")
        i0, i1 = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        tb = float(r[i0].replace(",", "")) * F[units[i0]] + float(r[i1].replace(",", "")) * F[units[i1]]
        out += [f"| DRAM traffic per launch | {tb / 1e9:.3f} | GB ({tb / 256**3:.1f} B per cell) |", ""]
        traffic[k] = tb / 256**3
rep5 = f"gpurun_out/{TAG}_prof_512_aapply.ncu-rep"
t512 = None
if os.path.exists(rep5):
    hdr, units, rows = raw(rep5)
    r = rows[-1]
    out += [f"## Momentum tile kernel at 512^3, one launch (`{short(r[hdr.index('Kernel Name')])}`, grid {r[hdr.index('launch__grid_size')]})", "", "| metric | value | unit |", "|---|---|---|"]
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            out.append(f"| {w} | {r[i]} | {units[i]} |")
    i0, i1 = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    tb = float(r[i0].replace(",", "")) * F[units[i0]] + float(r[i1].replace(",", "")) * F[units[i1]]
    t512 = tb / 512**3
    out += [f"| DRAM traffic per launch | {tb / 1e9:.3f} | GB ({t512:.1f} B per cell; algorithmic 120 B per cell with the dot partner: ratio {t512 / 120.0:.3f}) |", ""]
open(OUT, "w").write("\n".join(out) + "\n")
ma = [v for k, v in traffic.items() if "AApplyTile" in k]
pa = [v for k, v in traffic.items() if "PoissonTile" in k]
if t512 or ma:
    per = t512 if t512 else ma[0]
    d = {"_comment": "DRAM traffic per launch of the kernels bench.py reports a roofline for, from ncu --set full captures (dram__bytes_read.sum + dram__bytes_write.sum), per cell, so that bench.py can scale it to the launch it times. Written by tools/summarize_capture.py.",
         "momentum_apply": {"bytes_per_cell": round(per, 1), "algorithmic_bytes_per_cell": 120.0, "capture": f"{OUT}: {'one launch at 512^3' if t512 else 'launch at 256^3'} of k_tma_march<AApplyTile<4>> (with the dot partner r^: 120 B per cell algorithmic, ratio {per / 120.0:.3f}); the wall launches (< 1 % of the cells) are not in it"}}
    if pa:
        d["poisson_apply"] = {"bytes_per_cell": round(pa[0], 1), "algorithmic_bytes_per_cell": 24.0, "capture": f"{OUT} (256^3, BiCGStab form that also reads the shadow residual)"}
    json.dump(d, open("profiles/traffic.json", "w"), indent=1)
print("\n".join(out[:12]))
