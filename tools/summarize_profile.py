"""tools/summarize_profile.py TAG OUT.md -- turn gpurun_out/launches_TAG.csv and gpurun_out/prof_<kernel>_TAG.ncu-rep into the
markdown summary committed under profiles/ and refresh profiles/traffic.json (DRAM bytes per cell of the captured launches)."""
import collections
import csv
import io
import json
import re
import subprocess
import sys

TAG, OUT = sys.argv[1], sys.argv[2]
out = [f"# ncu capture {TAG}: sphere workload (BASELINE config 4 code path: IBM, outlet, flexible GMRES) at 256^3 on one B200", "",
       "command (tools/profile_final.sh): `python bench.py --n 256 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e`, first without ncu (exit 0), then",
       "`ncu --metrics gpu__time_duration.sum --clock-control none -s 1400 -c 2600` (launch list from the second step on) and",
       "`ncu --set full --clock-control none --import-source on -k regex:<kernel> -s 2 -c 1`.",
       "The 512^3 bench line runs the same kernels on 8x the cells; ncu replays at 512^3 save and restore ~150 GB per pass, so captures are taken at 256^3.", ""]
rows = list(csv.reader(open(f"gpurun_out/launches_{TAG}.csv")))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
h = rows[hi]
idx = {n: i for i, n in enumerate(h)}
agg = collections.defaultdict(lambda: [0, 0.0])
tot = n = 0
for r in rows[hi + 2:]:
    if len(r) < len(h):
        continue
    v = float(r[idx["Metric Value"]].replace(",", ""))
    u = r[idx["Metric Unit"]]
    v = v / 1000 if u == "ns" else (v * 1000 if u == "ms" else v)
    key = re.sub(r"fluca::|<unnamed>::|\(anonymous namespace\)::", "", r[idx["Kernel Name"]])
    key = re.sub(r"\(Solver &[^)]*\)", "()", key)
    key = re.sub(r"\(.*$", "", key)[:100]
    agg[key][0] += 1
    agg[key][1] += v
    tot += v
    n += 1
out += [f"## Launch list: {n} launches, {tot / 1000:.1f} ms of kernel time (cold-cache, serialised: compare shares)", "", "| share | total ms | launches | avg us | kernel |", "|---|---|---|---|---|"]
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:26]:
    out.append(f"| {t / tot * 100:.1f} % | {t / 1000:.2f} | {c} | {t / c:.1f} | `{k}` |")
out.append("")
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__shared_mem_per_block_dynamic", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "lts__t_sector_hit_rate.pct", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio"]
traffic = {}
f = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
for K in ("AApplyTile", "PoissonTile", "MGSmoothTile", "MGFirstTwoTile", "CoupledCells", "MGResidRestrict"):
    try:
        txt = subprocess.run(["ncu", "-i", f"gpurun_out/prof_{K}_{TAG}.ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rr = list(csv.reader(io.StringIO(txt)))
        hdr, units, r = rr[0], rr[1], rr[-1]
    except Exception as e:  # the kernel no longer launches in this build
        out.append(f"## {K}: no launch captured ({type(e).__name__})")
        out.append("")
        continue
    out += [f"## {K}  (`{r[hdr.index('Kernel Name')][:100]}`, grid {r[hdr.index('launch__grid_size')]})", "", "| metric | value | unit |", "|---|---|---|"]
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            out.append(f"| {w} | {r[i]} | {units[i]} |")
    i0, i1 = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    tb = float(r[i0].replace(",", "")) * f[units[i0]] + float(r[i1].replace(",", "")) * f[units[i1]]
    out += [f"| DRAM traffic per launch | {tb / 1e9:.3f} | GB |", ""]
    traffic[K] = tb
open(OUT, "w").write("\n".join(out) + "\n")
c = 256 ** 3
if "AApplyTile" in traffic and "PoissonTile" in traffic:
    d = {"_comment": "DRAM traffic per launch of the kernels bench.py reports a roofline for, from ncu --set full captures (dram__bytes_read.sum + dram__bytes_write.sum), expressed per cell so that bench.py can scale it to the launch it times. Written by tools/summarize_profile.py whenever the kernels change.",
         "momentum_apply": {"bytes_per_cell": round(traffic["AApplyTile"] / c * 108.0 / 120.0, 1), "algorithmic_bytes_per_cell": 108.0, "capture": f"{OUT} (256^3 sphere workload): {traffic['AApplyTile'] / 1e9:.3f} GB per launch for the variant with a separate dot partner (120 B/cell algorithmic, ratio {traffic['AApplyTile'] / c / 120.0:.3f}); scaled by that ratio to the 108 B/cell mean of the two applies of a BiCGStab iteration"},
         "poisson_apply": {"bytes_per_cell": round(traffic["PoissonTile"] / c, 1), "algorithmic_bytes_per_cell": 24.0, "capture": f"{OUT} (256^3 sphere workload, BiCGStab variant that also reads the shadow residual: {traffic['PoissonTile'] / 1e9:.3f} GB per launch)"}}
    json.dump(d, open("profiles/traffic.json", "w"), indent=1)
print("\n".join(out[:40]))
