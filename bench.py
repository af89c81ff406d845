#!/usr/bin/env python
"""bench.py -- Mcell-updates/s of one Navier-Stokes time step (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W                (our arm: CUDA kernels on N B200s)
  python bench.py --impl reference --gpus N --steps K --warmup W   (reference arm: CPU restatement)

One "step" = one full NS time step (NSStep) of the configured workload.  Prints ONE JSON line.
Workload: BASELINE config 4 (3-D sphere by IBM, 512^3, 100k markers).  At N > 1 the headline `value` is STRONG scaling of
that one grid (what config 4 names: "512^3 grid ... at 1/2/4/8 GPUs"), and the same line carries the weak-scaling run
(512^3 per GPU) under "weak".  Before anything is timed an N-rank parity case is run against the CPU oracle ("parity").
See DESIGN.md section "Measurement" for every definition used here.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


print_json = print
T0 = time.time()


def note(msg):
    """progress marker on stderr (rank-tagged, seconds since start): a stuck multi-GPU run shows where it stopped"""
    sys.stderr.write(f"[bench r{os.environ.get('RANK', '0')} +{time.time() - T0:6.1f}s] {msg}\n")
    sys.stderr.flush()


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="sphere", choices=["sphere", "cavity", "channel"], help="sphere: BASELINE config 4 (3-D IBM sphere Re=300, 512^3, 100k markers), the configuration the metric is quoted on; cavity: config 3 (256^3 lid-driven cavity, no IBM); channel: config 5 (multi-body channel, periodic x, 2048x1024x1024 on 8 GPUs = 268 M cells per GPU, 1 M markers; weak scaling)")
    ap.add_argument("--n", type=int, default=0, help="cells per direction (default: 512 sphere, 256 cavity; channel: cells in y and z, x = 2 n)")
    ap.add_argument("--markers", type=int, default=100000)
    ap.add_argument("--scaling", default="both", choices=["both", "strong", "weak"], help="N > 1: strong (the one n^3 grid split in z-slabs; the headline), weak (n^3 per GPU), or both (default)")
    ap.add_argument("--strong", action="store_true", help="same as --scaling strong")
    ap.add_argument("--mode", default="coupled", choices=["coupled", "fractional"])
    ap.add_argument("--restart", type=int, default=0, help="outer GMRES restart (memory: (restart+1) x 7 fields; default 3 at 512^3, 10 below; the flexible form also keeps restart x 7 fields of preconditioned vectors)")
    ap.add_argument("--schur-ainv", default="ID", choices=["ID", "DIAG", "ROWSUM"], help="-ns_pc_abf_schur_ainv_type (abfpc.c:246); the headline numbers use the reference default ID")
    ap.add_argument("--upper-ainv", default="ID", choices=["ID", "DIAG", "ROWSUM"], help="-ns_pc_abf_upper_ainv_type (abfpc.c:247)")
    ap.add_argument("--cpu-n", type=int, default=64, help="cells per direction of the bounded CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-reference-sources", action="store_true", help="skip the timed sample of the reference's own sources (oracle/_ref) inside the CPU legs")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the untimed N-rank parity case against the oracle")
    a = ap.parse_args()
    if a.strong:
        a.scaling = "strong"
    return a


# ---------------------------------------------------------------------------------------------- workload
def default_n(args):
    return args.n or {"sphere": 512, "cavity": 256, "channel": 1024}[args.workload]


def make_case(args, n, nz):
    """n: cells in x and y (channel: y and z-per-unit), nz: global cells in z."""
    from fluca_b200 import workloads as W

    if args.workload == "sphere":
        return W.sphere_bench_case(n, nz)
    if args.workload == "cavity":
        return W.cavity_bench_case(n, nz)
    # config 5: 2048 x 1024 x 1024 on 8 GPUs (h = 1/64); the slab direction z carries the GPUs.  The x extent is 2 n.
    return W.channel_bench_case((2 * n, n, nz))


def markers_for(args, case, n, nz):
    from fluca_b200 import workloads as W

    if args.workload == "sphere":
        return W.sphere_markers((0.0, 0.0, 0.0), 1.0, args.markers, 16.0 / n)
    if args.workload == "channel":
        # 80 spheres x 12 500 markers at full size (1.0 M); the count scales with the volume of the box that is run
        h = case.hi[0] / case.n[0]
        full = 2048.0 * 1024 * 1024
        nsph = max(1, int(round(80 * (case.n[0] * case.n[1] * case.n[2]) / full)))
        centres = W.channel_sphere_centres(case.lo, case.hi, nsph)
        return W.multi_sphere_markers(centres, 1.0, 12500, h)
    return None


def set_initial_state(args, case, solver):
    """Uniform stream for the external flows (zero state for the cavity), written slab by slab: every rank only ever
    allocates its own planes on the host."""
    from fluca_b200 import workloads as W

    if args.workload in ("sphere", "channel"):
        v, U, p = W.uniform_inflow_state(case, slab=(solver.k0, solver.nzl, solver.last_z))
        solver.set_state(v=v, U=U, p=p, phalf=p)


def workload_text(args, case, scaling, world, restart, nmark):
    nx, ny, nz = case.n
    tol = f"NS type b200 mode={args.mode}, reference default tolerances (outer/momentum/Schur rtol 1e-5; inner tolerances relaxed by the inexact-Krylov rule as the outer residual drops, DESIGN.md 5), flexible GMRES restart {restart}"
    part = "one GPU" if world == 1 else (f"STRONG scaling: the one grid in {world} z-slabs of {nz // world} planes" if scaling == "strong" else f"WEAK scaling: {nz // world} planes ({nx}x{ny}x{nz // world} cells) per GPU, {world} z-slabs")
    if args.workload == "sphere":
        return f"BASELINE config 4: 3-D flow past a sphere by IBM, Re=300, {nx}x{ny}x{nz} cells, {part}, h=16/{nx}, {nmark} Fibonacci markers, 4-point delta, inflow/pressure-outlet/symmetry, dt=0.5h, uniform initial state, {tol}"
    if args.workload == "channel":
        return f"BASELINE config 5: 3-D multi-body channel by IBM, {nx}x{ny}x{nz} cells, {part}, h=32/{nx}, periodic x and z, no-slip walls in y, {nmark} markers on spheres of D=1 (12500 Fibonacci markers each), bulk velocity 1, dt=0.5h, {tol}"
    return f"BASELINE config 3: 3-D lid-driven cavity Re=400, {nx}x{ny}x{nz} cells, {part}, dt=0.5h, zero initial state, {tol}"


# ---------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        self.index, self.samples, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        threading.Thread(target=self._read, daemon=True).start()

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.samples:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------- CPU arm
def reference_build_probe():
    """a PETSc build of the reference may be provided on the GPU box; it never was.  oracle/_ref holds the
    reference's NS sources compiled on a single-rank PETSc MODEL: the checker of the oracle, and (cpu_baseline.reference_sources) a
    one-core timing of the reference's own assembly + PCABF with the model's GMRES(30) + ILU(0) -- not a PETSc build."""
    petsc = os.environ.get("PETSC_DIR")
    ref = os.path.isdir(os.path.join(ROOT, "baseline", "_ref"))
    model = os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libfluca_ref_ns.so"))
    note = "; oracle/_ref = the reference's NS sources on a single-rank PETSc model (parity checker; timed on one core as cpu_baseline.reference_sources)" if model else ""
    if petsc or ref:
        return f"present (PETSC_DIR={petsc!r}, baseline/_ref={ref}) but not used: no build recipe for the reference exists in this repo (it needs MPI, HDF5 and parallel CGNS as well)" + note
    return "absent (no $PETSC_DIR, no baseline/_ref): the reference needs PETSc >= 3.23 + MPI + HDF5 + CGNS; the CPU arm is the repo's C restatement (oracle/, kind 'port')" + note


def cpu_sample(args, n, steps, warmup, mode):
    """The CPU restatement of the reference algorithm (oracle, kind "port": assembled CSR operators, GMRES(30) + block-Jacobi
    ILU(0), reference default tolerances) on an n^3 sample of the sphere / cavity workload, on all host threads.  The thread
    count is forced through OpenMP's API (a launcher such as torchrun pre-sets OMP_NUM_THREADS=1) and what a parallel
    region actually gets is what is reported."""
    # all host cores, up to 32: beyond that the block-Jacobi ILU of a 64^3 sample has more blocks than it can use and the
    # OpenMP barriers of its short loops cost more than they save (the count actually used is what the line reports)
    want = min(os.cpu_count() or 1, 32)
    os.environ["OMP_NUM_THREADS"] = str(want)
    from oracle import oracle as O
    from tests import cases

    threads = O.set_threads(want)
    sub = argparse.Namespace(**vars(args))
    if sub.workload == "channel":
        sub.workload = "sphere"
    case = make_case(sub, n, n)
    orc = cases.make_oracle_fast(case)
    nm = 0
    if sub.workload == "sphere":
        orc.set_state(*cases.uniform_inflow_state(case))
        nm = max(64, int(args.markers * (n / 512.0) ** 2))
        mk = cases.sphere_markers((0.0, 0.0, 0.0), 1.0, nm, 16.0 / n)
        orc.set_markers(mk["X"], mk["Ud"], mk["dV"], 4)
    else:
        orc.set_state(*case.initial_state())
    ainv = {"ID": 0, "DIAG": 1, "ROWSUM": 2}
    opt = O.default_options(mode=0 if mode == "coupled" else 1, ilu_blocks=threads, schur_ainv=ainv[args.schur_ainv], upper_ainv=ainv[args.upper_ainv])
    infos = []
    for _ in range(warmup):
        orc.step(opt)
    t0 = time.perf_counter()
    for _ in range(steps):
        infos.append(orc.step(opt))
    dt = time.perf_counter() - t0
    cells = float(n) ** 3
    res = dict(value=cells * steps / dt / 1e6, seconds=dt, steps=steps, warmup=warmup, n=n, threads=threads, markers=nm, outer=[i.outer_its for i in infos], mom=[i.mom_its for i in infos], schur=[i.schur_its for i in infos])
    res["reference_sources"] = reference_sources_sample(case, orc.get_state(), mode, args)
    return res


def reference_sources_sample(case, state, mode, args, budget_s=20.0):
    """The REFERENCE'S OWN NS sources timed on the same sample: oracle/_ref/libfluca_ref_ns.so (cartdiscret.c, cnlinear*.c, abfpc.c
    of thecasterian/fluca compiled on the single-rank PETSc model of oracle/ref_model/, built where /root/reference is and shipped
    with the repo) with the solver stack a serial PETSc run has by default -- outer right-preconditioned GMRES(30) on the true
    residual + PCABF with GMRES(30) + ILU(0) inside, rtol 1e-5 everywhere (ref_model/petsc_model_ksp.c) -- on ONE core.  It starts
    from the flow field the port has just produced (for the sphere: the developing flow around the immersed body); the reference
    has no immersed boundary, so its steps run without the forcing.  What is timed is the reference's assembly
    (MatSetValuesStencil per entry, every step), its PCSetUp_ABF (three sparse products per step) and its solves on a MODEL of PETSc:
    the reference's algorithm and code, not PETSc's performance.  Bounded: one warm-up step, then steps until 2 are done or
    budget_s is spent.  Reported beside the port, never instead of it."""
    try:
        from oracle import ref as R

        if getattr(args, "no_reference_sources", False):
            return {"available": False, "note": "--no-reference-sources"}
        if not os.path.exists(R.LIB):
            return {"available": False, "note": "oracle/_ref/libfluca_ref_ns.so is not on this machine (it is built from /root/reference)"}
        from oracle import oracle as O  # its BC container only

        R.set_inner_solvers(True)
        try:
            ref = R.Reference(case.n, case.faces(), case.rho, case.mu, case.dt, [O.BC(b["type"], velocity=b["velocity"], pressure=b["pressure"]) for b in case.bcs])
            ref.set_state(state["v"], state["U"], state["p"], state["phalf"])
            rmode = R.GMRES_ABF if mode == "coupled" else R.ABF_ONCE
            ref.step(mode=rmode, rtol=1e-5, maxit=200)  # warm-up
            m0, s0 = ref.inner_iterations()
            outer, t0 = [], time.perf_counter()
            while len(outer) < 2 and (not outer or time.perf_counter() - t0 < budget_s / 2):
                outer.append(int(ref.step(mode=rmode, rtol=1e-5, maxit=200)[0]))
            dt = time.perf_counter() - t0
            m1, s1 = ref.inner_iterations()
            del ref
        finally:
            R.set_inner_solvers(False)
        cells = float(np.prod(case.n))
        return {
            "available": True,
            "value": cells * len(outer) / dt / 1e6,
            "unit": "Mcell-updates/s",
            "cores": 1,
            "kind": "reference sources on a PETSc model",
            "sample": f"{'x'.join(str(k) for k in case.n)} sample from the port's final state, no immersed boundary (the reference has none), 1 warm-up + {len(outer)} timed step(s) in {dt:.1f} s, outer its {outer}, momentum its {m1 - m0}, Schur its {s1 - s0} (GMRES(30) + ILU(0), rtol 1e-5)",
        }
    except Exception as e:  # the checker's library is optional equipment: never lose the line over it
        return {"available": False, "error": f"{type(e).__name__}: {e}"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = args.cpu_n
    r = cpu_sample(args, n, args.steps, args.warmup, args.mode)
    sub = argparse.Namespace(**vars(args))
    full = default_n(args)
    case = make_case(sub, full, full if args.workload != "channel" else full)
    line = {
        "impl": "reference",
        "metric": "Mcell-updates/s per NS step",
        "value": r["value"],
        "unit": "Mcell-updates/s",
        "n_gpus": args.gpus,
        "steps": args.steps,
        "warmup": r["warmup"],
        "ms_per_step": 1e3 * r["seconds"] / r["steps"],
        "higher_is_better": True,
        "scaling": "strong" if args.scaling != "weak" and args.workload != "channel" else "weak",
        "vs_baseline": None,
        "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": workload_text(args, case, "strong", 1, args.restart or 3, args.markers) + f"; CPU arm: bounded sample {n}^3 of the same case ({r['markers']} markers: scaled with the surface cell count)", "mode": args.mode},
        "cpu_baseline": {"value": r["value"], "unit": "Mcell-updates/s", "cores": r["threads"], "kind": "port", "sample": f"{n}^3 sample of the workload, {r['warmup']} warm-up + {r['steps']} timed steps on {r['threads']} OpenMP threads (of {os.cpu_count()} host CPUs), outer its {r['outer']}, momentum its {r['mom']}, Schur its {r['schur']}", "reference_build": reference_build_probe(), "reference_sources": r["reference_sources"]},
        "e2e": {"value": r["value"], "unit": "Mcell-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print_json(json.dumps(line))


# ---------------------------------------------------------------------------------------------- GPU arm
class Ctx:
    """torch / torch.distributed plumbing of one rank."""

    def __init__(self):
        import torch
        import torch.distributed as dist

        self.torch, self.dist = torch, dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=self.dev)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def comm(self):
        """NS communicator description: every solver gets its own NCCL communicator (a fresh unique id from rank 0)."""
        if self.world == 1:
            return None
        import fluca_b200 as fb

        torch, dist, rank, world, dev = self.torch, self.dist, self.rank, self.world, self.dev

        def factory(L):
            if rank == 0:
                tid = torch.from_numpy(np.frombuffer(fb.Comm.unique_id(L), dtype=np.uint8).copy()).to(dev)
            else:
                tid = torch.zeros(128, dtype=torch.uint8, device=dev)
            dist.broadcast(tid, 0)
            return fb.Comm.nccl(L, tid.cpu().numpy().tobytes(), rank, world)

        return dict(rank=rank, nranks=world, make_comm=factory)

    def stream_timer(self, solver):
        """(start, stop) around the timed region: CUDA events on the solver's own stream; stop() returns milliseconds."""
        torch = self.torch
        stream = torch.cuda.ExternalStream(solver.stream(), device=self.dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

        def start():
            e0.record(stream)

        def stop():
            e1.record(stream)
            e1.synchronize()
            return e0.elapsed_time(e1)

        return start, stop

    def pinned(self, shape):
        return self.torch.empty(shape, dtype=self.torch.float64).pin_memory().numpy()

    def gather_z(self, arrs, axis):
        """rank 0 gets the z-concatenation of every rank's slab (small arrays only).  One node: through files in a scratch
        directory and a barrier, not through an object collective."""
        if self.world == 1:
            return arrs
        import tempfile

        self._gather_id = getattr(self, "_gather_id", 0) + 1
        d = os.path.join(tempfile.gettempdir(), f"fluca_b200_bench_{os.environ.get('MASTER_PORT', '0')}")
        os.makedirs(d, exist_ok=True)
        np.save(os.path.join(d, f"g{self._gather_id}_r{self.rank}.npy"), arrs)
        self.barrier()
        out = None
        if self.rank == 0:
            out = np.concatenate([np.load(os.path.join(d, f"g{self._gather_id}_r{r}.npy")) for r in range(self.world)], axis=axis)
        self.barrier()
        if os.path.exists(os.path.join(d, f"g{self._gather_id}_r{self.rank}.npy")):
            os.remove(os.path.join(d, f"g{self._gather_id}_r{self.rank}.npy"))
        return out


def parity_selfcheck(ctx, lib):
    """Untimed N-rank correctness of what is about to be timed: the sphere workload builder (inflow / outlet / symmetry /
    immersed boundary), coupled mode, tight tolerances, on the NCCL slab partition of this run, against the CPU oracle on
    rank 0 (the oracle is the checker here, never the thing measured).  Two steps with the reference-default ABF factors
    and one step with the DIAG / ROWSUM variants."""
    import fluca_b200 as fb
    from fluca_b200 import workloads as W

    world, rank = ctx.world, ctx.rank
    nx, ny, nz = 32, 16, max(16, 8 * world)
    case = W.sphere_bench_case(ny, nz)
    case.n, case.dt = (nx, ny, nz), 0.5 * 16.0 / nx
    case.hi = (12.0, 8.0, 8.0)  # the same box whatever the rank count: z in [-8, 8]
    mk = W.sphere_markers((0.0, 0.0, 0.0), 1.2, 300, 16.0 / ny)
    state = W.uniform_inflow_state(case)
    tight = {"ns_ksp_rtol": 1e-13, "ns_abf_momentum_ksp_rtol": 1e-13, "ns_abf_schur_ksp_rtol": 1e-13, "ns_ksp_max_it": 60}
    runs = (("ID/ID", "ID", "ID", 2), ("DIAG/ROWSUM", "DIAG", "ROWSUM", 1))
    got = {}
    for label, sa, ua, nsteps in runs:
        note(f"parity case {label}: building the solver")
        ns = W.make_ns(case, lib, "coupled", comm=ctx.comm(), ns_pc_abf_schur_ainv_type=sa, ns_pc_abf_upper_ainv_type=ua, **tight)
        W.set_initial_slab(ns, state)
        fb.NSB200SetMarkers(ns, mk["X"], mk["Ud"], mk["dV"], 4)
        outer = []
        for _ in range(nsteps):
            fb.NSStep(ns)
            outer.append(fb.NSB200GetStats(ns).outer_its)
        st = fb.NSB200GetSolver(ns).get_state()
        g = dict(v=ctx.gather_z(st["v"], 1), p=ctx.gather_z(st["p"], 0), U=[ctx.gather_z(u, 0) for u in st["U"]], outer=outer)
        fb.NSDestroy(ns)
        got[label] = g
    note("parity case: GPU side done")
    if rank != 0:
        return None
    from oracle import oracle as O  # checker
    from tests import cases

    # a 16 k-cell case at tight tolerances is ~1e5 short OpenMP regions: a handful of threads, not the 100+ of an 8-GPU host
    O.set_threads(min(os.cpu_count() or 1, 8))

    def rel(a, b):
        return float(np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-300))

    res, ok, tol = {}, True, 1e-10
    for label, sa, ua, nsteps in runs:
        orc = cases.make_oracle_fast(case)
        orc.set_state(*state)
        orc.set_markers(mk["X"], mk["Ud"], mk["dV"], 4)
        opt = O.default_options(mode=0, outer_rtol=1e-13, mom_rtol=1e-13, schur_rtol=1e-13, schur_ainv={"ID": 0, "DIAG": 1, "ROWSUM": 2}[sa], upper_ainv={"ID": 0, "DIAG": 1, "ROWSUM": 2}[ua])
        oo = [orc.step(opt).outer_its for _ in range(nsteps)]
        a, g = orc.get_state(), got[label]
        eU = float(np.sqrt(sum(np.sum((x - y) ** 2) for x, y in zip(g["U"], a["U"]))) / np.sqrt(sum(np.sum(y**2) for y in a["U"])))
        e = dict(v=rel(g["v"], a["v"]), U=eU, p=rel(g["p"], a["p"]), steps=nsteps, outer_its_gpu=g["outer"], outer_its_oracle=oo)
        ok = ok and e["v"] <= tol and e["U"] <= tol and e["p"] <= 10 * tol
        res[label] = e
    note("parity case: oracle side done")
    return {"ok": bool(ok), "ranks": world, "tolerance": "rel. L2 <= 1e-10 (v, U), 1e-9 (p) against the CPU oracle at tight tolerances", "case": f"sphere workload builder at {nx}x{ny}x{nz} (inflow/outlet/symmetry, 300 IBM markers), coupled mode, {world} z-slab(s) over {'NCCL' if world > 1 else 'one GPU'}; ABF factor variants schur/upper", "runs": res, "oracle": "this repo's restatement, pinned to the reference's own NS sources compiled on a PETSc model (operators, RHS, ABF, steps: tests/test_oracle_vs_reference.py); PETSc's solver arithmetic and the IBM section are unpinned (DESIGN.md 2)"}


def reference_selfcheck(lib):
    """Rank 0 only, untimed, no collective: the library about to be timed against the REFERENCE'S OWN compiled NS sources
    (oracle/_ref/libfluca_ref_ns.so: cartdiscret.c, cnlinear*.c, abfpc.c of thecasterian/fluca on a single-rank PETSc model, built in the
    container where /root/reference lives and shipped with the repo).  The sphere workload's box and boundary set
    (inflow / pressure outlet / symmetry) without markers -- the reference has no IBM -- at 8 x 5 x 5, seeded smooth initial state, two steps: one ABF application per
    step against -ns_ksp_type preonly, and the coupled solve against the exact solution of the reference's J x = b."""
    import fluca_b200 as fb
    from fluca_b200 import workloads as W

    from oracle import ref as R  # checker

    if not os.path.exists(R.LIB):
        return {"available": False, "note": "oracle/_ref/libfluca_ref_ns.so is not on this machine (it is built from /root/reference)"}
    from oracle import oracle as O  # its BC container only

    case = W.sphere_bench_case(5, 5)
    case.n, case.dt, case.hi = (8, 5, 5), 0.5 * 16.0 / 8, (12.0, 8.0, 8.0)
    state = case.initial_state(seed=5)  # a seeded smooth field: the uniform stream is an exact solution of the empty box (p = 0)
    tight = {"ns_ksp_rtol": 1e-13, "ns_abf_momentum_ksp_rtol": 1e-13, "ns_abf_schur_ksp_rtol": 1e-13, "ns_ksp_max_it": 60}

    def rel(a, b):
        return float(np.linalg.norm((np.asarray(a) - np.asarray(b)).ravel()) / max(np.linalg.norm(np.asarray(b).ravel()), 1e-300))

    res, ok = {}, True
    for mode, rmode in (("fractional", R.ABF_ONCE), ("coupled", R.EXACT)):
        ref = R.Reference(case.n, case.faces(), case.rho, case.mu, case.dt, [O.BC(b["type"], velocity=b["velocity"], pressure=b["pressure"]) for b in case.bcs])
        ref.set_state(*state)
        ns = W.make_ns(case, lib, mode, **tight)
        W.set_initial(ns, state)
        for _ in range(2):
            ref.step(mode=rmode)
            fb.NSStep(ns)
        a, g = ref.get_state(), fb.NSB200GetSolver(ns).get_state()
        fb.NSDestroy(ns)
        eU = float(np.sqrt(sum(np.sum((x - y) ** 2) for x, y in zip(g["U"], a["U"]))) / np.sqrt(sum(np.sum(y**2) for y in a["U"])))
        res[mode] = dict(v=rel(g["v"], a["v"]), U=eU, p=rel(g["p"], a["p"]), phalf=rel(g["phalf"], a["phalf"]))
        ok = ok and res[mode]["v"] <= 1e-10 and eU <= 1e-10 and res[mode]["p"] <= 1e-9
    return {"available": True, "ok": bool(ok), "against": "the reference's own NS sources (cartdiscret.c, cnlinear*.c, abfpc.c) compiled on a single-rank PETSc model with exact linear solves", "case": "sphere workload box and boundary set without markers, 8x5x5, 2 steps, rank 0's device", "runs": res}


def e2e_resident_loop(ns, solver, ksteps, barrier, clock):
    """The time loop the PETSc glue runs with an output every step (glue/nsb200.c, -ns_b200_sync_interval 1), with the
    download taken off the critical path: the state stays on the device, each step's result is staged into pinned host
    memory on the view stream (fluca_b200_stage_state) and collected after the NEXT step has been computed, so the
    device-to-host copy of step k runs behind the compute of step k + 1; the last copy is drained inside the timed region.
    Returns (seconds, d2h bytes per step, checksum of what the host read)."""
    import fluca_b200 as fb

    solver.stage_state()  # untimed: the first staging allocates the pinned buffers
    solver.staged_state(copy=False)
    barrier()
    t0 = clock()
    acc, nbytes = 0.0, 0
    for k in range(ksteps):
        fb.NSStep(ns)  # boundary planes that changed go up inside; the copy of step k - 1 is in flight meanwhile
        if k > 0:
            view = solver.staged_state(copy=False)  # result of step k - 1, now complete in host memory
            acc += float(view["p"].flat[0]) + float(view["v"].flat[-1])
        solver.stage_state()
    view = solver.staged_state(copy=False)
    acc += float(view["p"].flat[0]) + float(view["v"].flat[-1])
    nbytes = view["v"].nbytes + view["p"].nbytes + view["phalf"].nbytes + sum(u.nbytes for u in view["U"])
    barrier()
    return clock() - t0, nbytes, acc


def load_peaks():
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        peaks = {}
    peak = float(peaks.get("hbm_gbs", 6650.0))
    src = "measured copy bandwidth (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    return peak, src


def measure(ctx, lib, args, scaling, with_e2e):
    """Builds the workload on this run's ranks, warms up, times K steps device-resident, and (optionally) end to end."""
    import fluca_b200 as fb
    from fluca_b200 import workloads as W

    world, rank = ctx.world, ctx.rank
    n = default_n(args)
    nzg = n if scaling == "strong" else n * world
    if args.workload == "channel":
        nzg = n * world // 8 if scaling == "weak" else n  # config 5: 2048 x 1024 x 1024 on 8 GPUs = 128 z-planes per GPU
        nzg = max(nzg, 8 * world)
    big = (float(n) ** 2 * nzg / world) * (2 if args.workload == "channel" else 1) >= 100e6
    # config 5 (2.2 GB per field and GPU): the coupled mode fits with restart 1 (73 fields)
    restart = args.restart or ((1 if args.workload == "channel" else 3) if big else 10)
    case = make_case(args, n, nzg)
    note(f"measure {scaling}: building {case.n}")
    opts = {"ns_ksp_gmres_restart": restart, "ns_pc_abf_schur_ainv_type": args.schur_ainv, "ns_pc_abf_upper_ainv_type": args.upper_ainv}
    ns = W.make_ns(case, lib, args.mode, comm=ctx.comm(), **opts)
    s = fb.NSB200GetSolver(ns)
    cells_total = float(case.n[0]) * case.n[1] * case.n[2]
    cells_rank = float(case.n[0]) * case.n[1] * s.nzl
    set_initial_state(args, case, s)
    mk = markers_for(args, case, n, nzg)
    nmark = 0
    if mk is not None:
        fb.NSB200SetMarkers(ns, mk["X"], mk["Ud"], mk["dV"], 4)
        nmark = int(mk["dV"].shape[0])

    # ---- device-resident throughput ("value"): state lives in HBM, nothing crosses PCIe in the timed region
    note(f"measure {scaling}: warm-up")
    for _ in range(args.warmup):
        fb.NSStep(ns)
    ctx.barrier()
    note(f"measure {scaling}: timed region")
    sampler = ClockSampler(ctx.local)
    if rank == 0:
        sampler.start()
    # (1) the headline: K steps with NO per-launch instrumentation (event pairs around ~700 launches per step cost time, and
    #     the multigrid cycles replay as CUDA graphs only when nothing is recorded between their launches)
    s.kernel_timing(False)
    l0 = s.launch_count()
    stats = []
    start, stop = ctx.stream_timer(s)
    ctx.barrier()
    start()
    for _ in range(args.steps):
        fb.NSStep(ns)
        stats.append(fb.NSB200GetStats(ns))
    ms = stop()
    ctx.barrier()
    ms = ctx.max_over_ranks(ms)
    note(f"measure {scaling}: {ms / args.steps:.2f} ms per step")
    launches = s.launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    value = cells_total * args.steps / (ms * 1e-3) / 1e6
    # (2) the breakdown: a few more steps with a CUDA-event pair around every launch (class times, per-class rooflines)
    nb = max(1, min(args.steps, 5))
    s.kernel_times(reset=True)
    s.kernel_timing(True)
    bstats = []
    start, stop = ctx.stream_timer(s)
    ctx.barrier()
    start()
    for _ in range(nb):
        fb.NSStep(ns)
        bstats.append(fb.NSB200GetStats(ns))
    bms = ctx.max_over_ranks(stop())
    ctx.barrier()
    ktimes = s.kernel_times(reset=True)
    s.kernel_timing(False)

    # ---- rooflines (live CUDA-event pairs around every launch of the timed region; this rank's kernels)
    peak, peak_src = load_peaks()
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    except Exception:
        traffic = {}
    roof = None
    t_ma, cnt_ma = ktimes["momentum_apply"]
    if cnt_ma:
        # A_apply = R x(3) v0(3) U0(3) W y(3) = 96 B per cell.  Every other apply of a BiCGStab iteration also
        # reads the separate dot partner r^ (3 more fields, 120 B), which 8(d) books under the 15 vector passes; both are given.
        ach = 96.0 * cells_rank * cnt_ma / (t_ma * 1e-3) / 1e9
        tr = traffic.get("momentum_apply", {})
        roof = {"kernel": "k_tma_march<AApplyTile> + wall launches (momentum operator y = A x fused with <a, y>, <y, y>)", "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                "traffic": (tr["bytes_per_cell"] * cells_rank / 1e9 if "bytes_per_cell" in tr else None), "traffic_unit": "GB per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum per cell of the committed capture x cells of this launch)", "traffic_source": tr.get("capture"),
                "algorithmic_GB_per_launch": 96.0 * cells_rank / 1e9, "algorithmic_bytes_per_cell": 96.0, "launches": cnt_ma, "avg_ms": t_ma / cnt_ma, "share_of_step": t_ma / bms, "peak_source": peak_src,
                "achieved_counting_the_dot_partner": 108.0 * cells_rank * cnt_ma / (t_ma * 1e-3) / 1e9, "note": "achieved = SURVEY 8(d)'s 96 B/cell x cells of one launch / mean CUDA-event duration of one operator application in the timed region; the second figure adds the dot partner r^ that every other application reads (108 B/cell mean)"}
    # per class: the step model (fluca_b200_step_model_bytes_split) over the event time of the class
    split = {}
    for st in bstats:
        for k, v in s.model_bytes_split(st).items():
            split[k] = split.get(k, 0.0) + v
    classes = {}
    for k, (t, cnt) in ktimes.items():
        if cnt and t > 0:
            by = split.get(k, 0.0)
            classes[k] = {"ms_per_step": t / nb, "launch_groups_per_step": cnt / nb, "share_of_step": round(t / bms, 4), "model_GB_per_step": by / nb / 1e9, "achieved": (by / (t * 1e-3) / 1e9) if by else None, "frac": (by / (t * 1e-3) / 1e9 / peak) if by else None}
    t_solve = sum(ktimes[k][0] for k in ("poisson_apply", "poisson_vec", "mg_smooth", "mg_transfer"))
    b_solve = sum(split.get(k, 0.0) for k in ("poisson_apply", "poisson_vec", "mg_smooth", "mg_transfer"))
    its_p = int(sum(st.schur_its for st in bstats))
    poisson = None
    if t_solve > 0 and its_p > 0:
        ach = b_solve / (t_solve * 1e-3) / 1e9
        poisson = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "iterations": its_p, "bytes_per_iteration_per_cell": b_solve / its_p / cells_rank, "ms": t_solve, "krylov": "bicgstab+mg" if (args.workload == "sphere" or args.schur_ainv != "ID") else "pcg+mg"}
    model_bytes = sum(s.model_bytes(st) for st in stats)
    step_ach = model_bytes / (ms * 1e-3) / 1e9
    kernel_ms = sum(t for t, c in ktimes.values())

    res = {
        "value": value,
        "ms_per_step": ms / args.steps,
        "scaling": scaling,
        "cells": cells_total,
        "workload": workload_text(args, case, scaling, world, restart, nmark),
        "iterations_per_step": {"outer": [st.outer_its for st in stats], "momentum": [st.mom_its for st in stats], "schur": [st.schur_its for st in stats], "abf": [st.abf_applies for st in stats]},
        "roofline": roof,
        "poisson_solve": poisson,
        "step_roofline": {"algorithmic_GB_per_step": model_bytes / args.steps / 1e9 * world, "achieved": step_ach * world, "peak": peak * world, "unit": "GB/s", "frac": step_ach / peak, "peak_source": peak_src, "note": "SURVEY 8(d) step model bytes (this rank's cells) / step time / peak"},
        "class_rooflines": classes,
        "breakdown": {"steps": nb, "ms_per_step": bms / nb, "iterations_per_step": {"outer": [st.outer_its for st in bstats], "momentum": [st.mom_its for st in bstats], "schur": [st.schur_its for st in bstats]}, "note": "class_rooflines, roofline and poisson_solve come from these extra steps, run after the headline steps with a CUDA-event pair around every launch (eager multigrid cycles); the headline value is timed without any of that"},
        "time_outside_kernels_frac": max(0.0, 1.0 - kernel_ms / bms),
        "gpu_launches": int(launches),
        "clocks": clocks,
        "l2": f"inputs larger than L2 (each field {8 * cells_rank / 1e6:.0f} MB per GPU vs 126 MB L2; >50 fields streamed per step), no flush",
    }

    # ---- end to end through the public API with HOST buffers (pinned): H2D of the step's inputs + step + D2H of the result
    if with_e2e:
        st0 = s.get_state()
        host = {k: ctx.pinned(v.shape) for k, v in (("v", st0["v"]), ("p", st0["p"]), ("phalf", st0["phalf"]))}
        hostU = [ctx.pinned(u.shape) for u in st0["U"]]
        for k in host:
            host[k][...] = st0[k]
        for a, u in zip(hostU, st0["U"]):
            a[...] = u
        del st0
        h2d = sum(a.nbytes for a in host.values()) + sum(a.nbytes for a in hostU)
        d2h = h2d
        import ctypes as C

        def ptrs(arrs):
            p = (C.c_void_p * 3)()
            for i, a in enumerate(arrs):
                p[i] = a.ctypes.data
            return p

        ksteps = max(1, min(args.steps, 2))
        ctx.barrier()
        t0 = time.perf_counter()
        for _ in range(ksteps):
            fb._lib.check(lib, lib.fluca_b200_set_state(s._h, host["v"].ctypes.data, ptrs(hostU), host["p"].ctypes.data, host["phalf"].ctypes.data))
            fb.NSStep(ns)
            fb._lib.check(lib, lib.fluca_b200_get_state(s._h, host["v"].ctypes.data, ptrs(hostU), host["p"].ctypes.data, host["phalf"].ctypes.data))
        ctx.barrier()
        te = ctx.max_over_ranks(time.perf_counter() - t0)
        res["e2e"] = {"value": cells_total * ksteps / te / 1e6, "unit": "Mcell-updates/s", "h2d_bytes_per_step": int(h2d) * world, "d2h_bytes_per_step": int(d2h) * world, "steps": ksteps, "note": "NSStep through the NS API with pinned host buffers: set_state (H2D) + boundary planes + step + get_state (D2H) inside the timed region; bytes summed over ranks"}
        del host, hostU
        # ---- the same loop as the glue runs it: resident state, result staged to pinned memory behind the next step's compute
        if world == 1:  # N=1 only: an exception on one rank must not strand the others in a collective
            try:
                ko = max(2, args.steps)
                to, d2h_o, _ = e2e_resident_loop(ns, s, ko, ctx.barrier, time.perf_counter)
                res["e2e_resident"] = {"value": cells_total * ko / to / 1e6, "unit": "Mcell-updates/s", "ms_per_step": to / ko * 1e3, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": int(d2h_o), "steps": ko, "note": "NOT the contract's e2e (above): state resident on the device as in glue/nsb200.c, constant boundary planes (nothing to re-send), every step's full result copied to pinned host memory by fluca_b200_stage_state on a second stream and read by the host after the next step; the final copy is drained inside the timed region"}
            except Exception as exc:  # an extra: it must never cost the bench line
                res["e2e_resident"] = {"error": repr(exc)[:200]}
    fb.NSDestroy(ns)
    return res


def run_b200(args, ctx=None, lib=None):
    import fluca_b200 as fb

    ctx = ctx or Ctx()
    lib = lib or fb._lib.load()  # CUDA library or a loud failure: there is no fallback (tests pass the host-emulation double)
    world, rank = ctx.world, ctx.rank
    parity = None
    if not args.no_parity:
        try:
            parity = parity_selfcheck(ctx, lib)
        except Exception as exc:  # reported, never fatal for the timing that follows; every rank leaves the same way
            parity = {"ok": False, "error": repr(exc)[:300]}
        if rank == 0 and isinstance(parity, dict):
            try:
                parity["reference_sources"] = reference_selfcheck(lib)
            except Exception as exc:  # an extra: never fatal, never a collective
                parity["reference_sources"] = {"available": True, "ok": False, "error": repr(exc)[:300]}
    ctx.barrier()
    if args.workload == "channel":
        order = ["weak"]  # config 5 is sized per GPU (268 M cells each); the full grid only exists at 8 GPUs
    elif world == 1:
        order = ["strong"]
    else:
        order = {"both": ["strong", "weak"], "strong": ["strong"], "weak": ["weak"]}[args.scaling]
    runs = {}
    for i, sc in enumerate(order):
        runs[sc] = measure(ctx, lib, args, sc, with_e2e=(i == 0 and not args.no_e2e))
        ctx.barrier()
    head = runs[order[0]]
    line = {
        "metric": "Mcell-updates/s per NS step",
        "value": head["value"],
        "unit": "Mcell-updates/s",
        "n_gpus": world,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": head["ms_per_step"],
        "higher_is_better": True,
        "scaling": order[0] if world > 1 else ("weak" if args.workload == "channel" else "strong"),
        "vs_baseline": None,
        "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": head["workload"], "mode": args.mode, "abf_ainv": {"schur": args.schur_ainv, "upper": args.upper_ainv}, "l2": head["l2"], "iterations_per_step": head["iterations_per_step"]},
        "roofline": head["roofline"],
        "poisson_solve": head["poisson_solve"],
        "step_roofline": head["step_roofline"],
        "class_rooflines": head["class_rooflines"],
        "breakdown": head["breakdown"],
        "time_outside_kernels_frac": head["time_outside_kernels_frac"],
        "gpu_launches": head["gpu_launches"],
        "clocks": head["clocks"],
        "parity": parity,
    }
    for k in ("e2e", "e2e_resident"):
        if k in head:
            line[k] = head[k]
    if "weak" in runs and order[0] != "weak":
        w = runs["weak"]
        line["weak"] = {k: w[k] for k in ("value", "ms_per_step", "cells", "workload", "iterations_per_step", "step_roofline", "class_rooflines", "time_outside_kernels_frac", "gpu_launches")}
        line["weak"]["unit"] = "Mcell-updates/s"

    # ---- CPU baseline on the box's host cores (rank 0, N=1 only), bounded sample
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        r = cpu_sample(args, args.cpu_n, 3, 1, args.mode)
        line["cpu_baseline"] = {"value": r["value"], "unit": "Mcell-updates/s", "cores": r["threads"], "kind": "port", "sample": f"{args.cpu_n}^3 sample of the workload (same BCs, dt=0.5h, mode={args.mode}, tolerances 1e-5), 1 warm-up + 3 timed steps in {r['seconds']:.1f} s on {r['threads']} OpenMP threads", "reference_build": reference_build_probe(), "reference_sources": r["reference_sources"]}
    if rank == 0:
        print_json(json.dumps(line))
    if world > 1:
        ctx.dist.destroy_process_group()
    return line


def main():
    import faulthandler

    faulthandler.enable()
    faulthandler.dump_traceback_later(240, repeat=True, file=sys.stderr)  # a stuck run says where, every 4 minutes
    args = parse()
    # stdout carries exactly ONE JSON line: libraries that print to fd 1 (NCCL prints its version there)
    # are sent to stderr for the duration of the run
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    out = os.fdopen(real_stdout, "w")

    def emit(text):
        out.write(text + "\n")
        out.flush()

    global print_json
    print_json = emit
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
