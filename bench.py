#!/usr/bin/env python
"""bench.py -- Mcell-updates/s of one Navier-Stokes time step (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W                (our arm: CUDA kernels on N B200s)
  python bench.py --impl reference --gpus N --steps K --warmup W   (reference arm: CPU restatement)

One "step" = one full NS time step (NSStep) of the configured workload.  Prints ONE JSON line.
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


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="sphere", choices=["sphere", "cavity"], help="sphere: BASELINE config 4 (3-D IBM sphere Re=300, 512^3, 100k markers), the configuration the metric is quoted on; cavity: config 3 (256^3 lid-driven cavity, no IBM)")
    ap.add_argument("--n", type=int, default=0, help="cells per direction per GPU (default: 512 sphere, 256 cavity)")
    ap.add_argument("--markers", type=int, default=100000)
    ap.add_argument("--strong", action="store_true", help="strong scaling: the n^3 grid is split over the GPUs (default: weak, n^3 per GPU)")
    ap.add_argument("--mode", default="coupled", choices=["coupled", "fractional"])
    ap.add_argument("--restart", type=int, default=0, help="outer GMRES restart (memory: (restart+1) x 7 fields; default 3 at 512^3, 10 below; the flexible form also keeps restart x 7 fields of preconditioned vectors)")
    ap.add_argument("--schur-ainv", default="ID", choices=["ID", "DIAG", "ROWSUM"], help="-ns_pc_abf_schur_ainv_type (abfpc.c:246); the headline numbers use the reference default ID")
    ap.add_argument("--upper-ainv", default="ID", choices=["ID", "DIAG", "ROWSUM"], help="-ns_pc_abf_upper_ainv_type (abfpc.c:247)")
    ap.add_argument("--cpu-n", type=int, default=64, help="cells per direction of the bounded CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------- workload
def cavity_case(n, nz, Re=400.0):
    """BASELINE config 3: 3-D lid-driven cavity Re=400 on the unit cube, uniform h = 1/n, dt = 0.5 h,
    zero initial state.  For the weak-scaling runs the box is extended in z (nz = n * N
    cells, length N) so that every GPU keeps an n^3 slab."""
    from tests import cases

    c = cases.cavity3d_full(n=(n, n, nz), Re=Re, dt=0.5 / n)
    c.hi = (1.0, 1.0, float(nz) / n)
    return c


def sphere_case(n, nz, Re=300.0):
    """BASELINE config 4: flow past a sphere (D = 1 at the origin, U_inf = 1) by the immersed-boundary coupling on
    [-4,12] x [-8,8]^2 with n^3 cells (h = 16/n; 512^3 -> h = 1/32), inflow LEFT, pressure outlet RIGHT (p = 0), symmetry on
    the four side boundaries, dt = 0.5 h (CFL 0.5), initial state uniform U_inf.  Weak scaling extends the
    box in z (nz = n * N cells) so that every GPU keeps an n^3 slab; the sphere stays at the origin."""
    from tests import cases

    c = cases.channel3d(n=(n, n, nz), Re=Re, dt=0.5 * 16.0 / n)
    c.lo, c.hi = (-4.0, -8.0, -8.0), (12.0, 8.0, -8.0 + 16.0 * nz / n)
    for b in c.bcs:  # constant boundary data, evaluated once per plane
        for k in ("velocity", "pressure"):
            if b[k] is not None:
                b[k] = cases._const(1.0, 0.0, 0.0) if k == "velocity" else cases._constp(0.0)
    return c


def make_case(args, n, nz):
    return sphere_case(n, nz) if args.workload == "sphere" else cavity_case(n, nz)


def markers_for(args, n):
    from tests import cases

    if args.workload != "sphere":
        return None
    return cases.sphere_markers((0.0, 0.0, 0.0), 1.0, args.markers, 16.0 / n)


def uniform_inflow_state(case):
    cell, face = case.shapes()
    v = np.zeros((3,) + cell)
    v[0] = 1.0
    U = [np.zeros(s) for s in face]
    U[0][...] = 1.0
    return v, U, np.zeros(cell)


def workload_text(args, n, nzg, restart):
    if args.workload == "sphere":
        return (f"BASELINE config 4: 3-D flow past a sphere by IBM, Re=300, {n}x{n}x{nzg} cells ({'z-slabs of the one grid' if getattr(args, 'strong', False) else f'{n}^3 per GPU, z-slabs'}), h=16/{n}, {args.markers} Fibonacci markers, 4-point delta, "
                f"inflow/pressure-outlet/symmetry, dt=0.5h, uniform initial state, NS type b200 mode={args.mode}, reference default tolerances (outer/momentum/Schur rtol 1e-5; inner tolerances relaxed by the inexact-Krylov rule as the outer residual drops, DESIGN.md 5), flexible GMRES restart {restart}")
    return (f"BASELINE config 3: 3-D lid-driven cavity Re=400, {n}x{n}x{nzg} cells ({n}^3 per GPU, z-slabs), dt=0.5h, zero initial state, NS type b200 mode={args.mode}, "
            f"reference default tolerances (outer/momentum/Schur rtol 1e-5; inner tolerances relaxed by the inexact-Krylov rule, DESIGN.md 5), flexible GMRES restart {restart}")


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
def poisson_solve_rate(ktimes, schur_its, cells_rank, outlet, variant, peak):
    """Algorithmic HBM rate of the pressure solves of the timed region: iterations x bytes per iteration and cell (3-D: PCG +
    V(2,2) cycle = 227 B; BiCGStab with a pressure outlet or a DIAG / ROWSUM Schur complement = two applies + two V-cycles,
    2 x 227 resp. 2 x (227 - 16 + 104) B -- the model of fluca_b200_step_model_bytes) over the event-timed duration of the
    four kernel classes that make up the solve.  None if nothing was timed."""
    try:
        t_ms = sum(float(ktimes[k][0]) for k in ("poisson_apply", "poisson_vec", "mg_smooth", "mg_transfer") if k in ktimes)
        its = int(sum(schur_its))
        if t_ms <= 0.0 or its <= 0:
            return None
        per_iter = 2.0 * (227.0 - 16.0 + 104.0) if variant else (2.0 * 227.0 if outlet else 227.0)
        ach = its * per_iter * cells_rank / (t_ms * 1e-3) / 1e9
        return {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "iterations": its, "bytes_per_iteration_per_cell": per_iter, "ms": t_ms, "krylov": "bicgstab+mg" if (variant or outlet) else "pcg+mg"}
    except Exception:  # a reporting extra must never cost the bench line
        return None


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


def cpu_sample(args, n, steps, warmup, mode, threads):
    """The CPU restatement of the reference algorithm (oracle, kind "port": assembled CSR operators,
    GMRES(30) + block-Jacobi ILU(0), reference default tolerances) on an n^3 sample of the workload, all host threads."""
    from oracle import oracle as O
    from tests import cases

    os.environ.setdefault("OMP_NUM_THREADS", str(threads))
    case = make_case(args, n, n)
    orc = cases.make_oracle_fast(case)
    if args.workload == "sphere":
        orc.set_state(*uniform_inflow_state(case))
        mk = cases.sphere_markers((0.0, 0.0, 0.0), 1.0, max(64, int(args.markers * (n / 512.0) ** 2)), 16.0 / n)
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
    return dict(value=cells * steps / dt / 1e6, seconds=dt, steps=steps, n=n, outer=[i.outer_its for i in infos], mom=[i.mom_its for i in infos], schur=[i.schur_its for i in infos])


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    n = args.cpu_n
    r = cpu_sample(args, n, args.steps, min(args.warmup, 1), args.mode, threads)
    line = {
        "impl": "reference",
        "metric": "Mcell-updates/s per NS step",
        "value": r["value"],
        "unit": "Mcell-updates/s",
        "n_gpus": args.gpus,
        "steps": args.steps,
        "warmup": min(args.warmup, 1),
        "ms_per_step": 1e3 * r["seconds"] / r["steps"],
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": workload_text(args, args.n or (512 if args.workload == "sphere" else 256), (args.n or (512 if args.workload == "sphere" else 256)) * args.gpus, args.restart or 3) + f"; CPU arm: bounded sample {n}^3 of the same case (markers scaled with the surface cell count)", "mode": args.mode},
        "cpu_baseline": {"value": r["value"], "unit": "Mcell-updates/s", "cores": threads, "kind": "port", "sample": f"{n}^3 sample of the workload, {r['steps']} steps, outer its {r['outer']}, momentum its {r['mom']}, Schur its {r['schur']}; the reference (PETSc) cannot be built in this image, this is the repo's C restatement (oracle/)"},
        "e2e": {"value": r["value"], "unit": "Mcell-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print_json(json.dumps(line))


# ---------------------------------------------------------------------------------------------- GPU arm
def run_b200(args):
    import torch
    import torch.distributed as dist

    import fluca_b200 as fb
    from tests import parity

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    lib = fb._lib.load()  # CUDA library or a loud failure: there is no fallback
    comm = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
        # rank 0 makes the NCCL unique id of the solver's own communicator and ships it with torch.distributed
        if rank == 0:
            uid = np.frombuffer(fb.Comm.unique_id(lib), dtype=np.uint8).copy()
            tid = torch.from_numpy(uid).to(dev)
        else:
            tid = torch.zeros(128, dtype=torch.uint8, device=dev)
        dist.broadcast(tid, 0)
        uid_bytes = tid.cpu().numpy().tobytes()
        comm = dict(rank=rank, nranks=world, make_comm=lambda L: fb.Comm.nccl(L, uid_bytes, rank, world))

    n = args.n or (512 if args.workload == "sphere" else 256)
    nzg = n if args.strong else n * world  # weak scaling (default): an n^3 slab per GPU; strong: n / N planes per GPU
    restart = args.restart or (3 if n >= 512 else 10)
    case = make_case(args, n, nzg)
    opts = {"ns_ksp_gmres_restart": restart, "ns_pc_abf_schur_ainv_type": args.schur_ainv, "ns_pc_abf_upper_ainv_type": args.upper_ainv}
    ns = parity.make_ns(case, lib, args.mode, comm=comm, **opts)
    s = fb.NSB200GetSolver(ns)
    cells_total = float(n) * n * nzg
    if args.workload == "sphere":
        v0_, U0_, p0_ = uniform_inflow_state(case)
        k0_, nzl_ = s.k0, s.nzl
        s.set_state(v=v0_[:, k0_ : k0_ + nzl_], U=[U0_[0][k0_ : k0_ + nzl_], U0_[1][k0_ : k0_ + nzl_], U0_[2][k0_ : k0_ + nzl_ + (1 if s.last_z else 0)]], p=p0_[k0_ : k0_ + nzl_], phalf=p0_[k0_ : k0_ + nzl_])
        del v0_, U0_, p0_
        mk = markers_for(args, n)
        fb.NSB200SetMarkers(ns, mk["X"], mk["Ud"], mk["dV"], 4)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput ("value"): state lives in HBM, nothing crosses PCIe in the timed region
    stream = torch.cuda.ExternalStream(s.stream(), device=dev)
    for _ in range(args.warmup):
        fb.NSStep(ns)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    s.kernel_times(reset=True)
    s.kernel_timing(True)
    l0 = s.launch_count()
    stats = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    with torch.cuda.stream(stream):
        e0.record(stream)
        for _ in range(args.steps):
            fb.NSStep(ns)
            stats.append(fb.NSB200GetStats(ns))
        e1.record(stream)
    e1.synchronize()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = s.launch_count() - l0
    ktimes = s.kernel_times(reset=True)
    s.kernel_timing(False)
    clocks = sampler.stop() if rank == 0 else None
    tms = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms = float(tms.item())
    value = cells_total * args.steps / (ms * 1e-3) / 1e6

    # ---- roofline of the dominant kernel class (live CUDA-event pairs around every launch in the timed region)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    cells_rank = float(n) * n * (nzg / world)
    d3 = True
    # algorithmic bytes per launch and cell (DESIGN.md "Kernels"): A apply + 2 dots reads x(3) v0(3) U0(3) a(3) writes y(3)
    # momentum apply fused with two dots: reads x(3) v0(3) U0(3) (+ rhat(3) in the first of the two applies of a
    # BiCGStab iteration), writes y(3): (120 + 96) / 2 = 108 B per cell and launch on average
    per_launch = {"momentum_apply": 108.0, "poisson_apply": 16.0}
    if args.schur_ainv != "ID":
        # DIAG / ROWSUM Schur complement = two launches per apply: w = (1 - a1) G0 p reads p, a1(3) writes w(3) = 56 B;
        # out = P p + vol D T w reads p, w(3), a writes out = 48 B; mean 52 B per cell and launch
        per_launch["poisson_apply"] = 52.0
    shares = {k: v[0] / ms for k, v in ktimes.items() if v[1] > 0}
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    except Exception:
        traffic = {}
    roof = None
    for name in ("momentum_apply", "poisson_apply"):
        t, cnt = ktimes[name]
        if cnt:
            ach = per_launch[name] * cells_rank * cnt / (t * 1e-3) / 1e9
            r = {"kernel": name, "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": (traffic[name]["bytes_per_cell"] * cells_rank / 1e9 if name in traffic else None), "traffic_unit": "GB per launch (ncu dram bytes per cell of the committed capture x cells of this launch)", "traffic_source": traffic.get(name, {}).get("capture"), "algorithmic_GB_per_launch": per_launch[name] * cells_rank / 1e9, "launches": cnt, "avg_ms": t / cnt, "share_of_step": t / ms, "peak_source": peak_src}
            if roof is None or t > roof["_t"]:
                roof = dict(r, _t=t)
    if roof:
        roof.pop("_t")
    # second half of BASELINE.json's metric, "Poisson solve HBM GB/s vs peak" (SURVEY.md 8d): n_cg x 227 B x cells / t_poisson
    poisson = poisson_solve_rate(ktimes, [st.schur_its for st in stats], cells_rank, outlet=(args.workload == "sphere"), variant=(args.schur_ainv != "ID"), peak=peak)
    # whole-step model: algorithmic bytes of all kernels / elapsed
    model_bytes = sum(s.model_bytes(st) for st in stats)
    step_ach = model_bytes / (ms * 1e-3) / 1e9

    line = {
        "metric": "Mcell-updates/s per NS step",
        "value": value,
        "unit": "Mcell-updates/s",
        "n_gpus": world,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": ms / args.steps,
        "higher_is_better": True,
        "scaling": "strong" if args.strong else "weak",
        "vs_baseline": None,
        "dtype": "f64",
        "data": "synthetic",
        "config": {
            "workload": workload_text(args, n, nzg, restart),
            "mode": args.mode,
            "abf_ainv": {"schur": args.schur_ainv, "upper": args.upper_ainv},
            "l2": f"inputs larger than L2 (each field {8 * n**3 / 1e6:.0f} MB vs 126 MB L2; >50 fields streamed per step), no flush",
            "iterations_per_step": {"outer": [st.outer_its for st in stats], "momentum": [st.mom_its for st in stats], "schur": [st.schur_its for st in stats], "abf": [st.abf_applies for st in stats]},
        },
        "roofline": roof,
        "poisson_solve": poisson,
        "step_roofline": {"algorithmic_GB_per_step": model_bytes / args.steps / 1e9, "achieved": step_ach * world, "peak": peak * world, "unit": "GB/s", "frac": step_ach / peak, "peak_source": peak_src},
        "kernel_shares": {k: round(v, 4) for k, v in shares.items()},
        "gpu_launches": int(launches),
        "clocks": clocks,
    }

    # ---- end to end through the public API with HOST buffers (pinned): H2D of the step's inputs + step + D2H of the result
    if not args.no_e2e:
        st0 = s.get_state()
        host = {k: torch.empty(v.shape, dtype=torch.float64).pin_memory().numpy() for k, v in (("v", st0["v"]), ("p", st0["p"]), ("phalf", st0["phalf"]))}
        hostU = [torch.empty(u.shape, dtype=torch.float64).pin_memory().numpy() for u in st0["U"]]
        for k in host:
            host[k][...] = st0[k]
        for a, u in zip(hostU, st0["U"]):
            a[...] = u
        h2d = sum(a.nbytes for a in host.values()) + sum(a.nbytes for a in hostU)
        d2h = h2d
        import ctypes as C

        def ptrs(arrs):
            p = (C.c_void_p * 3)()
            for i, a in enumerate(arrs):
                p[i] = a.ctypes.data
            return p

        ksteps = max(1, min(args.steps, 2))
        barrier()
        t0 = time.perf_counter()
        for _ in range(ksteps):
            fb._lib.check(lib, lib.fluca_b200_set_state(s._h, host["v"].ctypes.data, ptrs(hostU), host["p"].ctypes.data, host["phalf"].ctypes.data))
            fb.NSStep(ns)
            fb._lib.check(lib, lib.fluca_b200_get_state(s._h, host["v"].ctypes.data, ptrs(hostU), host["p"].ctypes.data, host["phalf"].ctypes.data))
        barrier()
        te = time.perf_counter() - t0
        tt = torch.tensor([te], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        te = float(tt.item())
        line["e2e"] = {"value": cells_total * ksteps / te / 1e6, "unit": "Mcell-updates/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "steps": ksteps, "note": "NSStep through the NS API with pinned host buffers: set_state (H2D) + boundary planes + step + get_state (D2H) inside the timed region"}

        # ---- the same loop as the glue runs it: resident state, result staged to pinned memory behind the next step's compute
        try:  # N=1 only: an exception on one rank must not strand the others in a collective
            if world > 1:
                raise StopIteration
            ko = max(2, args.steps)
            to, d2h_o, _ = e2e_resident_loop(ns, s, ko, barrier, time.perf_counter)
            line["e2e_resident"] = {"value": cells_total * ko / to / 1e6, "unit": "Mcell-updates/s", "ms_per_step": to / ko * 1e3, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": int(d2h_o), "steps": ko, "note": "NOT the contract's e2e (above): state resident on the device as in glue/nsb200.c, constant boundary planes (nothing to re-send), every step's full result copied to pinned host memory by fluca_b200_stage_state on a second stream and read by the host after the next step; the final copy is drained inside the timed region"}
        except StopIteration:
            pass
        except Exception as exc:  # an extra: it must never cost the bench line
            line["e2e_resident"] = {"error": repr(exc)[:200]}

    # ---- CPU baseline on the box's host cores (rank 0, N=1 only), bounded sample
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        r = cpu_sample(args, args.cpu_n, 1, 0, args.mode, threads)
        line["cpu_baseline"] = {"value": r["value"], "unit": "Mcell-updates/s", "cores": threads, "kind": "port", "sample": f"{args.cpu_n}^3 sample of the workload (same BCs, dt=0.5h, mode={args.mode}, tolerances 1e-5), 1 step in {r['seconds']:.1f} s: the reference needs PETSc (absent) so this is the repo's C restatement"}
    if rank == 0:
        print_json(json.dumps(line))
    fb.NSDestroy(ns)
    if world > 1:
        dist.destroy_process_group()


def main():
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
