"""Worker of the world_size>1 CPU test: one process per rank, gloo backend, the solver library's
host-emulation test double with a callback communicator (halo exchange = dist.isend/irecv of ghost
planes, Krylov reductions = dist.all_reduce).  Exercises the slab partition, the halo protocol and
the multi-rank multigrid / Krylov orchestration of the product sources without a GPU."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch
import torch.distributed as dist

import fluca_b200 as fb
from tests import cases, parity


def make_comm_factory(rank, nranks):
    def factory(L):
        def view(ptr, count):
            return np.ctypeslib.as_array(ptr, shape=(count,))

        def halo(ctx, send_down, recv_down, send_up, recv_up, count, periodic):
            down = rank - 1 if rank > 0 else (nranks - 1 if periodic else None)
            up = rank + 1 if rank < nranks - 1 else (0 if periodic else None)
            reqs, recvs = [], []
            if down is not None:
                reqs.append(dist.isend(torch.from_numpy(view(send_down, count).copy()), down, tag=0))
                t = torch.empty(count, dtype=torch.float64)
                reqs.append(dist.irecv(t, down, tag=1))
                recvs.append((recv_down, t))
            if up is not None:
                reqs.append(dist.isend(torch.from_numpy(view(send_up, count).copy()), up, tag=1))
                t = torch.empty(count, dtype=torch.float64)
                reqs.append(dist.irecv(t, up, tag=0))
                recvs.append((recv_up, t))
            for r in reqs:
                r.wait()
            for ptr, t in recvs:
                view(ptr, count)[:] = t.numpy()
            return 0

        def allsum(ctx, vals, n):
            a = view(vals, n)
            t = torch.from_numpy(a.copy())
            dist.all_reduce(t)
            a[:] = t.numpy()
            return 0

        def allgather(ctx, send, recv, count):
            t = torch.from_numpy(view(send, count).copy())
            out = [torch.empty(count, dtype=torch.float64) for _ in range(nranks)]
            dist.all_gather(out, t)
            view(recv, count * nranks)[:] = torch.cat(out).numpy()
            return 0

        return fb.Comm.callbacks(L, rank, nranks, halo, allsum, allgather)

    return factory


def random_case3d(seed):
    """A random 3-D mesh / boundary set with at least 3 planes per slab for up to 3 ranks (tests/test_multirank_gloo.py)."""
    import math

    from fluca_b200.workloads import BC_PERIODIC, BC_PRESSURE_OUTLET, BC_SYMMETRY, BC_VELOCITY, Case

    rng = np.random.default_rng(seed)
    n = (int(rng.integers(4, 9)), int(rng.integers(4, 8)), int(rng.integers(9, 13)))
    lo = tuple(float(rng.uniform(-2, 0)) for _ in range(3))
    hi = tuple(x + float(rng.uniform(1, 4)) for x in lo)
    a = [float(rng.uniform(-1, 1)) for _ in range(8)]
    plan = []
    for _ in range(3):
        kind = rng.choice(["per", "walls", "mixed"], p=[0.25, 0.25, 0.5])
        plan.append([BC_PERIODIC] * 2 if kind == "per" else [int(rng.choice([BC_VELOCITY, BC_PRESSURE_OUTLET, BC_SYMMETRY], p=[0.5, 0.25, 0.25])) if kind == "mixed" else BC_VELOCITY for _ in range(2)])
    free = any(t == BC_PRESSURE_OUTLET for p in plan for t in p)

    def make_vel(dn):
        def vel(dim, t, x):
            return tuple(((a[c] + a[c + 3] * math.sin(1.3 * t + 0.7 * x[(c + 1) % dim])) if (free or c != dn) else 0.0) for c in range(dim))

        return vel

    def prs(dim, t, x):
        return a[6] * (1 + 0.3 * math.cos(2 * t)) * (1 + 0.2 * x[0] - 0.1 * x[2])

    bcs = [dict(type=plan[d][s], velocity=make_vel(d) if plan[d][s] == BC_VELOCITY else None, pressure=prs if plan[d][s] == BC_PRESSURE_OUTLET else None) for d in range(3) for s in range(2)]
    h = min((hi[d] - lo[d]) / n[d] for d in range(3))
    case = Case("random3d", n, lo, hi, float(rng.uniform(0.5, 2)), float(rng.uniform(0.01, 0.2)), float(rng.uniform(0.1, 0.3)) * h, bcs)
    case.stretch = float(rng.choice([0.0, 0.2]))
    return case


def main():
    case_name, mode, out_path = sys.argv[1], sys.argv[2], sys.argv[3]
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    backend = os.environ.get("FLUCA_WORKER_BACKEND", "gloo")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    if backend == "nccl":
        # the real thing: CUDA product library, one GPU per rank, the solver's own NCCL communicator
        torch.cuda.set_device(rank)
        lib = fb._lib.load()
        box = [fb.Comm.unique_id(lib) if rank == 0 else None]
        dist.broadcast_object_list(box, 0)
        uid = box[0]
    else:
        lib = parity.hostemu_library() if rank == 0 else None
        dist.barrier()
        lib = lib or fb._lib.load(parity.HOSTEMU)
    case = {
        "channel3d": lambda: cases.channel3d(n=(8, 6, 8), pout=0.2, dt=0.05),
        "z_outlet": lambda: cases.channel3d_z(n=(6, 6, 8), pout=0.2, dt=0.05),
        "z_outlet9": lambda: cases.channel3d_z(n=(6, 5, 9), pout=0.2, dt=0.05),
        "cavity3d": lambda: cases.cavity3d_full(n=(8, 8, 8)),
        "periodic_z": lambda: cases.channel3d(n=(8, 6, 8), periodic_z=True, dt=0.05),
        "uneven": lambda: cases.cavity3d_full(n=(8, 6, 7)),
        "three": lambda: cases.cavity3d_full(n=(8, 6, 9)),
        "cavity32": lambda: cases.cavity3d_full(n=(32, 32, 32)),
        "sphere_ibm": lambda: cases.channel3d(n=(12, 8, 8), pout=0.1, dt=0.05),
        "sphere_ibm_tma": lambda: cases.channel3d(n=(40, 16, 16), pout=0.1, dt=0.02),
        "sphere_ibm_periodic": lambda: cases.channel3d(n=(12, 8, 12), periodic_z=True, dt=0.05),
        "channel5": lambda: cases.channel_bench_case((8, 6, 8), periodic_z=True),  # BASELINE config 5's boundary set: periodic x and z, walls in y
    }.get(case_name, lambda: random_case3d(int(case_name.split(":")[1])))()
    markers = None
    if case_name.startswith("sphere_ibm"):  # the body straddles the slab interface: gather and scatter both cross it
        markers = cases.sphere_markers((0.1, 0.0, 0.05), 1.2, 120, 4.0 / case.n[1])
        if case_name == "sphere_ibm_periodic":  # near the periodic z boundary: the support wraps from the last slab into the first
            markers = cases.sphere_markers((0.1, 0.0, 1.7), 1.0, 150, 4.0 / case.n[1])
    if backend == "nccl":
        comm = dict(rank=rank, nranks=world, make_comm=lambda L: fb.Comm.nccl(L, uid, rank, world))
    else:
        comm = dict(rank=rank, nranks=world, make_comm=make_comm_factory(rank, world))
    ainv = [int(a) for a in os.environ.get("FLUCA_WORKER_AINV", "0,0").split(",")]  # PCABFAinvType of (Schur, upper)
    ns = parity.make_ns(case, lib, mode, comm=comm, ns_pc_abf_schur_ainv_type=parity.AINV_OPTION[ainv[0]], ns_pc_abf_upper_ainv_type=parity.AINV_OPTION[ainv[1]], **parity.TIGHT)
    s = fb.NSB200GetSolver(ns)
    v, U, p = case.initial_state(seed=31)
    k0, nzl = s.k0, s.nzl
    sl = slice(k0, k0 + nzl)
    Uz = U[2][k0 : k0 + nzl + (1 if s.last_z else 0)]
    s.set_state(v=v[:, sl], U=[U[0][sl], U[1][sl], Uz], p=p[sl])
    if markers is not None:
        fb.NSB200SetMarkers(ns, markers["X"], markers["Ud"], markers["dV"], 4, iterations=2)
    its = []
    for _ in range(2):
        fb.NSStep(ns)
        st = fb.NSB200GetStats(ns)
        its.append((st.outer_its, st.mom_its, st.schur_its))
    loc = s.get_state()
    # gather slabs on rank 0
    forces = fb.NSB200GetMarkerForces(ns) if markers is not None else None  # collective: every rank calls it
    infos = [None] * world
    if markers is not None:
        dist.all_gather_object(infos, s.ibm_info())
    parts = [None] * world
    dist.all_gather_object(parts, dict(k0=k0, nzl=nzl, v=loc["v"], U=loc["U"], p=loc["p"], phalf=loc["phalf"]))
    if rank == 0:
        parts.sort(key=lambda d: d["k0"])
        gv = np.concatenate([d["v"] for d in parts], axis=1)
        gp = np.concatenate([d["p"] for d in parts], axis=0)
        gph = np.concatenate([d["phalf"] for d in parts], axis=0)
        gU = [np.concatenate([d["U"][a] for d in parts], axis=0) for a in range(3)]
        extra = {}
        if markers is not None:
            F, Um = forces
            extra = dict(F=F, Um=Um)
            extra["ibm_info"] = np.array([[a, b, c, int(d)] for a, b, c, d in infos])
        np.savez(out_path, v=gv, p=gp, phalf=gph, U0=gU[0], U1=gU[1], U2=gU[2], its=np.array(its), **extra)
    fb.NSDestroy(ns)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
