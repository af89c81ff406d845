"""Worker of the multi-rank CPU test of bench.py's GPU arm: one process per rank, gloo, the host-emulation test double with the
callback communicator of tests/multirank_worker.py.  Runs bench.run_b200 (parity self-check + strong + weak measurement) on a
tiny grid so that mismatched collectives between ranks show up here and not on an 8-GPU box."""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch.distributed as dist

import bench
import fluca_b200 as fb
from tests import parity
from tests.multirank_worker import make_comm_factory


class GlooCtx:
    def __init__(self):
        self.rank, self.world, self.local, self.dev = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), 0, None
        self.dist = dist

    def barrier(self):
        dist.barrier()

    def max_over_ranks(self, x):
        import torch

        t = torch.tensor([x], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def comm(self):
        return dict(rank=self.rank, nranks=self.world, make_comm=make_comm_factory(self.rank, self.world))

    def stream_timer(self, solver):
        t = [0.0]

        def start():
            t[0] = time.perf_counter()

        return start, lambda: 1e3 * (time.perf_counter() - t[0])

    def pinned(self, shape):
        return np.empty(shape)

    def gather_z(self, arrs, axis):
        out = [None] * self.world if self.rank == 0 else None
        dist.gather_object(arrs, out, dst=0)
        return np.concatenate(out, axis=axis) if self.rank == 0 else None


def main():
    out_path = sys.argv[1]
    dist.init_process_group("gloo", rank=int(os.environ["RANK"]), world_size=int(os.environ["WORLD_SIZE"]))
    ctx = GlooCtx()
    lib = parity.hostemu_library() if ctx.rank == 0 else None
    dist.barrier()
    lib = lib or fb._lib.load(parity.HOSTEMU)
    lines = []
    bench.print_json = lines.append
    args = argparse.Namespace(gpus=ctx.world, steps=2, warmup=1, impl="b200", workload=os.environ.get("BENCH_WORKLOAD", "sphere"), n=int(os.environ.get("BENCH_N", "16")), markers=300, scaling="both", strong=False, mode="coupled", restart=0, schur_ainv="ID", upper_ainv="ID", cpu_n=8, no_cpu_baseline=True, no_e2e=False, no_parity=False)

    bench.run_b200(args, ctx=ctx, lib=lib)
    if ctx.rank == 0:
        open(out_path, "w").write(lines[0])


if __name__ == "__main__":
    main()
