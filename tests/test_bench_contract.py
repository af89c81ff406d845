"""bench.py contract, CPU side: the reference arm prints ONE JSON line with the keys the driver reads, on a bounded sample
(a tiny one here), and the workload builders produce the documented BASELINE configurations."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-n", "12", "--markers", "2000"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["dtype"] == "f64" and d["unit"] == "Mcell-updates/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    rs = d["cpu_baseline"]["reference_sources"]  # the reference's own sources timed beside the port, where oracle/_ref exists
    assert "error" not in rs, rs
    if os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libfluca_ref_ns.so")):
        assert rs["available"] and rs["value"] > 0 and rs["cores"] == 1 and "GMRES(30) + ILU(0)" in rs["sample"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]


def test_workload_builders_match_the_baseline_configs():
    from fluca_b200 import workloads as W

    c = W.sphere_bench_case(512, 512)  # config 4: [-4,12] x [-8,8]^2, h = 1/32, Re = 300, dt = 0.5 h
    f = c.faces()
    assert (f[0][0], f[0][-1], f[1][0], f[1][-1], f[2][0], f[2][-1]) == (-4.0, 12.0, -8.0, 8.0, -8.0, 8.0)
    assert np.isclose(f[0][1] - f[0][0], 1.0 / 32.0) and np.isclose(c.dt, 0.5 / 32.0) and np.isclose(c.mu, 1.0 / 300.0)
    assert [b["type"] for b in c.bcs] == [1, 2, 4, 4, 4, 4]  # inflow, pressure outlet, symmetry x 4
    c2 = W.sphere_bench_case(512, 1024)  # weak scaling: the box grows in z, the cell size does not
    assert np.isclose(c2.faces()[2][-1], 24.0)
    k = W.cavity_bench_case(256, 256)  # config 3
    assert np.isclose(k.mu, 1.0 / 400.0) and np.isclose(k.dt, 0.5 / 256.0)
    ch = W.channel_bench_case()  # config 5: 32 x 16 x 16 at h = 1/64, periodic x (and z), walls in y
    assert ch.n == (2048, 1024, 1024) and np.isclose(ch.hi[0], 32.0) and np.isclose(ch.hi[1], 16.0) and np.isclose(ch.dt, 0.5 / 64.0)
    assert [b["type"] for b in ch.bcs] == [3, 3, 1, 1, 3, 3]
    cen = np.array(W.channel_sphere_centres(ch.lo, ch.hi, 80))
    d = np.linalg.norm(cen[:, None] - cen[None], axis=2) + 99.0 * np.eye(80)
    assert cen.shape == (80, 3) and d.min() >= 1.5 and cen[:, 1].min() >= 1.0 and cen[:, 1].max() <= 15.0
    mk = W.multi_sphere_markers(cen[:3], 1.0, 12500, 1.0 / 64.0)
    assert mk["X"].shape == (3, 37500) and mk["dV"].shape == (37500,)


class _CpuCtx:
    """bench.Ctx without CUDA: one rank, wall-clock timer (the GPU arm's logic run on the host-emulation double)."""

    rank, world, local, dev = 0, 1, 0, None

    def barrier(self):
        pass

    def max_over_ranks(self, x):
        return x

    def comm(self):
        return None

    def stream_timer(self, solver):
        import time

        t = [0.0]

        def start():
            t[0] = time.perf_counter()

        return start, lambda: 1e3 * (time.perf_counter() - t[0])

    def pinned(self, shape):
        return np.empty(shape)

    def gather_z(self, arrs, axis):
        return arrs


def test_gpu_arm_logic_on_the_test_double():
    """The whole GPU arm of bench.py (parity self-check against the oracle, timed loop, rooflines per class, end-to-end loops)
    on the host-emulation library at a tiny size: the JSON line has every key of the contract and the parity case is green."""
    sys.path.insert(0, ROOT)
    import argparse

    import bench
    from tests import parity

    out = []
    bench.print_json = out.append
    args = argparse.Namespace(gpus=1, steps=2, warmup=1, impl="b200", workload="sphere", n=16, markers=300, scaling="both", strong=False, mode="coupled", restart=0, schur_ainv="ID", upper_ainv="ID", cpu_n=8, no_cpu_baseline=False, no_e2e=False, no_parity=False)
    bench.run_b200(args, ctx=_CpuCtx(), lib=parity.hostemu_library())
    assert len(out) == 1
    d = json.loads(out[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks", "parity", "class_rooflines", "step_roofline", "poisson_solve"):
        assert key in d, key
    assert d["parity"]["ok"], d["parity"]
    assert set(d["parity"]["runs"]) == {"ID/ID", "DIAG/ROWSUM"}
    rs = d["parity"]["reference_sources"]  # the same library against the reference's own compiled NS sources (where oracle/_ref exists)
    assert rs["available"] is False or (rs["ok"] and set(rs["runs"]) == {"fractional", "coupled"}), rs
    rs = d["cpu_baseline"]["reference_sources"]
    assert "error" not in rs and (rs["available"] is False or (rs["value"] > 0 and rs["cores"] == 1)), rs
    assert d["value"] > 0 and d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["value"] > 0 and d["gpu_launches"] > 0
    assert d["cpu_baseline"]["cores"] >= 1 and "absent" in d["cpu_baseline"]["reference_build"]
    assert d["config"]["workload"].startswith("BASELINE config 4")
    for args.workload, args.n in (("cavity", 16), ("channel", 16)):
        out.clear()
        args.no_parity = args.no_cpu_baseline = args.no_e2e = True
        bench.run_b200(args, ctx=_CpuCtx(), lib=parity.hostemu_library())
        d = json.loads(out[0])
        assert d["value"] > 0 and d["config"]["iterations_per_step"]["outer"]


def test_resident_e2e_loop_reads_every_step_result():
    """bench.py's e2e_resident loop (the glue's time loop with staged downloads), run here on the host-emulation double."""
    sys.path.insert(0, ROOT)
    import time

    import bench
    import fluca_b200 as fb
    from tests import cases, parity

    case = cases.cavity3d(n=(8, 8, 4))
    ns = parity.make_ns(case, parity.hostemu_library(), "fractional")
    parity.set_initial(ns, case.initial_state(seed=2))
    s = fb.NSB200GetSolver(ns)
    step0 = fb.NSGetTimeStep(ns)
    seconds, nbytes, acc = bench.e2e_resident_loop(ns, s, 3, lambda: None, time.perf_counter)
    assert fb.NSGetTimeStep(ns) == step0 + 3 and seconds > 0 and np.isfinite(acc)
    st = s.get_state()
    assert nbytes == st["v"].nbytes + st["p"].nbytes + st["phalf"].nbytes + sum(u.nbytes for u in st["U"])
    last = s.staged_state()  # the drained copy is the final state
    assert np.array_equal(last["v"], st["v"]) and np.array_equal(last["p"], st["p"])
    fb.NSDestroy(ns)


def _run_bench_world(world, tmp_path, workload="sphere", n=16):
    import socket

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / f"bench_{workload}_{world}.json")
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), OMP_NUM_THREADS="1", BENCH_WORKLOAD=workload, BENCH_N=str(n))
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "bench_worker.py"), out], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    logs = []
    for p in procs:
        try:
            o, _ = p.communicate(timeout=900)
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
        logs.append(o)
    assert all(p.returncode == 0 for p in procs), "\n".join(lg[-3000:] for lg in logs)
    return json.loads(open(out).read())


def test_gpu_arm_logic_multi_rank_gloo(tmp_path):
    """bench.py's whole N > 1 flow (N-rank parity case against the oracle, strong-scaling headline, weak run) with 2 and 4 ranks
    on the test double: every rank must issue the same collectives; the line carries parity, strong value and the weak run."""
    for world in (2, 4):
        d = _run_bench_world(world, tmp_path)
        assert d["n_gpus"] == world and d["scaling"] == "strong" and d["value"] > 0
        assert d["parity"]["ok"] and d["parity"]["ranks"] == world, d["parity"]
        assert d["weak"]["value"] > 0 and "WEAK" in d["weak"]["workload"] and "STRONG" in d["config"]["workload"]
