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
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]


def test_workload_builders_match_the_baseline_configs():
    sys.path.insert(0, ROOT)
    import bench

    c = bench.sphere_case(512, 512)  # config 4: [-4,12] x [-8,8]^2, h = 1/32, Re = 300, dt = 0.5 h
    f = c.faces()
    assert (f[0][0], f[0][-1], f[1][0], f[1][-1], f[2][0], f[2][-1]) == (-4.0, 12.0, -8.0, 8.0, -8.0, 8.0)
    assert np.isclose(f[0][1] - f[0][0], 1.0 / 32.0) and np.isclose(c.dt, 0.5 / 32.0) and np.isclose(c.mu, 1.0 / 300.0)
    assert [b["type"] for b in c.bcs] == [1, 2, 4, 4, 4, 4]  # inflow, pressure outlet, symmetry x 4
    c2 = bench.sphere_case(512, 1024)  # weak scaling: the box grows in z, the cell size does not
    assert np.isclose(c2.faces()[2][-1], 24.0)
    k = bench.cavity_case(256, 256)  # config 3
    assert np.isclose(k.mu, 1.0 / 400.0) and np.isclose(k.dt, 0.5 / 256.0)


def test_poisson_solve_rate_is_the_second_half_of_the_baseline_metric():
    """'Poisson solve HBM GB/s vs peak' = iterations x bytes per iteration and cell x cells / event-timed solve time."""
    sys.path.insert(0, ROOT)
    import bench

    kt = {"poisson_apply": (2.0, 10), "poisson_vec": (3.0, 30), "mg_smooth": (4.0, 100), "mg_transfer": (1.0, 50), "momentum_apply": (99.0, 5)}
    r = bench.poisson_solve_rate(kt, [5, 5], 1.0e6, outlet=False, variant=False, peak=6530.3)
    assert np.isclose(r["achieved"], 10 * 227.0 * 1.0e6 / 10.0e-3 / 1e9) and np.isclose(r["frac"], r["achieved"] / 6530.3) and r["krylov"] == "pcg+mg"
    assert bench.poisson_solve_rate(kt, [5, 5], 1.0e6, outlet=True, variant=False, peak=1.0)["bytes_per_iteration_per_cell"] == 454.0
    assert bench.poisson_solve_rate(kt, [5, 5], 1.0e6, outlet=False, variant=True, peak=1.0)["bytes_per_iteration_per_cell"] == 630.0
    # never raises, never divides by zero: a reporting extra must not cost the bench line
    assert bench.poisson_solve_rate({}, [5], 1.0e6, False, False, 1.0) is None
    assert bench.poisson_solve_rate(kt, [], 1.0e6, False, False, 1.0) is None
    assert bench.poisson_solve_rate(None, None, None, False, False, 0.0) is None


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
