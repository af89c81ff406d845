"""Run the CUDA library and the host-emulation double side by side on the GPU box (debug aid)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import fluca_b200 as fb
from tests import cases, parity

n = int(sys.argv[1]) if len(sys.argv) > 1 else 160
gl, el = fb._lib.load(), parity.hostemu_library()
case = cases.cavity3d_full(n=(n, n, n), Re=400.0)
rng = np.random.default_rng(0)
r = rng.standard_normal((n, n, n)); r -= r.mean()
out = {}
for name, L in (("gpu", gl), ("emu", el)):
    ns = parity.make_ns(case, L, "fractional")
    s = fb.NSB200GetSolver(ns)
    t = time.time()
    out[name] = dict(z=s.apply_vcycle(r), sp=s.apply_schur(r))
    print(name, f"{time.time()-t:.1f}s", flush=True)
    fb.NSDestroy(ns)
for k in ("sp", "z"):
    a, b = out["gpu"][k], out["emu"][k]
    d = np.abs(a - b)
    print(k, "rel diff", np.linalg.norm(a - b) / np.linalg.norm(b), "max at", np.unravel_index(d.argmax(), d.shape), d.max())
    if d.max() > 1e-9 * np.abs(b).max():
        bad = np.argwhere(d > 1e-9 * np.abs(b).max())
        print("  bad count", len(bad), "k range", bad[:, 0].min(), bad[:, 0].max(), "j range", bad[:, 1].min(), bad[:, 1].max(), "i range", bad[:, 2].min(), bad[:, 2].max())
        ks, cnt = np.unique(bad[:, 0], return_counts=True)
        print("  bad per plane (first 20):", list(zip(ks[:20], cnt[:20])))
