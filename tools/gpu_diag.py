"""GPU diagnostic: inner iteration counts vs grid size (fractional step, 1e-10)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fluca_b200 as fb
from tests import cases, parity

lib = fb._lib.load()
sizes = [int(a) for a in sys.argv[1:]] or [32, 64, 128, 256]
for n in sizes:
    case = cases.cavity3d_full(n=(n, n, n), Re=400.0)
    ns = parity.make_ns(case, lib, "fractional", ns_abf_momentum_ksp_rtol=1e-10, ns_abf_schur_ksp_rtol=1e-10)
    for k in range(2):
        t = time.time()
        fb.NSStep(ns)
        st = fb.NSB200GetStats(ns)
        print(n, "step", k, "mom", st.mom_its, "schur", st.schur_its, st.schur_last_rel, f"{time.time()-t:.2f}s", flush=True)
    fb.NSDestroy(ns)
