import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import fluca_b200 as fb
from tests import cases, parity
n = int(sys.argv[1]) if len(sys.argv) > 1 else 192
os.environ["FLUCA_B200_DEBUG"] = "1"
gl, el = fb._lib.load(), parity.hostemu_library()
case = cases.cavity3d_full(n=(n, n, n), Re=400.0)
for name, L in (("gpu", gl), ("emu", el)):
    print("=====", name, flush=True); sys.stderr.flush()
    ns = parity.make_ns(case, L, "fractional", ns_abf_momentum_ksp_rtol=1e-10, ns_abf_schur_ksp_rtol=1e-10)
    fb.NSStep(ns)
    st = fb.NSB200GetStats(ns)
    print(name, "mom", st.mom_its, "schur", st.schur_its, flush=True)
    fb.NSDestroy(ns)
