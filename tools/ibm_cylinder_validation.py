"""Physical check of the immersed-boundary coupling (SURVEY.md 8c "IBM": no reference code exists, so the method is validated
against the literature): 2-D flow past a circular cylinder at Re = 100 on BASELINE config 2's domain [-8,24] x [-8,8] with its
boundary set (inflow, pressure outlet, symmetry), D = 1 at (0, 0.0137) -- off the grid's symmetry line so that shedding starts
by itself --, marker spacing ~ h, two direct-forcing passes, fractional mode at the default tolerances.  Runs the product's own
algorithm in the host-emulation build (CPU; the CUDA kernels are checked against the same definition by tests/test_ibm.py).
Literature: C_D ~ 1.33-1.38, St ~ 0.164-0.166 (not from the reference).

    python tools/ibm_cylinder_validation.py <cells per unit length> <steps> <history.json> [marker retraction / h] [half height] [marker count] [mode] [upstream length]

The two optional arguments probe the two effects that raise the drag above the unbounded-domain literature value: markers
placed on a circle of radius D/2 - retraction * h (the regularised delta makes the body act larger), and symmetry planes at
+- half height instead of BASELINE config 2's +- 8 (blockage).
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests import cases, parity
import fluca_b200 as fb
lib = parity.hostemu_library()
hinv = int(sys.argv[1]); nsteps = int(sys.argv[2]); out = sys.argv[3]
retract = float(sys.argv[4]) if len(sys.argv) > 4 else 0.0
half = float(sys.argv[5]) if len(sys.argv) > 5 else 8.0
mode = sys.argv[7] if len(sys.argv) > 7 else "fractional"
upstream = float(sys.argv[8]) if len(sys.argv) > 8 else 8.0
n = (int((upstream + 24.0)*hinv), int(2*half*hinv))
h = 1.0/hinv
def inflow(dim, t, x):
    shape = np.shape(x[0])
    return [np.full(shape, 1.0), np.zeros(shape)] if shape else (1.0, 0.0)
inflow.vectorized = True; inflow.time_independent = True
def pout(dim, t, x):
    shape = np.shape(x[0]); return np.zeros(shape) if shape else 0.0
pout.vectorized = True; pout.time_independent = True
inl = dict(type=cases.BC_VELOCITY, velocity=inflow, pressure=None)
outl = dict(type=cases.BC_PRESSURE_OUTLET, velocity=None, pressure=pout)
sym = dict(type=cases.BC_SYMMETRY, velocity=None, pressure=None)
case = cases.Case("cylinder2d", n, (-upstream, -half), (24.0, half), 1.0, 1.0/100.0, 0.5*h, [inl, outl, sym, sym])
ns = parity.make_ns(case, lib, mode)
v, U, p = case.initial_state()
v[0] = 1.0; U[0][...] = 1.0
parity.set_initial(ns, (v, U, p))
nm = int(sys.argv[6]) if len(sys.argv) > 6 and int(sys.argv[6]) > 0 else int(np.ceil(np.pi*1.0/h))  # default: marker spacing ~ h; BASELINE config 2 fixes 1024
mk = cases.cylinder_markers((0.0, 0.0137), 1.0 - 2.0*retract*h, nm, h)
fb.NSB200SetMarkers(ns, mk["X"], mk["Ud"], mk["dV"], 4, 2)
hist = []
t0 = time.time()
for k in range(nsteps):
    fb.NSStep(ns)
    F, Um = fb.NSB200GetMarkerForces(ns)
    cd = -2.0*F[0].sum(); cl = -2.0*F[1].sum()   # force on the body = -force on the fluid; 1/2 rho U^2 D = 1/2
    hist.append((float((k+1)*case.dt), float(cd), float(cl), float(np.abs(Um).max())))
    if k % 50 == 0 or k == nsteps-1:
        print(k, "t=%.2f" % hist[-1][0], "CD=%.4f CL=%.4f slip=%.3e" % hist[-1][1:], "elapsed %.0fs" % (time.time()-t0), flush=True)
        json.dump(hist, open(out, "w"))
