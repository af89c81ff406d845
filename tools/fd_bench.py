"""Times the generated FlucaFDApply kernel (csrc/fd.cu, v1) on device-resident fields: a 7-point Laplacian (sum of three
second derivatives, Dirichlet / Neumann / periodic mix) on n^3 elements.  Algorithmic traffic: 8 B read + 8 B written per
output point.  Not part of bench.py's contract; prints one JSON line.  --assembled times the v2 ELL path.

    python tools/fd_bench.py --n 512 --reps 20
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=512)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--assembled", action="store_true", help="time the v2 assembled (ELL) apply of the same operator instead of the class-table kernel")
    args = ap.parse_args()
    if args.assembled:
        os.environ["FLUCA_B200_FD_ASSEMBLED"] = "1"
    import torch

    import fluca_b200 as fb
    from fluca_b200 import fd as FD

    lib = fb._lib.load()
    n = args.n
    g = FD.FDGrid.uniform([n, n, n], [0.0] * 3, [1.0] * 3, periodic=[False, False, True], library=lib)
    ops = [FD.FlucaFDDerivativeCreate(g, d, 2, 2, FD.DMSTAG_ELEMENT, 0, FD.DMSTAG_ELEMENT, 0).SetUp() for d in range(3)]
    lap = FD.FlucaFDSumCreate(ops)
    lap.SetBoundaryCondition(0, FD.FLUCAFD_BC_DIRICHLET, 1.0)
    lap.SetBoundaryCondition(2, FD.FLUCAFD_BC_NEUMANN, 0.0)
    lap.SetUp()
    assert lap.ApplyInputs() == [(FD.DMSTAG_ELEMENT, 0)]
    x = torch.randn(n, n, n, dtype=torch.float64, device="cuda")
    y = torch.empty_like(x)
    torch.cuda.synchronize()
    stream = torch.cuda.ExternalStream(lap.Stream())
    for _ in range(3):
        lap.ApplyDevice([x.data_ptr()], y.data_ptr())
    lap.Sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record(stream)
        for _ in range(args.reps):
            lap.ApplyDevice([x.data_ptr()], y.data_ptr())
        e1.record(stream)
    e1.synchronize()
    ms = e0.elapsed_time(e1) / args.reps
    # interior check against torch on the same data: (x[i-1] - 2 x[i] + x[i+1]) / h^2 summed over the directions
    h2 = (1.0 / n) ** 2
    ref = (x[2:, 1:-1, 1:-1] + x[:-2, 1:-1, 1:-1] + x[1:-1, 2:, 1:-1] + x[1:-1, :-2, 1:-1] + x[1:-1, 1:-1, 2:] + x[1:-1, 1:-1, :-2] - 6.0 * x[1:-1, 1:-1, 1:-1]) / h2
    err = float((y[1:-1, 1:-1, 1:-1] - ref).abs().max() / ref.abs().max())
    peak = 6650.0
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", peak))
    except Exception:
        pass
    ach = 16.0 * n**3 / (ms * 1e-3) / 1e9
    if args.assembled:  # the table is the traffic: 7 entries x (8 B weight + 4 B index) + 8 B constant + 8 B result, + the gathered x
        ach = (7 * 12.0 + 8.0 + 8.0 + 8.0) * n**3 / (ms * 1e-3) / 1e9
    print(json.dumps({"kernel": "fd_apply_laplacian_assembled_ell" if args.assembled else "fd_apply_laplacian", "n": n, "avg_ms": ms, "reps": args.reps, "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak}, "interior_rel_err_vs_torch": err}))


if __name__ == "__main__":
    main()
