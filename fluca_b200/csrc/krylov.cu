// krylov.cu -- inner Krylov solvers of the ABF application.
//
//   momentum_solve : A v* = r_mom  (KSPSolve(abf->kspA), abfpc.c:72).  A is non-symmetric and
//                    strongly diagonally dominant (I + O(dt)): unpreconditioned BiCGStab with the
//                    dot products fused into the matrix-free A apply and the vector updates.
//   poisson_solve  : P p' = s      (KSPSolve(abf->kspS), abfpc.c:77).  P = vol * (-D Gst0) is
//                    symmetric positive semi-definite unless a pressure outlet exists: CG
//                    preconditioned by one geometric-multigrid V-cycle; with an outlet the one-sided
//                    wall derivative makes the boundary rows non-symmetric and the same V-cycle
//                    right-preconditions BiCGStab.
// Convergence is tested on the true residual 2-norm relative to the right-hand side (the reference's
// inner KSPs use PETSc's default preconditioned norm with ILU(0); inner histories cannot match
// pointwise across preconditioners, SURVEY.md section 7 "hard parts").
#include "solver.h"

namespace fluca {

// relative tolerance of an inner solve: the user's, relaxed by the outer solver as its residual drops (step.cu outer_gmres)
static double inner_rtol(const Solver &s, double rtol)
{
  const double r = rtol > s.tol_floor ? rtol : s.tol_floor;
  return r < 0.1 ? r : 0.1;
}

static Box cell_box(const Solver &s)
{
  Box b = {s.gh.g.nx, s.gh.g.ny, s.gh.g.nzl};
  return b;
}

// y = A x; out = {<a, y>, <y, y>, <x, y>, <a, x>}
static void a_apply_dots(Solver &s, const V3 &x, const V3 &y, const V3 &a, double out[4])
{
  halo_cells(s, x);
#ifndef FLUCA_HOSTEMU
  if (tma_usable(s)) {
    a_apply_dots_tma(s, x, y, a, true);
    reduce_finish(s, 4, out);
    return;
  }
#endif
  {
    KScope kt(s.ex, KT_MOMENTUM_APPLY);
    if (s.dim == 2) {
      AApplyDots<2> f;
      f.g = s.gh.g, f.sp = s.sp, f.bc = s.bc, f.x = CV3(x), f.v0 = CV3(s.v0), f.U0 = CV3(s.U0), f.a = CV3(a), f.y = y;
      for_box_reduce<4>(s.ex, cell_box(s), f);
    } else {
      AApplyDots<3> f;
      f.g = s.gh.g, f.sp = s.sp, f.bc = s.bc, f.x = CV3(x), f.v0 = CV3(s.v0), f.U0 = CV3(s.U0), f.a = CV3(a), f.y = y;
      for_box_reduce<4>(s.ex, cell_box(s), f);
    }
  }
  reduce_finish(s, 4, out);
}

struct P3 { // three component pointers offset to the interior planes
  double *c[3];
};
static P3 off3(const Solver &s, const V3 &v)
{
  P3 r;
  for (int q = 0; q < 3; ++q) r.c[q] = v.c[q] ? v.c[q] + interior_off(s) : nullptr;
  return r;
}

// the system solved is A x = bscale * b.  guess: x holds an initial guess on entry (the previous velocity, or the forced predictor of the IBM coupling); the
// tolerance stays relative to |b|, as PETSc's default convergence test does with a non-zero guess
int momentum_solve(Solver &s, const V3 &b, const V3 &x, bool guess, double bscale)
{
  KScope ks(s.ex, KT_MOMENTUM_VEC);
  const int    nc  = s.dim;
  const long   len = interior_len(s);
  const P3     B = off3(s, b), X = off3(s, x), R = off3(s, s.kr), RH = off3(s, s.krh), PV = off3(s, s.kp), VV = off3(s, s.kv), SV = off3(s, s.ks), TV = off3(s, s.kt);
  double       red[4];
  if (!guess) {
    // x = 0, r = rhat = p = b, rho = <b, b>
    for_range_reduce<2>(s.ex, len, FL_LAMBDA(long i, double acc[2]) {
      double t = 0.;
      _Pragma("unroll") for (int q = 0; q < 3; ++q) if (q < nc) {
        const double bv = bscale * B.c[q][i];
        X.c[q][i]  = 0.;
        R.c[q][i]  = bv;
        RH.c[q][i] = bv;
        PV.c[q][i] = bv;
        t += bv * bv;
      }
      acc[0] += t;
      acc[1] += t;
    });
  } else {
    // r = rhat = p = b - A x, rho = <r, r>
    a_apply(s, x, s.kv);
    for_range_reduce<2>(s.ex, len, FL_LAMBDA(long i, double acc[2]) {
      double t = 0., tb = 0.;
      _Pragma("unroll") for (int q = 0; q < 3; ++q) if (q < nc) {
        const double bv = bscale * B.c[q][i], rv = bv - VV.c[q][i];
        R.c[q][i]  = rv;
        RH.c[q][i] = rv;
        PV.c[q][i] = rv;
        t += rv * rv;
        tb += bv * bv;
      }
      acc[0] += t;
      acc[1] += tb;
    });
  }
  reduce_finish(s, 2, red);
  const double bnorm = std::sqrt(red[1]);
  s.stats.mom_last_rel = 0.;
  if (bnorm == 0.) {
    if (guess) for_range(s.ex, len, FL_LAMBDA(long i) { _Pragma("unroll") for (int q = 0; q < 3; ++q) if (q < nc) X.c[q][i] = 0.; });
    return 0;
  }
  if (!(bnorm == bnorm)) throw Error(FL_ERR_DIVERGED, "momentum right-hand side is NaN");
  const double tol = inner_rtol(s, s.opt.mom_rtol) * bnorm;
  double       rho = red[0], alpha = 1., omega = 1.;
  s.stats.mom_last_rel = std::sqrt(red[0]) / bnorm;
  s.monitor(0, 0, std::sqrt(red[0]));
  if (std::sqrt(red[0]) <= tol) return 0; // the guess already meets the tolerance
  int          it = 0;
  for (; it < s.opt.inner_maxit;) {
    // v = A p, <rhat, v>
    a_apply_dots(s, s.kp, s.kv, s.krh, red);
    const double rhv = red[0];
    if (rhv == 0. || !(rhv == rhv)) throw Error(FL_ERR_DIVERGED, "BiCGStab breakdown in the momentum solve (<rhat, A p> = 0)");
    alpha = rho / rhv;
    // s = r - alpha v, |s|^2
    for_range_reduce<1>(s.ex, len, FL_LAMBDA(long i, double acc[1]) {
      double t = 0.;
      _Pragma("unroll") for (int q = 0; q < 3; ++q) if (q < nc) {
        const double sv = R.c[q][i] - alpha * VV.c[q][i];
        SV.c[q][i]      = sv;
        t += sv * sv;
      }
      acc[0] += t;
    });
    reduce_finish(s, 1, red);
    ++it;
    if (std::sqrt(red[0]) <= tol) {
      for_range(s.ex, len, FL_LAMBDA(long i) {
        _Pragma("unroll") for (int q = 0; q < 3; ++q) if (q < nc) X.c[q][i] += alpha * PV.c[q][i];
      });
      s.stats.mom_last_rel = std::sqrt(red[0]) / bnorm;
      s.monitor(0, it, std::sqrt(red[0]));
      break;
    }
    // t = A s with four fused sums: <r^, t>, <t, t>, <s, t>, <r^, s>
    a_apply_dots(s, s.ks, s.kt, s.krh, red);
    if (red[1] == 0.) throw Error(FL_ERR_DIVERGED, "BiCGStab breakdown in the momentum solve (A s = 0)");
    omega = red[2] / red[1];
    // rho_new = <r^, r_new> = <r^, s> - omega <r^, t>: both sums come with the operator application (s and r^ are in registers
    // there), so the next rho and beta are known BEFORE r is formed, and the x / r update and the direction update are ONE pass
    // over the vectors (x, p, s, t, v -> x, r, p: 192 B per cell instead of 168 + 96) with one reduction instead of two.
    // (<r^, s> vanishes in exact arithmetic; dropping it stalls tight solves at 1e-8, so it is kept.)
    const double rho_new = red[3] - omega * red[0];
    if (rho_new == 0. || omega == 0. || !(rho_new == rho_new)) throw Error(FL_ERR_DIVERGED, "BiCGStab breakdown in the momentum solve (rho = 0)");
    const double beta = (rho_new / rho) * (alpha / omega);
    rho               = rho_new;
    // x += alpha p + omega s ; r = s - omega t ; p = r + beta (p - omega v) ; |r|^2
    for_range_reduce<1>(s.ex, len, FL_LAMBDA(long i, double acc[1]) {
      double t0 = 0.;
      _Pragma("unroll") for (int q = 0; q < 3; ++q) if (q < nc) {
        const double sv = SV.c[q][i], pv = PV.c[q][i];
        X.c[q][i] += alpha * pv + omega * sv;
        const double rv = sv - omega * TV.c[q][i];
        R.c[q][i]       = rv;
        PV.c[q][i]      = rv + beta * (pv - omega * VV.c[q][i]);
        t0 += rv * rv;
      }
      acc[0] += t0;
    });
    reduce_finish(s, 1, red);
    s.stats.mom_last_rel = std::sqrt(red[0]) / bnorm;
    s.monitor(0, it, std::sqrt(red[0]));
    if (std::sqrt(red[0]) <= tol) break;
    if (!(red[0] == red[0])) throw Error(FL_ERR_DIVERGED, "momentum residual is NaN");
  }
  s.stats.mom_its += it;
  return (it >= s.opt.inner_maxit) ? 1 : 0;
}

// ------------------------------------------------------------------ Poisson
template <int DIM>
struct PoissonApplyDot { // out = P p ; acc[0] += <a, out>
  Geom          g;
  const double *p, *a;
  double       *out;
  FL_HD void operator()(int i, int j, int kl, double acc[1]) const
  {
    Nbr<DIM> nb;
    nbr<DIM>(g, i, j, kl, nb);
    const double v = poisson_apply_cell<DIM>(g, p, nb);
    out[nb.c]      = v;
    acc[0] += a[nb.c] * v;
  }
};

// DIAG / ROWSUM Schur complement (abfpc.c:155-170): out = P p + vol D T ((1 - a1) .* G0 p), two passes with the
// cell field w = (1 - a1) .* G0 p in between (its halo plane is needed by T across a slab face); returns <a, out>
double schur_variant_apply_dot(Solver &s, double *pin, double *out, const double *a)
{
  if (!s.ainv_s.c[0] || !s.prepared) throw Error(FL_ERR_ARG, "the DIAG / ROWSUM Schur complement needs the momentum operator of a prepared step");
  halo_scalar(s, pin);
  {
    KScope kt(s.ex, KT_POISSON_APPLY);
    if (s.dim == 2) {
      GradScaleCells<2> gs;
      gs.g = s.gh.g, gs.scale = 1., gs.ainv = CV3(s.ainv_s), gs.p = pin, gs.vs = CV3(), gs.v = s.tw, gs.w = s.tw;
      for_box(s.ex, cell_box(s), gs);
    } else {
      GradScaleCells<3> gs;
      gs.g = s.gh.g, gs.scale = 1., gs.ainv = CV3(s.ainv_s), gs.p = pin, gs.vs = CV3(), gs.v = s.tw, gs.w = s.tw;
      for_box(s.ex, cell_box(s), gs);
    }
  }
  halo_cells(s, s.tw);
  {
    KScope kt(s.ex, KT_POISSON_APPLY);
    if (s.dim == 2) {
      SchurVariantApplyDot<2> f;
      f.g = s.gh.g, f.p = pin, f.a = a, f.w = CV3(s.tw), f.out = out;
      for_box_reduce<1>(s.ex, cell_box(s), f);
    } else {
      SchurVariantApplyDot<3> f;
      f.g = s.gh.g, f.p = pin, f.a = a, f.w = CV3(s.tw), f.out = out;
      for_box_reduce<1>(s.ex, cell_box(s), f);
    }
  }
  double r;
  reduce_finish(s, 1, &r);
  return r;
}

static double poisson_apply_dot(Solver &s, double *pin, double *out, const double *a)
{
  if (s.opt.schur_ainv != 0) return schur_variant_apply_dot(s, pin, out, a);
  halo_scalar(s, pin);
#ifndef FLUCA_HOSTEMU
  if (tma_usable(s)) {
    KScope kt(s.ex, KT_POISSON_APPLY);
    poisson_apply_dot_tma(s, pin, out, a);
  } else
#endif
  {
    KScope kt(s.ex, KT_POISSON_APPLY);
    if (s.dim == 2) {
      PoissonApplyDot<2> f;
      f.g = s.gh.g, f.p = pin, f.a = a, f.out = out;
      for_box_reduce<1>(s.ex, cell_box(s), f);
    } else {
      PoissonApplyDot<3> f;
      f.g = s.gh.g, f.p = pin, f.a = a, f.out = out;
      for_box_reduce<1>(s.ex, cell_box(s), f);
    }
  }
  double r;
  reduce_finish(s, 1, &r);
  return r;
}

void poisson_apply(Solver &s, double *pin, double *out)
{
  (void)poisson_apply_dot(s, pin, out, pin);
}

static int poisson_pcg(Solver &s, double *b, double *x)
{
  KScope ks(s.ex, KT_POISSON_VEC);
  const long off = interior_off(s), len = interior_len(s);
  double    *B = b + off, *X = x + off, *R = s.pr + off, *PP = s.pp + off, *Q = s.pq + off;
  double     red[2];
  for_range_reduce<1>(s.ex, len, FL_LAMBDA(long i, double acc[1]) {
    const double bv = B[i];
    X[i] = 0.;
    R[i] = bv;
    acc[0] += bv * bv;
  });
  reduce_finish(s, 1, red);
  const double bnorm = std::sqrt(red[0]);
  s.stats.schur_last_rel = 0.;
  if (bnorm == 0.) return 0;
  if (!(bnorm == bnorm)) throw Error(FL_ERR_DIVERGED, "Poisson right-hand side is NaN");
  const double tol = inner_rtol(s, s.opt.schur_rtol) * bnorm;
  s.monitor(1, 0, bnorm);
  // The V-cycle is not a symmetric operator (summed-residual restriction against trilinear prolongation; measured: 8 % asymmetry on a
  // 6^3 grid with a symmetry plane), so the Fletcher-Reeves beta of textbook PCG loses conjugacy in LONG solves: a rough right-hand
  // side then needs 185 iterations for 1e-5 and stalls at 1e-5 for good, where the flexible (Polak-Ribiere) beta
  //   beta = (<r_new, z_new> - <r_old, z_new>) / <r_old, z_old>,   <r_old, z_new> = <r_new, z_new> + alpha <q, z_new>
  // takes 10 and converges to round-off (found by the randomised comparison with the compiled reference).  It costs one
  // more reduction over (q, z) per iteration, so it is used from iteration `pr_from` of a solve on: the 4-6 iteration solves of the
  // cavity configurations never pay for it.  FLUCA_B200_PCG_PR_FROM=n moves the switch (0: always flexible, -1: textbook always).
  static const int pr_from = getenv("FLUCA_B200_PCG_PR_FROM") ? atoi(getenv("FLUCA_B200_PCG_PR_FROM")) : 4;
  double            rz = 0., alpha = 0.;
  int               it = 0;
  for (; it < s.opt.inner_maxit;) {
    // z = V-cycle(r); its last smoothing sweep also accumulates rz_new = <r, z>
    const double *Z = mg_vcycle(s, s.pr, s.opt.mg_nu2 > 0) + off;
    if (s.opt.mg_nu2 <= 0) for_range_reduce<1>(s.ex, len, FL_LAMBDA(long i, double acc[1]) { acc[0] += R[i] * Z[i]; });
    reduce_finish(s, 1, red);
    const double rz_new = red[0];
    if (it == 0) {
      for_range(s.ex, len, FL_LAMBDA(long i) { PP[i] = Z[i]; });
    } else {
      double qz = 0.;
      const bool flexible = pr_from >= 0 && it >= pr_from;
      if (flexible) { // q still holds P p of the previous iteration
        double r2[1];
        for_range_reduce<1>(s.ex, len, FL_LAMBDA(long i, double acc[1]) { acc[0] += Q[i] * Z[i]; });
        reduce_finish(s, 1, r2);
        qz = r2[0];
      }
      const double beta = flexible ? -alpha * qz / rz : rz_new / rz;
      for_range(s.ex, len, FL_LAMBDA(long i) { PP[i] = Z[i] + beta * PP[i]; });
    }
    rz = rz_new;
    const double pq = poisson_apply_dot(s, s.pp, s.pq, s.pp);
    if (!(pq > 0.)) {
      if (pq == 0.) break;
      throw Error(FL_ERR_DIVERGED, "CG breakdown in the pressure solve (<p, P p> <= 0 or NaN)");
    }
    alpha = rz / pq;
    for_range_reduce<1>(s.ex, len, FL_LAMBDA(long i, double acc[1]) {
      X[i] += alpha * PP[i];
      const double rv = R[i] - alpha * Q[i];
      R[i]            = rv;
      acc[0] += rv * rv;
    });
    reduce_finish(s, 1, red);
    ++it;
    s.stats.schur_last_rel = std::sqrt(red[0]) / bnorm;
    s.monitor(1, it, std::sqrt(red[0]));
    if (std::sqrt(red[0]) <= tol) break;
  }
  s.stats.schur_its += it;
  return (it >= s.opt.inner_maxit) ? 1 : 0;
}

// right-preconditioned BiCGStab (pressure outlet present: boundary rows of P are not symmetric)
static int poisson_bicgstab(Solver &s, double *b, double *x)
{
  KScope ks(s.ex, KT_POISSON_VEC);
  const long off = interior_off(s), len = interior_len(s);
  double    *B = b + off, *X = x + off, *R = s.pr + off, *RH = s.prh + off, *PP = s.pp + off, *V = s.pq + off, *S = s.ps + off, *T = s.pt + off;
  double     red[2];
  for_range_reduce<1>(s.ex, len, FL_LAMBDA(long i, double acc[1]) {
    const double bv = B[i];
    X[i]  = 0.;
    R[i]  = bv;
    RH[i] = bv;
    PP[i] = bv;
    acc[0] += bv * bv;
  });
  reduce_finish(s, 1, red);
  const double bnorm = std::sqrt(red[0]);
  s.stats.schur_last_rel = 0.;
  if (bnorm == 0.) return 0;
  if (!(bnorm == bnorm)) throw Error(FL_ERR_DIVERGED, "Poisson right-hand side is NaN");
  const double tol = inner_rtol(s, s.opt.schur_rtol) * bnorm;
  double       rho = red[0], alpha = 1., omega = 1.;
  int          it = 0;
  s.monitor(1, 0, bnorm);
  for (; it < s.opt.inner_maxit;) {
    double       *yf = mg_vcycle(s, s.pp); // y = M^-1 p
    const double *Y  = yf + off;
    const double  rhv = poisson_apply_dot(s, yf, s.pq, s.prh);
    if (rhv == 0. || !(rhv == rhv)) throw Error(FL_ERR_DIVERGED, "BiCGStab breakdown in the pressure solve");
    alpha = rho / rhv;
    for_range_reduce<1>(s.ex, len, FL_LAMBDA(long i, double acc[1]) {
      const double sv = R[i] - alpha * V[i];
      S[i]            = sv;
      X[i] += alpha * Y[i];
      acc[0] += sv * sv;
    });
    reduce_finish(s, 1, red);
    ++it;
    s.stats.schur_last_rel = std::sqrt(red[0]) / bnorm;
    if (std::sqrt(red[0]) <= tol) {
      s.monitor(1, it, std::sqrt(red[0]));
      break;
    }
    double       *zf = mg_vcycle(s, s.ps); // z = M^-1 s  (y is dead by now: x was updated above)
    const double *Z  = zf + off;
    const double  ts = poisson_apply_dot(s, zf, s.pt, s.ps);
    for_range_reduce<1>(s.ex, len, FL_LAMBDA(long i, double acc[1]) { acc[0] += T[i] * T[i]; });
    reduce_finish(s, 1, red);
    if (red[0] == 0.) throw Error(FL_ERR_DIVERGED, "BiCGStab breakdown in the pressure solve (P z = 0)");
    omega = ts / red[0];
    for_range_reduce<2>(s.ex, len, FL_LAMBDA(long i, double acc[2]) {
      X[i] += omega * Z[i];
      const double rv = S[i] - omega * T[i];
      R[i]            = rv;
      acc[0] += rv * rv;
      acc[1] += RH[i] * rv;
    });
    reduce_finish(s, 2, red);
    s.stats.schur_last_rel = std::sqrt(red[0]) / bnorm;
    s.monitor(1, it, std::sqrt(red[0]));
    if (std::sqrt(red[0]) <= tol) break;
    if (!(red[0] == red[0])) throw Error(FL_ERR_DIVERGED, "pressure residual is NaN");
    if (red[1] == 0. || omega == 0.) throw Error(FL_ERR_DIVERGED, "BiCGStab breakdown in the pressure solve (rho = 0)");
    const double beta = (red[1] / rho) * (alpha / omega);
    rho               = red[1];
    for_range(s.ex, len, FL_LAMBDA(long i) { PP[i] = R[i] + beta * (PP[i] - omega * V[i]); });
  }
  s.stats.schur_its += it;
  return (it >= s.opt.inner_maxit) ? 1 : 0;
}

int poisson_solve(Solver &s, double *b, double *x)
{
  // the DIAG / ROWSUM Schur complements are not symmetric either; the V-cycle of P stays their preconditioner
  // (S' - S = D T (I - a1) G~ is O(dt) relative to S)
  return (s.has_outlet || s.opt.schur_ainv != 0) ? poisson_bicgstab(s, b, x) : poisson_pcg(s, b, x);
}

} // namespace fluca
