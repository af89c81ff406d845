// mg.cu -- matrix-free geometric multigrid V-cycle, the preconditioner of the pressure solve.
//
// Replaces the ILU(0) that PETSc applies to the explicitly formed Schur complement S
// (abfpc.c:151-180; S = -(dt/rho) D Gst is a 5/7-point Laplacian with Neumann rows at walls).
// Cell-centred, finite-volume (flux) form on the rectilinear product grid:
//   (P x)_c = sum_d area_d [ k_d(lo) (x_c - x_lo) + k_d(hi) (x_c - x_hi) ],   k = 1 / centre distance
// Components (chosen with tools/mg_prototype.py: 5 PCG iterations for 1e-5, 8 for 1e-8, mesh
// independent): damped-Jacobi smoothing (order independent => identical on any slab partition),
// residual fused with summation restriction (flux-form residuals are extensive), trilinear
// prolongation fused with the correction, rediscretised coarse operators, coarsening by 2 per
// direction while the (local) size stays even.  At pressure outlets the hierarchy uses a plain
// first-order Dirichlet row; the Krylov method outside applies the exact one-sided operator.
#include "solver.h"
#ifndef FLUCA_HOSTEMU
#include "tma.h"
#endif

namespace fluca {

namespace {

// row of the level operator from the centre value and the two neighbours per direction (out-of-range
// neighbours are passed as 0: they meet k = 0 at Neumann walls and are the Dirichlet ghost at outlets)
template <int DIM>
FL_HD void mg_row_core(const MGLevel &L, int i, int j, int kl, double xc, const double xm[3], const double xp[3], double &Ax, double &diag)
{
  const double hx = L.h[0][i], hy = L.h[1][j];
  const double hz = (DIM == 3) ? L.h[2][L.k0 + kl] : 1.;
  const double ax = hy * hz, ay = hx * hz, az = hx * hy;
  double       s = 0., dg = 0.;
  {
    const double kl_ = L.kf[0][i], ku = L.kf[0][i + 1];
    s += ax * (kl_ * (xc - xm[0]) + ku * (xc - xp[0]));
    dg += ax * (kl_ + ku);
  }
  {
    const double kl_ = L.kf[1][j], ku = L.kf[1][j + 1];
    s += ay * (kl_ * (xc - xm[1]) + ku * (xc - xp[1]));
    dg += ay * (kl_ + ku);
  }
  if (DIM == 3) {
    const int    kg = L.k0 + kl;
    const double kl_ = L.kf[2][kg], ku = L.kf[2][kg + 1];
    s += az * (kl_ * (xc - xm[2]) + ku * (xc - xp[2]));
    dg += az * (kl_ + ku);
  }
  Ax   = s;
  diag = dg;
}

template <int DIM>
FL_HD void mg_row(const MGLevel &L, const double *__restrict__ x, int i, int j, int kl, double &Ax, double &diag)
{
  const long c = L.idx(i, j, kl);
  double     xm[3], xp[3];
  {
    const int n = L.n[0];
    xm[0] = i > 0 ? x[c - 1] : (L.per[0] ? x[c + (n - 1)] : 0.);
    xp[0] = i < n - 1 ? x[c + 1] : (L.per[0] ? x[c - (n - 1)] : 0.);
  }
  {
    const int n = L.n[1];
    xm[1] = j > 0 ? x[c - L.px] : (L.per[1] ? x[c + (long)(n - 1) * L.px] : 0.);
    xp[1] = j < n - 1 ? x[c + L.px] : (L.per[1] ? x[c - (long)(n - 1) * L.px] : 0.);
  }
  // ghost planes: neighbour rank / periodic wrap, or zeros at a wall
  xm[2] = (DIM == 3) ? x[c - L.plane] : 0.;
  xp[2] = (DIM == 3) ? x[c + L.plane] : 0.;
  mg_row_core<DIM>(L, i, j, kl, x[c], xm, xp, Ax, diag);
}

template <int DIM>
struct MGSmooth {
  MGLevel       L;
  double        omega;
  int           zero_guess;
  const double *xin, *b;
  double       *xout;
  FL_HD double sweep(int i, int j, int kl) const
  {
    const long c = L.idx(i, j, kl);
    double     Ax, dg;
    if (zero_guess) {
      // diag only
      const double hx = L.h[0][i], hy = L.h[1][j], hz = (DIM == 3) ? L.h[2][L.k0 + kl] : 1.;
      dg = hy * hz * (L.kf[0][i] + L.kf[0][i + 1]) + hx * hz * (L.kf[1][j] + L.kf[1][j + 1]);
      if (DIM == 3) dg += hx * hy * (L.kf[2][L.k0 + kl] + L.kf[2][L.k0 + kl + 1]);
      const double v = dg > 0. ? omega * b[c] / dg : 0.;
      fl_store(xout + c, v);
      return b[c] * v;
    }
    if (DIM == 3 && kl + FL_PF < L.nzl) fl_prefetch(xin + c + FL_PF * L.plane), fl_prefetch(b + c + FL_PF * L.plane);
    mg_row<DIM>(L, xin, i, j, kl, Ax, dg);
    const double bc = b[c], v = dg > 0. ? xin[c] + omega * (bc - Ax) / dg : xin[c];
    fl_store(xout + c, v);
    return bc * v;
  }
  FL_HD void operator()(int i, int j, int kl) const { (void)sweep(i, j, kl); }
  // the last sweep of the cycle also accumulates <b, x> = <r, z> for the CG that called it
  FL_HD void operator()(int i, int j, int kl, double acc[1]) const { acc[0] += sweep(i, j, kl); }
};

#ifndef FLUCA_HOSTEMU
// the same sweep from TMA-staged tiles (3-D levels of at least one tile per plane, see tma.h)
template <int NRED>
struct MGSmoothTile : TileOpDefaults {
  static const int NIN = 1, NR = NRED, MINB = 4, STAGES = 8, PLANES = FL_TILE_PLANES;
  MGLevel          L;
  double           omega;
  const double    *b;
  double          *xout;
  struct Regs {
    double b;
  };
  __device__ int flags(int i, int j) const { return (L.uni && (L.per[0] || (i > 0 && i < L.n[0] - 1)) && j > 0 && j < L.n[1] - 1) ? 1 : 0; }
  __device__ void prefetch(Regs &rg, int off, int kl) const { rg.b = b[off]; }
  __device__ void cell(const TileView &tv, const Regs &rg, int fl, int i, int j, int kl, int off, double *acc) const
  {
    const int    lc = tv.lc;
    const double xc = tv.p0[lc];
    const double xm[3] = {tv.p0[lc - 1], tv.p0[lc - TLX], tv.pm[lc]}, xp[3] = {tv.p0[lc + 1], tv.p0[lc + TLX], tv.pp[lc]};
    const int    kg = L.k0 + kl;
    const bool   inter = fl && (L.per[2] || (kg > 0 && kg < L.n[2] - 1));
    double       v;
    if (__all_sync(__activemask(), inter)) {
      // uniform level, no wall in reach: constant row, no table load, no division
      const double Ax = L.cd[0] * (2. * xc - xm[0] - xp[0]) + L.cd[1] * (2. * xc - xm[1] - xp[1]) + L.cd[2] * (2. * xc - xm[2] - xp[2]);
      v               = xc + (omega * L.idg) * (rg.b - Ax);
    } else {
      double Ax, dg;
      mg_row_core<3>(L, i, j, kl, xc, xm, xp, Ax, dg);
      v = dg > 0. ? xc + omega * (rg.b - Ax) / dg : xc;
    }
    xout[off] = v;
    if (NRED > 0) acc[0] += rg.b * v;
  }
};

// diagonal of the level operator at cell (i, j, global plane kg)
__device__ __forceinline__ double mg_diag3(const MGLevel &L, int i, int j, int kg)
{
  const double hx = L.h[0][i], hy = L.h[1][j], hz = L.h[2][kg];
  return hy * hz * (L.kf[0][i] + L.kf[0][i + 1]) + hx * hz * (L.kf[1][j] + L.kf[1][j + 1]) + hx * hy * (L.kf[2][kg] + L.kf[2][kg + 1]);
}

// The first two pre-smoothing sweeps of a V-cycle in one pass: with a zero guess the first sweep is x1 = omega b / diag,
// so the second one, x2 = x1 + omega (b - P x1) / diag, needs only b at the cell and its six neighbours.  The tile holds b
// (ghost planes exchanged by the caller); x1 is never stored: 16 B/cell instead of 16 + 24.
struct MGFirstTwoTile : TileOpDefaults {
  static const int NIN = 1, NR = 0, MINB = 4, STAGES = 8, PLANES = FL_TILE_PLANES;
  MGLevel          L;
  double           omega, omega2; // weights of the first and the second sweep
  double          *xout;
  struct Regs {
  };
  __device__ int  flags(int i, int j) const { return (L.uni && (L.per[0] || (i >= 2 && i <= L.n[0] - 3)) && j >= 2 && j <= L.n[1] - 3) ? 1 : 0; }
  __device__ void prefetch(Regs &, int, int) const { }
  __device__ double first(double bv, int i, int j, int kg) const
  {
    const double dg = mg_diag3(L, i, j, kg);
    return dg > 0. ? omega * bv / dg : 0.;
  }
  __device__ void cell(const TileView &tv, const Regs &, int fl, int i, int j, int kl, int off, double *) const
  {
    const int    lc = tv.lc, kg = L.k0 + kl, n2 = L.n[2];
    const double bc = tv.p0[lc];
    const double bm[3] = {tv.p0[lc - 1], tv.p0[lc - TLX], tv.pm[lc]}, bp[3] = {tv.p0[lc + 1], tv.p0[lc + TLX], tv.pp[lc]};
    const bool   inter = fl && (L.per[2] || (kg >= 2 && kg <= n2 - 3));
    double       v;
    if (__all_sync(__activemask(), inter)) {
      // uniform level, the cell and its neighbours away from every wall: x1 = w b with one constant w
      const double w1 = omega * L.idg, w2 = omega2 * L.idg;
      const double Ab = L.cd[0] * (2. * bc - bm[0] - bp[0]) + L.cd[1] * (2. * bc - bm[1] - bp[1]) + L.cd[2] * (2. * bc - bm[2] - bp[2]);
      v               = (w1 + w2) * bc - w1 * w2 * Ab;
    } else {
      // neighbours outside the domain count as zero (Neumann: zero conductance; outlet: Dirichlet ghost), as in mg_row
      const double x1c = first(bc, i, j, kg);
      double       xm[3], xp[3];
      xm[0] = i > 0 ? first(bm[0], i - 1, j, kg) : (L.per[0] ? first(bm[0], L.n[0] - 1, j, kg) : 0.);
      xp[0] = i < L.n[0] - 1 ? first(bp[0], i + 1, j, kg) : (L.per[0] ? first(bp[0], 0, j, kg) : 0.);
      xm[1] = j > 0 ? first(bm[1], i, j - 1, kg) : 0.;
      xp[1] = j < L.n[1] - 1 ? first(bp[1], i, j + 1, kg) : 0.;
      xm[2] = kg > 0 ? first(bm[2], i, j, kg - 1) : (L.per[2] ? first(bm[2], i, j, n2 - 1) : 0.);
      xp[2] = kg < n2 - 1 ? first(bp[2], i, j, kg + 1) : (L.per[2] ? first(bp[2], i, j, 0) : 0.);
      double Ax, dg;
      mg_row_core<3>(L, i, j, kl, x1c, xm, xp, Ax, dg);
      v = dg > 0. ? x1c + omega2 * (bc - Ax) / dg : x1c;
    }
    xout[off] = v;
  }
};

// coarse b = sum over the 2 x 2 x 2 children of (b - P x), from TMA-staged tiles of x.  A thread forms the residual of its
// fine cell; x-pairs are summed with a warp shuffle, y-pairs through CTA scratch after the plane barrier, z-pairs in a
// register across two consecutive planes (chunks start at even planes: ZALIGN = 2).
struct MGResidTile : TileOpDefaults {
  // scratch: one half-row buffer per plane of two consecutive trips (a trip holds up to two planes; the next trip must not
  // overwrite what a slower thread still reads in post())
  static const int  NIN = 1, NR = 0, MINB = 4, STAGES = 8, ZALIGN = 2, PLANES = FL_TILE_PLANES, SCRATCH = 4 * TMY * (TMX / 2) * (int)sizeof(double);
  static const bool POST = true;
  MGLevel           F, C;
  struct Regs {
    double b;
  };
  __device__ int  flags(int i, int j) const { return (F.uni && (F.per[0] || (i > 0 && i < F.n[0] - 1)) && j > 0 && j < F.n[1] - 1) ? 1 : 0; }
  __device__ void prefetch(Regs &rg, int off, int) const { rg.b = F.b[off]; }
  __device__ void cell(const TileView &tv, const Regs &rg, int fl, int i, int j, int kl, int, double *) const
  {
    const int    lc = tv.lc, kg = F.k0 + kl, tx = threadIdx.x & (TMX - 1), ty = threadIdx.x / TMX;
    const double xc = tv.p0[lc];
    const double xm[3] = {tv.p0[lc - 1], tv.p0[lc - TLX], tv.pm[lc]}, xp[3] = {tv.p0[lc + 1], tv.p0[lc + TLX], tv.pp[lc]};
    const bool   inter = fl && (F.per[2] || (kg > 0 && kg < F.n[2] - 1));
    double       Ax, dg;
    if (__all_sync(__activemask(), inter)) Ax = F.cd[0] * (2. * xc - xm[0] - xp[0]) + F.cd[1] * (2. * xc - xm[1] - xp[1]) + F.cd[2] * (2. * xc - xm[2] - xp[2]);
    else mg_row_core<3>(F, i, j, kl, xc, xm, xp, Ax, dg);
    double r = rg.b - Ax;
    r += __shfl_xor_sync(__activemask(), r, 1); // cells (2m, 2m+1) of a row are owned together (tile origins are even)
    if (!(tx & 1)) tv.scratch[((kl & 3) * TMY + ty) * (TMX / 2) + (tx >> 1)] = r;
  }
  __device__ void post(const TileView &tv, int, bool owned, int i, int j, int kl, double &zsum) const
  {
    const int tx = threadIdx.x & (TMX - 1), ty = threadIdx.x / TMX;
    if (!owned || (tx & 1) || (ty & 1)) return;
    const double *row = tv.scratch + ((kl & 3) * TMY + ty) * (TMX / 2) + (tx >> 1);
    const double  s   = row[0] + row[TMX / 2];
    if (!(kl & 1)) zsum = s;
    else C.b[C.idx(i >> 1, j >> 1, kl >> 1)] = zsum + s;
  }
};
#endif

// coarse b = sum over children of (b - P x)
template <int DIM>
struct MGResidRestrict {
  MGLevel F, C;
  FL_HD void operator()(int I, int J, int K) const
  {
    double    s   = 0.;
    const int cfz = (DIM == 3) ? F.cf[2] : 1;
    for (int dk = 0; dk < cfz; ++dk)
      for (int dj = 0; dj < F.cf[1]; ++dj)
        for (int di = 0; di < F.cf[0]; ++di) {
          const int i = F.cf[0] * I + di, j = F.cf[1] * J + dj, kl = cfz * K + dk;
          if (DIM == 3 && kl + 2 * FL_PF < F.nzl) fl_prefetch(F.x + F.idx(i, j, kl + 2 * FL_PF)), fl_prefetch(F.b + F.idx(i, j, kl + 2 * FL_PF));
          double    Ax, dg;
          mg_row<DIM>(F, F.x, i, j, kl, Ax, dg);
          s += F.b[F.idx(i, j, kl)] - Ax;
        }
    fl_store(C.b + C.idx(I, J, K), s);
  }
};

// 1-D cell-centred linear interpolation: fine index i -> coarse (I0, w0), (I1, w1)
FL_HD void interp1(int i, int cf, int nc, int per, int &I0, int &I1, double &w1)
{
  if (cf == 1) {
    I0 = I1 = i;
    w1 = 0.;
    return;
  }
  I0 = i >> 1;
  I1 = (i & 1) ? I0 + 1 : I0 - 1;
  w1 = 0.25;
  if (I1 < 0) {
    if (per) I1 = nc - 1;
    else I1 = I0, w1 = 0.;
  } else if (I1 >= nc) {
    if (per) I1 = 0;
    else I1 = I0, w1 = 0.;
  }
}

template <int DIM>
struct MGProlong {
  MGLevel F, C;
  FL_HD void operator()(int i, int j, int kl) const
  {
    int    I0, I1, J0, J1;
    double wi, wj;
    interp1(i, F.cf[0], C.n[0], C.per[0], I0, I1, wi);
    interp1(j, F.cf[1], C.n[1], C.per[1], J0, J1, wj);
    const double *e = C.x;
    double        v;
    if (DIM == 3 && kl + FL_PF < F.nzl) fl_prefetch(F.x + F.idx(i, j, kl + FL_PF));
    if (DIM == 3) {
      // z: local coarse index with ghost planes; clamp only at a physical wall
      int    K0, K1;
      double wk;
      if (F.cf[2] == 1) K0 = K1 = kl, wk = 0.;
      else {
        K0 = kl >> 1;
        K1 = (kl & 1) ? K0 + 1 : K0 - 1;
        wk = 0.25;
        if (K1 < 0 && C.wall_lo_z) K1 = K0, wk = 0.;
        if (K1 >= C.nzl && C.wall_hi_z) K1 = K0, wk = 0.;
      }
      double a0 = (1. - wi) * ((1. - wj) * e[C.idx(I0, J0, K0)] + wj * e[C.idx(I0, J1, K0)]) + wi * ((1. - wj) * e[C.idx(I1, J0, K0)] + wj * e[C.idx(I1, J1, K0)]);
      double a1 = (1. - wi) * ((1. - wj) * e[C.idx(I0, J0, K1)] + wj * e[C.idx(I0, J1, K1)]) + wi * ((1. - wj) * e[C.idx(I1, J0, K1)] + wj * e[C.idx(I1, J1, K1)]);
      v         = (1. - wk) * a0 + wk * a1;
    } else {
      v = (1. - wi) * ((1. - wj) * e[C.idx(I0, J0, 0)] + wj * e[C.idx(I0, J1, 0)]) + wi * ((1. - wj) * e[C.idx(I1, J0, 0)] + wj * e[C.idx(I1, J1, 0)]);
    }
    fl_store(F.x + F.idx(i, j, kl), F.x[F.idx(i, j, kl)] + v);
  }
};

const double *upload_mg(Solver &s, const std::vector<double> &v)
{
  double *d = (double *)dev_alloc(sizeof(double) * v.size());
  copy_h2d(s.ex, d, v.data(), sizeof(double) * v.size());
  s.ex.sync();
  s.mg_owned.push_back(d);
  return d;
}

void level_tables(Solver &s, MGLevel &L, const std::vector<double> xf[3], const int bc[6])
{
  for (int d = 0; d < 3; ++d) {
    const int           n = L.n[d];
    std::vector<double> h(n), kf(n + 1, 0.), xc(n);
    if (d >= s.dim) {
      h.assign(1, 1.);
      kf.assign(2, 0.);
      L.h[d]  = upload_mg(s, h);
      L.kf[d] = upload_mg(s, kf);
      continue;
    }
    for (int i = 0; i < n; ++i) h[i] = xf[d][i + 1] - xf[d][i], xc[i] = 0.5 * (xf[d][i] + xf[d][i + 1]);
    for (int f = 1; f < n; ++f) kf[f] = 1. / (xc[f] - xc[f - 1]);
    if (L.per[d]) {
      const double len = xf[d][n] - xf[d][0];
      kf[0] = kf[n] = 1. / (xc[0] + len - xc[n - 1]);
    } else {
      if (bc[2 * d] == BC_PRESSURE_OUTLET) kf[0] = 1. / (xc[0] - xf[d][0]);
      if (bc[2 * d + 1] == BC_PRESSURE_OUTLET) kf[n] = 1. / (xf[d][n] - xc[n - 1]);
    }
    L.h[d]  = upload_mg(s, h);
    L.kf[d] = upload_mg(s, kf);
  }
  // constant interior rows of a level that is uniform in every direction
  L.uni = 1;
  double hu[3] = {1., 1., 1.};
  for (int d = 0; d < s.dim; ++d) {
    const int    n = L.n[d];
    const double hbar = (xf[d][n] - xf[d][0]) / n;
    for (int i = 0; i < n; ++i)
      if (std::fabs((xf[d][i + 1] - xf[d][i]) - hbar) > 1e-12 * hbar) L.uni = 0;
    hu[d] = hbar;
  }
  double dsum = 0.;
  for (int d = 0; d < 3; ++d) {
    L.cd[d] = d < s.dim ? (hu[0] * hu[1] * hu[2] / hu[d]) / hu[d] : 0.;
    dsum += 2. * L.cd[d];
  }
  L.idg = 1. / dsum;
}

Box level_box(const MGLevel &L)
{
  Box b = {L.n[0], L.n[1], L.nzl};
  return b;
}

void level_halo(Solver &s, MGLevel &L, double *x)
{
  if (s.dim != 3) return;
  double *f[1] = {x};
  Comm   *c    = L.replicated ? s.local_comm.get() : s.comm.get();
  KTimer  kt(s.ex, KT_HALO, c->nranks > 1 || L.per[2]);
  c->halo(s.ex, f, 1, L.plane, L.nzl, L.per[2] != 0);
}

bool level_tiled(const Solver &s, const MGLevel &L)
{
#ifndef FLUCA_HOSTEMU
  return tma_usable(s) && L.n[0] >= TMX && L.n[1] >= TMY;
#else
  (void)s, (void)L;
  return false;
#endif
}

// one damped-Jacobi sweep; with_dot leaves <b, x_new> in ex.d_result
// Weights of the m sweeps of one smoothing stage.  Plain damped Jacobi (6/7 in 3-D, 0.8 in 2-D) on the coarsest level;
// elsewhere the sweeps are a Chebyshev polynomial in D^-1 P on [0.15 lmax, lmax], lmax = 2 (Gershgorin bound of the
// diagonally dominant flux-form operator): measured on the 32^3 cavity / channel cases: 20-30 % fewer pressure iterations than two damped sweeps,
// at the same cost.  The post-smoother uses the weights in reverse order, which keeps the V-cycle symmetric for CG.
static double sweep_weight(int dim, bool coarsest, int m, int k, bool reverse)
{
  static const bool plain = getenv("FLUCA_B200_MG_JACOBI") != nullptr;
  if (coarsest || plain || m < 2) return dim == 3 ? 6. / 7. : 0.8;
  static const double lmin_env = getenv("FLUCA_B200_MG_LMIN") ? atof(getenv("FLUCA_B200_MG_LMIN")) : 0.3;
  const double lmax = 2., lmin = lmin_env, c = 0.5 * (lmax + lmin), d = 0.5 * (lmax - lmin);
  const int    kk = reverse ? m - 1 - k : k;
  return 1. / (c + d * std::cos(M_PI * (2. * kk + 1.) / (2. * m)));
}

template <int DIM>
void smooth(Solver &s, MGLevel &L, bool zero_guess, bool with_dot, double omega)
{
  if (!zero_guess) level_halo(s, L, L.x);
  KScope        kt(s.ex, KT_MG_SMOOTH);
#ifndef FLUCA_HOSTEMU
  if (DIM == 3 && !zero_guess && level_tiled(s, L)) {
    const double *fields[1] = {L.x};
    if (with_dot) {
      MGSmoothTile<1> op;
      op.L = L, op.omega = omega, op.b = L.b, op.xout = L.t;
      tma_launch(s.ex, op, fields, L.px, L.py, L.nzl + 2, L.n[0], L.n[1], 0, L.nzl, nullptr, L.per[0] != 0);
    } else {
      MGSmoothTile<0> op;
      op.L = L, op.omega = omega, op.b = L.b, op.xout = L.t;
      tma_launch(s.ex, op, fields, L.px, L.py, L.nzl + 2, L.n[0], L.n[1], 0, L.nzl, nullptr, L.per[0] != 0);
    }
    double *tmp = L.x;
    L.x         = L.t;
    L.t         = tmp;
    return;
  }
#endif
  MGSmooth<DIM> f;
  f.L = L, f.omega = omega, f.zero_guess = zero_guess ? 1 : 0, f.xin = L.x, f.b = L.b, f.xout = zero_guess ? L.x : L.t;
  if (with_dot) for_box_reduce<1>(s.ex, level_box(L), f);
  else for_box<2>(s.ex, level_box(L), f);
  if (!zero_guess) {
    double *tmp = L.x;
    L.x         = L.t;
    L.t         = tmp;
  }
}

template <int DIM>
void vcycle(Solver &s, std::vector<MGLevel> &levels, size_t l, bool want_dot)
{
  MGLevel &L = levels[l];
  const bool coarsest = (l + 1 == levels.size());
  if (coarsest && &levels == &s.mg && !s.mg_agg.empty()) {
    // Coarse-grid agglomeration: below this level a slab is a few planes thick and every smoothing sweep would cost a
    // latency-bound halo exchange (two thirds of the ~50 exchanges of a V-cycle sit on levels of <= 64^3 cells).  One
    // allgather puts the level's right-hand side on every rank, the rest of the cycle runs on the whole coarse grid
    // without communication (redundantly: it is tiny), and each rank copies its planes and their ghosts back.
    KScope   kt(s.ex, KT_MG_TRANSFER);
    MGLevel &A0 = s.mg_agg[0];
    {
      KTimer kh(s.ex, KT_HALO, s.comm->nranks > 1);
      s.comm->allgather(s.ex, L.b + L.plane, A0.b + A0.plane, L.plane * L.nzl);
    }
    vcycle<DIM>(s, s.mg_agg, 0, false);
    if (A0.per[2]) level_halo(s, A0, A0.x);
    copy_d2d(s.ex, L.x, A0.x + A0.plane * L.k0, sizeof(double) * (size_t)L.plane * (L.nzl + 2));
    s.ex.stats.launches++;
    return;
  }
  if (coarsest) {
    const int    ns = s.opt.mg_coarse_sweeps;
    const double om = sweep_weight(DIM, true, ns, 0, false);
    smooth<DIM>(s, L, true, want_dot && ns <= 1, om);
    for (int k = 1; k < ns; ++k) smooth<DIM>(s, L, false, want_dot && k == ns - 1, om);
    return;
  }
  int done = 0;
#ifndef FLUCA_HOSTEMU
  const bool no_fuse = getenv("FLUCA_B200_NO_MG_FUSION") != nullptr; // read per call: the parity test toggles it
  if (DIM == 3 && !no_fuse && s.opt.mg_nu1 >= 2 && level_tiled(s, L)) {
    level_halo(s, L, L.b);
    KScope         kt(s.ex, KT_MG_SMOOTH);
    MGFirstTwoTile op;
    op.L = L, op.omega = sweep_weight(DIM, false, s.opt.mg_nu1, 0, false), op.omega2 = sweep_weight(DIM, false, s.opt.mg_nu1, 1, false), op.xout = L.x;
    const double *fields[1] = {L.b};
    tma_launch(s.ex, op, fields, L.px, L.py, L.nzl + 2, L.n[0], L.n[1], 0, L.nzl, nullptr, L.per[0] != 0);
    done = 2;
  }
#endif
  if (!done) smooth<DIM>(s, L, true, false, sweep_weight(DIM, false, s.opt.mg_nu1, 0, false)), done = 1;
  for (int k = done; k < s.opt.mg_nu1; ++k) smooth<DIM>(s, L, false, false, sweep_weight(DIM, false, s.opt.mg_nu1, k, false));
  MGLevel &C = levels[l + 1];
  level_halo(s, L, L.x);
  {
    KScope kt(s.ex, KT_MG_TRANSFER);
    bool   tiled = false;
#ifndef FLUCA_HOSTEMU
    if (DIM == 3 && getenv("FLUCA_B200_NO_MG_FUSION") == nullptr && level_tiled(s, L) && L.cf[0] == 2 && L.cf[1] == 2 && L.cf[2] == 2 && L.nzl % 2 == 0) {
      MGResidTile op;
      op.F = L, op.C = C;
      const double *fields[1] = {L.x};
      tma_launch(s.ex, op, fields, L.px, L.py, L.nzl + 2, L.n[0], L.n[1], 0, L.nzl, nullptr, L.per[0] != 0);
      tiled = true;
    }
#endif
    if (!tiled) {
      MGResidRestrict<DIM> rr;
      rr.F = L, rr.C = C;
      for_box<2>(s.ex, level_box(C), rr);
    }
  }
  vcycle<DIM>(s, levels, l + 1, false);
  const bool gathered = (&levels == &s.mg && !s.mg_agg.empty() && l + 2 == levels.size()); // ghosts came with the copy
  if (!gathered) level_halo(s, C, C.x);
  {
    KScope         kt(s.ex, KT_MG_TRANSFER);
    MGProlong<DIM> pr;
    pr.F = L, pr.C = C;
    for_box<2>(s.ex, level_box(L), pr);
  }
  for (int k = 0; k < s.opt.mg_nu2; ++k) smooth<DIM>(s, L, false, want_dot && k == s.opt.mg_nu2 - 1, sweep_weight(DIM, false, s.opt.mg_nu2, k, true));
}

} // namespace

// builds a hierarchy from level L (sizes, slab and layout filled in) down; nranks = 1 for the replicated hierarchy.
// agg_cells > 0: stop at the first level below the finest with at most that many global cells (it will be gathered).
static void build_levels(Solver &s, std::vector<MGLevel> &levels, MGLevel L, std::vector<double> xf[3], const int bc[6], int nranks, bool equal_slabs, long agg_cells, bool borrow_b0)
{
  for (int lev = 0; lev < 24; ++lev) {
    level_tables(s, L, xf, bc);
    // level 0 of the distributed hierarchy borrows b from the Krylov solver (mg_vcycle); x and the Jacobi scratch t are owned
    const bool own_b = !(borrow_b0 && lev == 0);
    L.own_x = true, L.own_b = own_b;
    L.b = nullptr;
    L.x = (double *)dev_alloc(sizeof(double) * (size_t)L.nalloc);
    s.mg_owned.push_back(L.x);
    if (own_b) {
      L.b = (double *)dev_alloc(sizeof(double) * (size_t)L.nalloc);
      s.mg_owned.push_back(L.b);
    }
    L.t = (double *)dev_alloc(sizeof(double) * (size_t)L.nalloc);
    s.mg_owned.push_back(L.t);
    // coarsening decision (identical on every rank)
    int  cf[3] = {1, 1, 1};
    bool any   = false;
    for (int d = 0; d < s.dim; ++d) {
      bool ok = (L.n[d] % 2 == 0) && (L.n[d] >= 4);
      if (d == 2) ok = (L.n[2] % 2 == 0) && equal_slabs && (L.nzl % 2 == 0) && (nranks > 1 ? L.nzl >= 2 : L.n[2] >= 4);
      if (ok) cf[d] = 2, any = true;
    }
    const bool gather_here = agg_cells > 0 && lev > 0 && (long)L.n[0] * L.n[1] * L.n[2] <= agg_cells;
    if (!any || gather_here) {
      L.cf[0] = L.cf[1] = L.cf[2] = 0;
      levels.push_back(L);
      if (gather_here) {
        // the replicated hierarchy starts from this level's global grid
        MGLevel A = L;
        A.nzl = L.n[2], A.k0 = 0, A.nalloc = A.plane * (A.nzl + 2);
        A.wall_lo_z = A.wall_hi_z = L.per[2] ? 0 : 1;
        A.replicated = 1;
        build_levels(s, s.mg_agg, A, xf, bc, 1, true, 0, false);
      }
      break;
    }
    for (int d = 0; d < 3; ++d) L.cf[d] = cf[d];
    levels.push_back(L);
    // next level
    MGLevel Cn;
    memset(&Cn, 0, sizeof(Cn));
    for (int d = 0; d < 3; ++d) {
      Cn.n[d]   = L.n[d] / cf[d];
      Cn.per[d] = L.per[d];
      if (d < s.dim) {
        std::vector<double> c(Cn.n[d] + 1);
        for (int i = 0; i <= Cn.n[d]; ++i) c[i] = xf[d][(size_t)i * cf[d]];
        xf[d] = c;
      }
    }
    Cn.nzl = L.nzl / cf[2], Cn.k0 = L.k0 / cf[2];
    Cn.px     = ((Cn.n[0] + 7) / 8) * 8;
    Cn.py     = Cn.n[1];
    Cn.plane  = (long)Cn.px * Cn.py;
    Cn.nalloc = Cn.plane * (Cn.nzl + 2);
    Cn.wall_lo_z = L.wall_lo_z, Cn.wall_hi_z = L.wall_hi_z;
    Cn.replicated = L.replicated;
    L = Cn;
  }
}

void mg_setup(Solver &s)
{
  const Geom &g = s.gh.g;
  int         bc[6];
  for (int d = 0; d < 3; ++d) bc[2 * d] = g.t[d].bc_lo, bc[2 * d + 1] = g.t[d].bc_hi;
  std::vector<double> xf[3];
  for (int d = 0; d < s.dim; ++d) xf[d] = s.gh.xf[d];
  const bool equal_slabs = (g.nzl * g.nranks == g.nzg);
  if (!s.local_comm) s.local_comm.reset(new LocalComm());

  MGLevel L;
  memset(&L, 0, sizeof(L));
  L.n[0] = g.nx, L.n[1] = g.ny, L.n[2] = g.nzg;
  L.nzl = g.nzl, L.k0 = g.k0, L.px = g.px, L.py = g.py, L.plane = g.plane, L.nalloc = g.nalloc;
  for (int d = 0; d < 3; ++d) L.per[d] = d < s.dim ? g.t[d].per : 0;
  L.wall_lo_z = g.t[2].wall_lo, L.wall_hi_z = g.t[2].wall_hi;
  // coarse levels of at most 64^3 cells are gathered on every rank (multi-rank 3-D runs with equal slabs)
  long agg = 0;
  if (g.nranks > 1 && s.dim == 3 && equal_slabs) agg = getenv("FLUCA_B200_MG_AGG") ? atol(getenv("FLUCA_B200_MG_AGG")) : 64L * 64 * 64;
  build_levels(s, s.mg, L, xf, bc, g.nranks, equal_slabs, agg, true);
}

void mg_destroy(Solver &s)
{
#ifndef FLUCA_HOSTEMU
  for (Solver::VGraph &g : s.vgraphs)
    if (g.exec) cudaGraphExecDestroy((cudaGraphExec_t)g.exec);
#endif
  s.vgraphs.clear();
  for (void *p : s.mg_owned) dev_free(p);
  s.mg_owned.clear();
  s.mg.clear();
  s.mg_agg.clear();
}

// returns z = V-cycle(r) with zero initial guess; want_dot: the last sweep leaves <r, z> in ex.d_result.  r is a fine-level field (Geom layout); the result
// lives in a level-0 buffer (same layout, ghost planes included) that stays valid until the next call.
static double *mg_vcycle_eager(Solver &s, double *r, bool want_dot)
{
  MGLevel &L0 = s.mg[0];
  L0.b        = r;
  if (s.dim == 2) vcycle<2>(s, s.mg, 0, want_dot);
  else vcycle<3>(s, s.mg, 0, want_dot);
  L0.b = nullptr;
  return L0.x;
}

#ifndef FLUCA_HOSTEMU
// The V-cycle as a CUDA graph.  A cycle is ~100 launches -- most of them a few microseconds long on the coarse levels -- and, on
// several ranks, ~16 grouped NCCL exchanges; enqueueing them costs the host more than the GPU needs to run them (8-GPU strong
// scaling, profiles/r04: 200 communication calls per step at 43 us each, the stream running dry in between).  The cycle is a fixed
// sequence for given buffer roles, so it is captured once per (input field, dot flag, roles of the Jacobi double buffers) and
// replayed with ONE launch.  Not used while per-launch event timing is on (the events would be baked into the graph) or with a
// host-callback communicator.
static void mg_roles(Solver &s, std::vector<double *> &v)
{
  v.clear();
  for (auto *lv : {&s.mg, &s.mg_agg})
    for (MGLevel &L : *lv) v.push_back(L.x), v.push_back(L.t);
}
static void mg_set_roles(Solver &s, const std::vector<double *> &v)
{
  size_t at = 0;
  for (auto *lv : {&s.mg, &s.mg_agg})
    for (MGLevel &L : *lv) L.x = v[at++], L.t = v[at++];
}

static double *mg_vcycle_graph(Solver &s, double *r, bool want_dot)
{
  std::vector<double *> now;
  mg_roles(s, now);
  for (Solver::VGraph &g : s.vgraphs)
    if (g.r == r && g.dot == want_dot && g.pre == now) {
      FL_CUDA(cudaGraphLaunch((cudaGraphExec_t)g.exec, s.ex.stream));
      mg_set_roles(s, g.post);
      s.ex.stats.launches += g.launches;
      return s.mg[0].x;
    }
  if (s.vgraphs.size() >= 32) return mg_vcycle_eager(s, r, want_dot); // roles never settle: stay eager
  // capture: the eager code enqueues into the capturing stream; nothing runs until the launch below
  const long  l0 = s.ex.stats.launches;
  cudaGraph_t graph = nullptr;
  bool        ok = cudaStreamBeginCapture(s.ex.stream, cudaStreamCaptureModeRelaxed) == cudaSuccess;
  if (ok) {
    try {
      (void)mg_vcycle_eager(s, r, want_dot);
    } catch (...) {
      ok = false;
    }
    if (cudaStreamEndCapture(s.ex.stream, &graph) != cudaSuccess || !graph) ok = false;
  }
  cudaGraphExec_t exec = nullptr;
  if (ok && cudaGraphInstantiate(&exec, graph, 0) != cudaSuccess) ok = false;
  if (graph) cudaGraphDestroy(graph);
  if (!ok) {
    // not capturable here: undo the host-side role changes of the aborted pass and run the cycle the ordinary way from now on
    (void)cudaGetLastError();
    mg_set_roles(s, now);
    s.ex.stats.launches = l0;
    s.vgraph_state      = -1;
    return mg_vcycle_eager(s, r, want_dot);
  }
  Solver::VGraph g;
  g.pre = now, g.r = r, g.dot = want_dot, g.exec = (void *)exec, g.launches = s.ex.stats.launches - l0;
  mg_roles(s, g.post);
  s.vgraphs.push_back(g);
  FL_CUDA(cudaGraphLaunch(exec, s.ex.stream));
  return s.mg[0].x;
}
#endif

double *mg_vcycle(Solver &s, double *r, bool want_dot)
{
#ifndef FLUCA_HOSTEMU
  if (s.vgraph_state == 0) {
    static const bool off = getenv("FLUCA_B200_NO_GRAPH") != nullptr;
    s.vgraph_state = (!off && s.comm->capturable() && (!s.local_comm || s.local_comm->capturable())) ? 1 : -1;
  }
  // the first cycles run eagerly (one-time kernel attribute settings and tensor maps happen there)
  // (tests toggle FLUCA_B200_NO_MG_FUSION between calls: a captured cycle would not see it)
  if (s.vgraph_state == 1 && !s.ex.ktime_on && getenv("FLUCA_B200_NO_MG_FUSION") == nullptr && ++s.vcycles > 2) return mg_vcycle_graph(s, r, want_dot);
#endif
  return mg_vcycle_eager(s, r, want_dot);
}

} // namespace fluca
