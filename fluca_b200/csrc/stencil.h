// stencil.h -- matrix-free operators of the linearised Crank-Nicolson / ABF time step.
//
// These functors replace the assembled PETSc matrices of the reference's hot path
// (SURVEY.md 2.2): every operator is evaluated per cell from the fields and the 1-D tables of
// geom.h; nothing is assembled.  Reference call sites replaced (fluca/src/ns/...):
//   AApply       A = I + dt*C(V0, v0interp) - (nu dt/2) L     cnlinearcart3d.c:873-1294, :425-646, :2930-2941
//                (v0interp = B v0 + bc is recomputed on the fly, never stored: :2825-2831)
//   MomentumRhs  r_mom of NSFormFunction                        cnlinearcart3d.c:2967-2991
//   FaceStar     U* = r_int + T v*                              abfpc.c:73-74
//   PoissonRhs   (rho/dt) vol (r_con - D U*)                    abfpc.c:75-76
//   PoissonApply vol * (-D Gst0 p)   (S = -(dt/rho) D Gst0)     abfpc.c:77,151-170
//   Project      v = v* - G~ p, U = U* - G~st p                 abfpc.c:80-101
//   CoupledApply y = M x of the 3x3 block system                nsbasic.c:203-207 (outer KSP MatMult)
#pragma once
#include "geom.h"

namespace fluca {

struct V3 {
  double *c[3];
};
struct CV3 {
  const double *c[3];
  CV3() { c[0] = c[1] = c[2] = nullptr; }
  CV3(const V3 &v) { c[0] = v.c[0], c[1] = v.c[1], c[2] = v.c[2]; }
};

struct StepParams {
  double dt, rho, mu;
  double nu2;   // 0.5 * mu * dt / rho
  double dtrho; // dt / rho
  double sG;    // scale of the outlet-gradient BC vector in r_mom: dt/rho (2-D file) or 1 (3-D file)
};

// boundary values on the device: vel[b][slot] -> [comp * npts + pt], prs[b][slot] -> [pt]
// velocity slots: 0 = t^n, 1 = t^{n+1}; pressure slots: 0 = t_q, 1 = t^{n+1/2}
struct BcDev {
  const double *vel[6][2];
  const double *prs[6][2];
  long          npts[6];
};

// boundary-plane point index of cell (i, j, kl) for boundary b
FL_HD int bc_pt(const Geom &g, int b, int i, int j, int kl)
{
  const int d = b >> 1;
  return d == 0 ? j + g.ny * kl : (d == 1 ? i + g.nx * kl : i + g.nx * j);
}

template <int DIM>
struct Nbr {
  int c;
  int m[DIM], p[DIM];   // cell neighbours -/+ along each direction (wrapped or clamped)
  int fu[DIM];          // index of the upper face along each direction
  int ig[DIM];          // global cell index along each direction
  int m2[DIM], p2[DIM]; // cells at -2 / +2 (only valid where the one-sided stencils need them)
  bool interior;        // no non-periodic wall touches this cell
};

template <int DIM>
FL_HD void nbr(const Geom &g, int i, int j, int kl, Nbr<DIM> &n)
{
  n.c     = g.idx(i, j, kl);
  n.m[0]  = g.idx(g.im(i), j, kl);
  n.p[0]  = g.idx(g.ip(i), j, kl);
  n.fu[0] = g.idx(g.fxp(i), j, kl);
  n.ig[0] = i;
  n.m2[0] = n.c - 2;
  n.p2[0] = n.c + 2;
  n.m[1]  = g.idx(i, g.jm(j), kl);
  n.p[1]  = g.idx(i, g.jp(j), kl);
  n.fu[1] = g.idx(i, g.fyp(j), kl);
  n.ig[1] = j;
  n.m2[1] = n.c - 2 * g.px;
  n.p2[1] = n.c + 2 * g.px;
  bool inter = (g.t[0].per || (i > 0 && i < g.nx - 1)) && (g.t[1].per || (j > 0 && j < g.ny - 1));
  if (DIM == 3) {
    const int pl  = (int)g.plane;
    n.m[DIM - 1]  = n.c - pl; // ghost planes hold the halo (or zeros at a wall)
    n.p[DIM - 1]  = n.c + pl;
    n.fu[DIM - 1] = n.c + pl;
    n.ig[DIM - 1] = g.k0 + kl;
    n.m2[DIM - 1] = n.c - 2 * pl;
    n.p2[DIM - 1] = n.c + 2 * pl;
    inter = inter && (g.t[2].per || (n.ig[DIM - 1] > 0 && n.ig[DIM - 1] < g.t[2].n - 1));
  }
  n.interior = inter;
}

// ------------------------------------------------------------------ momentum operator A
// BND = false compiles the wall handling out: used for warps whose 32 cells are all interior
template <int DIM, bool BND>
FL_HD void a_apply_core(const Geom &g, const StepParams &sp, const BcDev &bc, const CV3 &x, const CV3 &v0, const CV3 &U0, const Nbr<DIM> &nb, int i, int j, int kl, double y[DIM])
{
  double xc[DIM], vc[DIM], conv[DIM], lap[DIM];
#pragma unroll
  for (int c = 0; c < DIM; ++c) {
    xc[c]   = x.c[c][nb.c];
    vc[c]   = v0.c[c][nb.c];
    conv[c] = 0.;
    lap[c]  = 0.;
  }
#pragma unroll
  for (int d = 0; d < DIM; ++d) {
    const Tab   &T  = g.t[d];
    const int    ig = nb.ig[d];
    const bool   lo = BND && (!T.per && ig == 0), hi = BND && (!T.per && ig == T.n - 1);
    const double hh = 0.5 * FL_LDG(T.hinv + ig);
    const double Ul = U0.c[d][nb.c], Uu = U0.c[d][nb.fu[d]];
    const double al = FL_LDG(T.itw + 2 * ig), bl = FL_LDG(T.itw + 2 * ig + 1), au = FL_LDG(T.itw + 2 * ig + 2), bu = FL_LDG(T.itw + 2 * ig + 3);
    double       xm[DIM], xp[DIM], vm[DIM], vp[DIM];
#pragma unroll
    for (int c = 0; c < DIM; ++c) {
      xm[c] = x.c[c][nb.m[d]];
      xp[c] = x.c[c][nb.p[d]];
      vm[c] = v0.c[c][nb.m[d]];
      vp[c] = v0.c[c][nb.p[d]];
    }
    const long pt = (lo || hi) ? bc_pt(g, 2 * d, i, j, kl) : 0;
    // interior rows of the second derivative are the same for normal and tangential components
    const double lw0 = FL_LDG(T.lapw + (size_t)ig * 3), lw1 = FL_LDG(T.lapw + (size_t)ig * 3 + 1), lw2 = FL_LDG(T.lapw + (size_t)ig * 3 + 2);
    // normal-component interpolation of x to the two faces (second convection term)
    const double Ild = lo ? T.cv2_lo[0] * xc[d] + T.cv2_lo[1] * xp[d] : al * xm[d] + bl * xc[d];
    const double Iud = hi ? T.cv2_hi[0] * xm[d] + T.cv2_hi[1] * xc[d] : au * xc[d] + bu * xp[d];
#pragma unroll
    for (int c = 0; c < DIM; ++c) {
      const int w = (c == d) ? 1 : 0;
      // v0interp_c on the two faces = B v0 + wall value at t^n (cnlinearcart3d.c:2825-2831)
      double vbl, vbu;
      if (lo) {
        vbl = T.it_lo[w][0] * vc[c] + T.it_lo[w][1] * vp[c];
        if (T.it_lo_bc != 0.) vbl += bc.vel[2 * d][0][c * bc.npts[2 * d] + pt];
      } else vbl = al * vm[c] + bl * vc[c];
      if (hi) {
        vbu = T.it_hi[w][0] * vm[c] + T.it_hi[w][1] * vc[c];
        if (T.it_hi_bc != 0.) vbu += bc.vel[2 * d + 1][0][c * bc.npts[2 * d + 1] + pt];
      } else vbu = au * vc[c] + bu * vp[c];
      const double Ilc = lo ? T.cv1_lo[w][0] * xc[c] + T.cv1_lo[w][1] * xp[c] : al * xm[c] + bl * xc[c];
      const double Iuc = hi ? T.cv1_hi[w][0] * xm[c] + T.cv1_hi[w][1] * xc[c] : au * xc[c] + bu * xp[c];
      conv[c] += hh * (Uu * Iuc + vbu * Iud - Ul * Ilc - vbl * Ild);
      double l;
      if (lo || hi) {
        const double *lw = T.lapw + ((size_t)w * T.n + ig) * 3;
        l                = lw[0] * xm[c] + lw[1] * xc[c] + lw[2] * xp[c];
        if (lo) l += T.lap_lo2[w] * x.c[c][nb.p2[d]];
        if (hi) l += T.lap_hi2[w] * x.c[c][nb.m2[d]];
      } else l = lw0 * xm[c] + lw1 * xc[c] + lw2 * xp[c];
      lap[c] += l;
    }
  }
#pragma unroll
  for (int c = 0; c < DIM; ++c) y[c] = xc[c] + sp.dt * conv[c] - sp.nu2 * lap[c];
}

template <int DIM>
FL_HD void a_apply_cell(const Geom &g, const StepParams &sp, const BcDev &bc, const CV3 &x, const CV3 &v0, const CV3 &U0, int i, int j, int kl, double y[DIM])
{
  Nbr<DIM> nb;
  nbr<DIM>(g, i, j, kl, nb);
  if (FL_WARP_ALL(nb.interior)) a_apply_core<DIM, false>(g, sp, bc, x, v0, U0, nb, i, j, kl, y);
  else a_apply_core<DIM, true>(g, sp, bc, x, v0, U0, nb, i, j, kl, y);
}

// ------------------------------------------------------------------ diagonal / row sums of A (PCABF DIAG and ROWSUM variants)
// The reference approximates A^-1 by the reciprocal of MatGetDiagonal(A) or MatGetRowSum(A) (abfpc.c:81-94, 151-168).
// Matrix-free: the row of a_apply_core is evaluated on x = e_P (the cell's own entry of component c: ROWSUM = false) or on
// x = 1 (every column: ROWSUM = true) with the same tables; out[c] = 1 / that value.
template <int DIM, bool ROWSUM>
FL_HD void a_row_inverse(const Geom &g, const StepParams &sp, const BcDev &bc, const CV3 &v0, const CV3 &U0, const Nbr<DIM> &nb, int i, int j, int kl, double out[DIM])
{
  const double e = ROWSUM ? 1. : 0.; // value of every column other than the diagonal one
  double       vc[DIM], conv[DIM], lap[DIM];
#pragma unroll
  for (int c = 0; c < DIM; ++c) vc[c] = v0.c[c][nb.c], conv[c] = 0., lap[c] = 0.;
#pragma unroll
  for (int d = 0; d < DIM; ++d) {
    const Tab   &T  = g.t[d];
    const int    ig = nb.ig[d];
    const bool   lo = (!T.per && ig == 0), hi = (!T.per && ig == T.n - 1);
    const double hh = 0.5 * T.hinv[ig];
    const double Ul = U0.c[d][nb.c], Uu = U0.c[d][nb.fu[d]];
    const double al = T.itw[2 * ig], bl = T.itw[2 * ig + 1], au = T.itw[2 * ig + 2], bu = T.itw[2 * ig + 3];
    const long   pt = (lo || hi) ? bc_pt(g, 2 * d, i, j, kl) : 0;
    // weights of (own cell, neighbour) in the normal-component interpolation to the two faces
    const double Ild_c = lo ? T.cv2_lo[0] : bl, Ild_n = lo ? T.cv2_lo[1] : al;
    const double Iud_c = hi ? T.cv2_hi[1] : au, Iud_n = hi ? T.cv2_hi[0] : bu;
#pragma unroll
    for (int c = 0; c < DIM; ++c) {
      const int w = (c == d) ? 1 : 0;
      double    vbl, vbu;
      if (lo) {
        vbl = T.it_lo[w][0] * vc[c] + T.it_lo[w][1] * v0.c[c][nb.p[d]];
        if (T.it_lo_bc != 0.) vbl += bc.vel[2 * d][0][c * bc.npts[2 * d] + pt];
      } else vbl = al * v0.c[c][nb.m[d]] + bl * vc[c];
      if (hi) {
        vbu = T.it_hi[w][0] * v0.c[c][nb.m[d]] + T.it_hi[w][1] * vc[c];
        if (T.it_hi_bc != 0.) vbu += bc.vel[2 * d + 1][0][c * bc.npts[2 * d + 1] + pt];
      } else vbu = au * vc[c] + bu * v0.c[c][nb.p[d]];
      // second convection term: columns of component d; the diagonal column is among them only when d == c
      const double own = (c == d) ? 1. : e;
      const double Ild = own * Ild_c + e * Ild_n, Iud = own * Iud_c + e * Iud_n;
      const double Ilc = lo ? T.cv1_lo[w][0] + e * T.cv1_lo[w][1] : e * al + bl;
      const double Iuc = hi ? e * T.cv1_hi[w][0] + T.cv1_hi[w][1] : au + e * bu;
      conv[c] += hh * (Uu * Iuc + vbu * Iud - Ul * Ilc - vbl * Ild);
      const double *lw = T.lapw + ((size_t)((lo || hi) ? w : 0) * T.n + ig) * 3;
      double        l  = e * lw[0] + lw[1] + e * lw[2];
      if (lo) l += e * T.lap_lo2[w];
      if (hi) l += e * T.lap_hi2[w];
      lap[c] += l;
    }
  }
#pragma unroll
  for (int c = 0; c < DIM; ++c) out[c] = 1. / (1. + sp.dt * conv[c] - sp.nu2 * lap[c]);
}

template <int DIM>
struct AinvCells { // ainv = 1 / diag(A) (type 1) or 1 / rowsum(A) (type 2)
  Geom       g;
  StepParams sp;
  BcDev      bc;
  CV3        v0, U0;
  int        type;
  V3         ainv;
  FL_HD void operator()(int i, int j, int kl) const
  {
    Nbr<DIM> nb;
    nbr<DIM>(g, i, j, kl, nb);
    double r[DIM];
    if (type == 2) a_row_inverse<DIM, true>(g, sp, bc, v0, U0, nb, i, j, kl, r);
    else a_row_inverse<DIM, false>(g, sp, bc, v0, U0, nb, i, j, kl, r);
#pragma unroll
    for (int c = 0; c < DIM; ++c) ainv.c[c][nb.c] = r[c];
  }
};

// y = A x fused with acc[0] += <a, y>, acc[1] += <y, y>, acc[2] += <x, y>, acc[3] += <a, x>
template <int DIM>
struct AApplyDots {
  Geom       g;
  StepParams sp;
  BcDev      bc;
  CV3        x, v0, U0, a;
  V3         y;
  FL_HD void operator()(int i, int j, int kl, double acc[4]) const
  {
    double r[DIM];
    a_apply_cell<DIM>(g, sp, bc, x, v0, U0, i, j, kl, r);
    const long c = g.idx(i, j, kl);
    double     d0 = 0., d1 = 0., d2 = 0., d3 = 0.;
#pragma unroll
    for (int q = 0; q < DIM; ++q) {
      y.c[q][c] = r[q];
      d0 += a.c[q][c] * r[q];
      d1 += r[q] * r[q];
      d2 += x.c[q][c] * r[q];
      d3 += a.c[q][c] * x.c[q][c];
    }
    acc[0] += d0;
    acc[1] += d1;
    acc[2] += d2;
    acc[3] += d3;
  }
};

// cell-centred pressure gradient, unscaled: (G0 q)_c at one cell (cnlinearcart3d.c:4-217)
template <int DIM>
FL_HD void grad_cell(const Geom &g, const double *__restrict__ q, const Nbr<DIM> &nb, double gq[DIM])
{
  const double qc = q[nb.c];
#pragma unroll
  for (int d = 0; d < DIM; ++d) {
    const Tab    &T  = g.t[d];
    const int     ig = nb.ig[d];
    const double *gw = T.grw + (size_t)ig * 3;
    double        s  = gw[0] * q[nb.m[d]] + gw[1] * qc + gw[2] * q[nb.p[d]];
    if (!T.per && ig == 0 && T.gr_lo2 != 0.) s += T.gr_lo2 * q[nb.p2[d]];
    if (!T.per && ig == T.n - 1 && T.gr_hi2 != 0.) s += T.gr_hi2 * q[nb.m2[d]];
    gq[d] = s;
  }
}

// r_mom = v0 + nu2 (L v0 + bcL(t0)) - dt bcC(t0,t1) - (dt/rho G q + sG bcG(tq)) + nu2 bcL(t1)
template <int DIM>
struct MomentumRhs {
  Geom       g;
  StepParams sp;
  BcDev      bc;
  CV3        v0;
  const double *q;
  V3         r;
  FL_HD void operator()(int i, int j, int kl) const
  {
    Nbr<DIM> nb;
    nbr<DIM>(g, i, j, kl, nb);
    if (DIM == 3 && kl + FL_PF < g.nzl) {
      const long ahead = nb.c + FL_PF * g.plane;
#pragma unroll
      for (int c = 0; c < DIM; ++c) fl_prefetch(v0.c[c] + ahead);
      fl_prefetch(q + ahead);
    }
    double vc[DIM], lap[DIM], bcc[DIM], gq[DIM], bcg[DIM];
#pragma unroll
    for (int c = 0; c < DIM; ++c) {
      vc[c]  = v0.c[c][nb.c];
      lap[c] = bcc[c] = bcg[c] = 0.;
    }
    grad_cell<DIM>(g, q, nb, gq);
#pragma unroll
    for (int d = 0; d < DIM; ++d) {
      const Tab &T  = g.t[d];
      const int  ig = nb.ig[d];
      const bool lo = (!T.per && ig == 0), hi = (!T.per && ig == T.n - 1);
#pragma unroll
      for (int c = 0; c < DIM; ++c) {
        const int     w  = (c == d) ? 1 : 0;
        const double *lw = T.lapw + ((size_t)w * T.n + ig) * 3;
        double        l  = lw[0] * v0.c[c][nb.m[d]] + lw[1] * vc[c] + lw[2] * v0.c[c][nb.p[d]];
        if (lo) l += T.lap_lo2[w] * v0.c[c][nb.p2[d]];
        if (hi) l += T.lap_hi2[w] * v0.c[c][nb.m2[d]];
        lap[c] += l;
      }
      for (int side = 0; side < 2; ++side) {
        if (!(side ? hi : lo)) continue;
        const int  b  = 2 * d + side;
        const long pt = bc_pt(g, b, i, j, kl), np = bc.npts[b];
        const int  ty = side ? T.bc_hi : T.bc_lo;
        if (ty == BC_VELOCITY) {
          const double *w0 = bc.vel[b][0], *w1 = bc.vel[b][1];
          const double  lb = side ? T.lap_bc_hi[0] : T.lap_bc_lo[0];
          const double  sg = (side ? 0.5 : -0.5) * T.hinv[ig];
          const double  n0 = w0[d * np + pt], n1 = w1[d * np + pt];
#pragma unroll
          for (int c = 0; c < DIM; ++c) {
            const double a0 = w0[c * np + pt], a1 = w1[c * np + pt];
            lap[c] += lb * (a0 + a1);                 // cnlinearcart3d.c:2984-2985,2990-2991
            bcc[c] += sg * (a1 * n0 + a0 * n1);       // cnlinearcart3d.c:1344,1374,...
          }
        } else if (ty == BC_PRESSURE_OUTLET) {
          bcg[d] += (side ? T.gr_bc_hi : T.gr_bc_lo) * bc.prs[b][0][pt]; // cnlinearcart3d.c:264,292,...
        }
      }
    }
#pragma unroll
    for (int c = 0; c < DIM; ++c) r.c[c][nb.c] = vc[c] + sp.nu2 * lap[c] - sp.dt * bcc[c] - sp.dtrho * gq[c] - sp.sG * bcg[c];
  }
};

// ------------------------------------------------------------------ faces
// value of T(field) at the lower face of cell (i,j,kl) along d; T has no wall-value term
template <int DIM>
FL_HD double t_face_lo(const Geom &g, int d, const double *__restrict__ f, const Nbr<DIM> &nb)
{
  const Tab &T  = g.t[d];
  const int  ig = nb.ig[d];
  if (!T.per && ig == 0) return T.tn_lo[0] * f[nb.c] + T.tn_lo[1] * f[nb.p[d]];
  return T.itw[2 * ig] * f[nb.m[d]] + T.itw[2 * ig + 1] * f[nb.c];
}
// value at the extra (upper wall) face of the last cell
template <int DIM>
FL_HD double t_face_wall_hi(const Geom &g, int d, const double *__restrict__ f, const Nbr<DIM> &nb)
{
  const Tab &T = g.t[d];
  return T.tn_hi[0] * f[nb.m[d]] + T.tn_hi[1] * f[nb.c];
}
// unscaled face-normal pressure derivative (Gst0 p) at the lower face / extra upper wall face
template <int DIM>
FL_HD double gst_face_lo(const Geom &g, int d, const double *__restrict__ p, const Nbr<DIM> &nb)
{
  const Tab &T  = g.t[d];
  const int  ig = nb.ig[d];
  if (!T.per && ig == 0) return T.gst_lo[0] * p[nb.c] + T.gst_lo[1] * p[nb.p[d]];
  return T.gstw[ig] * (p[nb.c] - p[nb.m[d]]);
}
template <int DIM>
FL_HD double gst_face_wall_hi(const Geom &g, int d, const double *__restrict__ p, const Nbr<DIM> &nb)
{
  const Tab &T = g.t[d];
  return T.gst_hi[0] * p[nb.m[d]] + T.gst_hi[1] * p[nb.c];
}

// out_f = a * in_f + b * T(w)_f + c * (Gst0 p)_f on every face this cell owns (LEFT/DOWN/BACK and the
// extra wall faces of a last cell).  Covers:
//   U* = r_int + T v*                 (a=1, b=1, c=0)
//   U  = U* - (dt/rho) Gst0 p         (a=1, b=0, c=-dt/rho)
//   y_U = U - T w + (dt/rho) Gst0 p   (a=1, b=-1, c=dt/rho)
template <int DIM>
struct FaceCombine {
  Geom          g;
  double        a, b, c;
  CV3           in;
  CV3           w;
  const double *p;
  V3            out;
  FL_HD void operator()(int i, int j, int kl) const
  {
    Nbr<DIM> nb;
    nbr<DIM>(g, i, j, kl, nb);
    if (DIM == 3 && kl + FL_PF < g.nzl) {
      const long ahead = nb.c + FL_PF * g.plane;
#pragma unroll
      for (int d = 0; d < DIM; ++d) {
        fl_prefetch(in.c[d] + ahead);
        if (b != 0.) fl_prefetch(w.c[d] + ahead);
      }
      if (c != 0.) fl_prefetch(p + ahead);
    }
#pragma unroll
    for (int d = 0; d < DIM; ++d) {
      const Tab &T = g.t[d];
      double     s = a * in.c[d][nb.c];
      if (b != 0.) s += b * t_face_lo<DIM>(g, d, w.c[d], nb);
      if (c != 0.) s += c * gst_face_lo<DIM>(g, d, p, nb);
      fl_store(out.c[d] + nb.c, s);
      if (!T.per && nb.ig[d] == T.n - 1) {
        const long fw = nb.fu[d];
        double     e  = a * in.c[d][fw];
        if (b != 0.) e += b * t_face_wall_hi<DIM>(g, d, w.c[d], nb);
        if (c != 0.) e += c * gst_face_wall_hi<DIM>(g, d, p, nb);
        fl_store(out.c[d] + fw, e);
      }
    }
  }
};

// divergence in flux form: sum_d area_d (U_hi - U_lo) = vol * (D U)
template <int DIM>
FL_HD double div_flux(const Geom &g, const CV3 &U, const Nbr<DIM> &nb, double &vol)
{
  double h[3] = {1., 1., 1.};
#pragma unroll
  for (int d = 0; d < DIM; ++d) h[d] = g.t[d].h[nb.ig[d]];
  const double area[3] = {h[1] * h[2], h[0] * h[2], h[0] * h[1]};
  vol = h[0] * h[1] * h[2];
  double s = 0.;
#pragma unroll
  for (int d = 0; d < DIM; ++d) s += area[d] * (U.c[d][nb.fu[d]] - U.c[d][nb.c]);
  return s;
}

// out = scale * (vol * rc - flux divergence of U); rc may be null
//   Poisson right-hand side:  scale = rho/dt        (abfpc.c:75-76 in flux form)
//   plain divergence D U:     use DivCell below
template <int DIM>
struct PoissonRhs {
  Geom          g;
  double        scale;
  double        rcscale; // the continuity right-hand side enters as rcscale * rc
  CV3           U;
  const double *rc;
  double       *out;
  FL_HD void operator()(int i, int j, int kl, double acc[1]) const
  {
    Nbr<DIM> nb;
    nbr<DIM>(g, i, j, kl, nb);
    if (DIM == 3 && kl + FL_PF < g.nzl) {
      const long ahead = nb.c + FL_PF * g.plane;
#pragma unroll
      for (int d = 0; d < DIM; ++d) fl_prefetch(U.c[d] + ahead);
      if (rc) fl_prefetch(rc + ahead);
    }
    double vol, fl = div_flux<DIM>(g, U, nb, vol);
    double s = scale * ((rc ? vol * rcscale * rc[nb.c] : 0.) - fl);
    out[nb.c] = s;
    acc[0] += s;
  }
};

template <int DIM>
struct DivCell { // y_p = D U (per unit volume, as the reference's D)
  Geom    g;
  CV3     U;
  double *out;
  FL_HD void operator()(int i, int j, int kl) const
  {
    Nbr<DIM> nb;
    nbr<DIM>(g, i, j, kl, nb);
    if (DIM == 3 && kl + FL_PF < g.nzl) {
      const long ahead = nb.c + FL_PF * g.plane;
#pragma unroll
      for (int d = 0; d < DIM; ++d) fl_prefetch(U.c[d] + ahead);
    }
    double vol, fl = div_flux<DIM>(g, U, nb, vol);
    fl_store(out + nb.c, fl / vol);
  }
};

// ------------------------------------------------------------------ fused stages of the ABF application
// Stage 1 in one pass (abfpc.c:73-76):  U* = a * in + T w  on every face the cell owns, and the Poisson right-hand side
//   out = scale * (vol * rcscale * rc - flux divergence of U*)
// from the cell's own faces: the upper face values are recomputed from (in, w) instead of being re-read after a halo exchange
// of U* (they are the neighbour cell's lower face, formed by the same expression).  Algorithmic traffic in 3-D: in(3) w(3) rc
// -> U*(3) out = 88 B per cell instead of 72 + 40 for FaceCombine followed by PoissonRhs.  acc[0] += out (for the mean removal).
template <int DIM>
struct FaceStarRhs {
  Geom          g;
  double        a;       // scale of `in` (the outer Krylov basis is stored unnormalised)
  double        scale;   // rho / dt
  double        rcscale; // the continuity right-hand side enters as rcscale * rc
  CV3           in;      // r_int
  CV3           w;       // v*
  const double *rc;      // may be null
  V3            Us;      // U*
  double       *out;     // Poisson right-hand side
  FL_HD void operator()(int i, int j, int kl, double acc[1]) const
  {
    Nbr<DIM> nb;
    nbr<DIM>(g, i, j, kl, nb);
    if (DIM == 3 && kl + FL_PF < g.nzl) {
      const long ahead = nb.c + FL_PF * g.plane;
#pragma unroll
      for (int d = 0; d < DIM; ++d) fl_prefetch(in.c[d] + ahead), fl_prefetch(w.c[d] + ahead);
      if (rc) fl_prefetch(rc + ahead);
    }
    double h[3] = {1., 1., 1.};
#pragma unroll
    for (int d = 0; d < DIM; ++d) h[d] = g.t[d].h[nb.ig[d]];
    const double area[3] = {h[1] * h[2], h[0] * h[2], h[0] * h[1]};
    const double vol     = h[0] * h[1] * h[2];
    double       fl      = 0.;
#pragma unroll
    for (int d = 0; d < DIM; ++d) {
      const Tab    &T  = g.t[d];
      const int     ig = nb.ig[d];
      const double *f  = w.c[d];
      double        lo = a * in.c[d][nb.c];
      lo += t_face_lo<DIM>(g, d, f, nb);
      fl_store(Us.c[d] + nb.c, lo);
      const long fw = nb.fu[d];
      double     up = a * in.c[d][fw];
      if (!T.per && ig == T.n - 1) {
        up += t_face_wall_hi<DIM>(g, d, f, nb);
        fl_store(Us.c[d] + fw, up); // the extra wall face of a last cell
      } else {
        // lower face of the next cell (face index wraps to 0 in a periodic direction: itw[n] = itw[0])
        up += T.itw[2 * ig + 2] * f[nb.c] + T.itw[2 * ig + 3] * f[nb.p[d]];
      }
      fl += area[d] * (up - lo);
    }
    const double sv = scale * ((rc ? vol * rcscale * rc[nb.c] : 0.) - fl);
    out[nb.c]       = sv;
    acc[0] += sv;
  }
};

// Stage 2 in one pass (abfpc.c:80-101 with the ID upper factor):  v = v* - (dt/rho) G0 p  and  U = U* - (dt/rho) Gst0 p on every
// face the cell owns.  3-D traffic: v*(3) U*(3) p -> v(3) U(3) = 104 B per cell (the figure of SURVEY.md 8d) in one launch
// instead of ProjectCells (56) + FaceCombine (56).
template <int DIM>
struct ProjectAll {
  Geom          g;
  double        dtrho;
  CV3           vs, Us;
  const double *p;
  V3            v, U;
  FL_HD void operator()(int i, int j, int kl) const
  {
    Nbr<DIM> nb;
    nbr<DIM>(g, i, j, kl, nb);
    if (DIM == 3 && kl + FL_PF < g.nzl) {
      const long ahead = nb.c + FL_PF * g.plane;
#pragma unroll
      for (int c = 0; c < DIM; ++c) fl_prefetch(vs.c[c] + ahead), fl_prefetch(Us.c[c] + ahead);
      fl_prefetch(p + ahead);
    }
    double gp[DIM];
    grad_cell<DIM>(g, p, nb, gp);
#pragma unroll
    for (int d = 0; d < DIM; ++d) {
      const Tab &T = g.t[d];
      fl_store(v.c[d] + nb.c, vs.c[d][nb.c] - dtrho * gp[d]);
      fl_store(U.c[d] + nb.c, Us.c[d][nb.c] - dtrho * gst_face_lo<DIM>(g, d, p, nb));
      if (!T.per && nb.ig[d] == T.n - 1) {
        const long fw = nb.fu[d];
        fl_store(U.c[d] + fw, Us.c[d][fw] - dtrho * gst_face_wall_hi<DIM>(g, d, p, nb));
      }
    }
  }
};

// ------------------------------------------------------------------ Poisson operator, fine level
// (P p)_c = vol * (-D Gst0 p)_c = sum_d area_d (g_lo - g_hi),  g = face-normal derivative
// row of the Poisson operator from the centre value and the two neighbours per direction
template <int DIM>
FL_HD double poisson_row(const Geom &g, const int ig[DIM], double pc, const double pm[DIM], const double pp[DIM])
{
  double h[3] = {1., 1., 1.};
#pragma unroll
  for (int d = 0; d < DIM; ++d) h[d] = FL_LDG(g.t[d].h + ig[d]);
  // face areas by products (a division per direction made this row instruction-bound: profiles/r01h)
  const double area[3] = {h[1] * h[2], h[0] * h[2], h[0] * h[1]};
  double       s = 0.;
#pragma unroll
  for (int d = 0; d < DIM; ++d) {
    const Tab &T = g.t[d];
    const int  i = ig[d];
    double     gl, gu;
    if (!T.per && i == 0) gl = T.gst_lo[0] * pc + T.gst_lo[1] * pp[d];
    else gl = FL_LDG(T.gstw + i) * (pc - pm[d]);
    if (!T.per && i == T.n - 1) gu = T.gst_hi[0] * pm[d] + T.gst_hi[1] * pc;
    else gu = FL_LDG(T.gstw + i + 1) * (pp[d] - pc);
    s += area[d] * (gl - gu);
  }
  return s;
}

template <int DIM>
FL_HD double poisson_apply_cell(const Geom &g, const double *__restrict__ p, const Nbr<DIM> &nb)
{
  double pm[DIM], pp[DIM];
#pragma unroll
  for (int d = 0; d < DIM; ++d) pm[d] = p[nb.m[d]], pp[d] = p[nb.p[d]];
  return poisson_row<DIM>(g, nb.ig, p[nb.c], pm, pp);
}

// v = v* - (dt/rho) G0 p   (abfpc.c:80,95)
template <int DIM>
struct ProjectCells {
  Geom          g;
  double        dtrho;
  CV3           vs;
  const double *p;
  V3            v;
  FL_HD void operator()(int i, int j, int kl) const
  {
    Nbr<DIM> nb;
    nbr<DIM>(g, i, j, kl, nb);
    if (DIM == 3 && kl + FL_PF < g.nzl) {
      const long ahead = nb.c + FL_PF * g.plane;
#pragma unroll
      for (int c = 0; c < DIM; ++c) fl_prefetch(vs.c[c] + ahead);
      fl_prefetch(p + ahead);
    }
    double gp[DIM];
    grad_cell<DIM>(g, p, nb, gp);
#pragma unroll
    for (int c = 0; c < DIM; ++c) fl_store(v.c[c] + nb.c, vs.c[c][nb.c] - dtrho * gp[c]);
  }
};

// w = (dt/rho) G0 p   (first half of the coupled velocity block when the momentum operator runs from TMA tiles)
template <int DIM>
struct GradCells {
  Geom          g;
  double        dtrho;
  const double *p;
  V3            w;
  FL_HD void operator()(int i, int j, int kl) const
  {
    Nbr<DIM> nb;
    nbr<DIM>(g, i, j, kl, nb);
    if (DIM == 3 && kl + FL_PF < g.nzl) fl_prefetch(p + nb.c + FL_PF * g.plane);
    double gp[DIM];
    grad_cell<DIM>(g, p, nb, gp);
#pragma unroll
    for (int c = 0; c < DIM; ++c) fl_store(w.c[c] + nb.c, dtrho * gp[c]);
  }
};

// PCABF with a DIAG / ROWSUM approximation a1 of A^-1 (abfpc.c:81-94, 151-168):
//   w = scale * (1 - a1) .* G0 p            and, when vs is given,   v = vs - scale * a1 .* G0 p  (abfpc.c:80-95)
// w is what the ID variant cancels analytically: T w is the face term left over in S and in the face update.
template <int DIM>
struct GradScaleCells {
  Geom          g;
  double        scale;
  CV3           ainv;
  const double *p;
  CV3           vs; // may be empty
  V3            v;
  V3            w;
  FL_HD void operator()(int i, int j, int kl) const
  {
    Nbr<DIM> nb;
    nbr<DIM>(g, i, j, kl, nb);
    double gp[DIM];
    grad_cell<DIM>(g, p, nb, gp);
#pragma unroll
    for (int c = 0; c < DIM; ++c) {
      const double a1 = ainv.c[c][nb.c], sg = scale * gp[c];
      w.c[c][nb.c]    = (1. - a1) * sg;
      if (vs.c[0]) v.c[c][nb.c] = vs.c[c][nb.c] - a1 * sg;
    }
  }
};

// Schur complement of the DIAG / ROWSUM variants in the scaling of the pressure solve:
//   out = vol * (rho/dt) S' p = P p + vol * D T w,   w = (1 - a1) .* G0 p   (abfpc.c:155-170; S' = D((-T) a1 G~ - (-R)))
// fused with acc[0] += <a, out>
template <int DIM>
struct SchurVariantApplyDot {
  Geom          g;
  const double *p, *a;
  CV3           w;
  double       *out;
  FL_HD void operator()(int i, int j, int kl, double acc[1]) const
  {
    Nbr<DIM> nb;
    nbr<DIM>(g, i, j, kl, nb);
    double s = poisson_apply_cell<DIM>(g, p, nb);
    double h[3] = {1., 1., 1.};
#pragma unroll
    for (int d = 0; d < DIM; ++d) h[d] = g.t[d].h[nb.ig[d]];
    const double area[3] = {h[1] * h[2], h[0] * h[2], h[0] * h[1]};
#pragma unroll
    for (int d = 0; d < DIM; ++d) {
      const Tab    &T  = g.t[d];
      const int     ig = nb.ig[d];
      const double *f  = w.c[d];
      const double  lo = t_face_lo<DIM>(g, d, f, nb);
      double        up;
      if (!T.per && ig == T.n - 1) up = t_face_wall_hi<DIM>(g, d, f, nb);
      else up = T.itw[2 * ig + 2] * f[nb.c] + T.itw[2 * ig + 3] * f[nb.p[d]];
      s += area[d] * (up - lo);
    }
    out[nb.c] = s;
    acc[0] += a[nb.c] * s;
  }
};

// coupled operator, velocity block: y_v = A v + (dt/rho) G0 p ; also w = v + (dt/rho) G0 p (skipped when w is empty)
template <int DIM>
struct CoupledCells {
  Geom          g;
  StepParams    sp;
  BcDev         bc;
  CV3           x, v0, U0;
  const double *p;
  V3            y, w;
  FL_HD void operator()(int i, int j, int kl) const
  {
    Nbr<DIM> nb;
    nbr<DIM>(g, i, j, kl, nb);
    if (DIM == 3 && kl + FL_PF < g.nzl) {
      const long ahead = nb.c + FL_PF * g.plane;
#pragma unroll
      for (int c = 0; c < DIM; ++c) fl_prefetch(x.c[c] + ahead), fl_prefetch(v0.c[c] + ahead), fl_prefetch(U0.c[c] + ahead);
      fl_prefetch(p + ahead);
    }
    double av[DIM], gp[DIM];
    a_apply_cell<DIM>(g, sp, bc, x, v0, U0, i, j, kl, av);
    grad_cell<DIM>(g, p, nb, gp);
#pragma unroll
    for (int c = 0; c < DIM; ++c) {
      fl_store(y.c[c] + nb.c, av[c] + sp.dtrho * gp[c]);
      if (w.c[0]) fl_store(w.c[c] + nb.c, x.c[c][nb.c] + sp.dtrho * gp[c]);
    }
  }
};

} // namespace fluca
