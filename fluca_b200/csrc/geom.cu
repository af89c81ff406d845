// geom.cu -- host construction of the per-direction coefficient tables (see geom.h).
#include "geom.h"

namespace fluca {

GeomHost::~GeomHost()
{
  for (void *p : owned) dev_free(p);
}

namespace {

struct Dir {
  int                        n, per, bc_lo, bc_hi;
  const std::vector<double> *xf, *xc;
  double                     len;
  double F(int i) const
  {
    if (i < 0) return (*xf)[i + n] - len;
    if (i > n) return (*xf)[i - n] + len;
    return (*xf)[i];
  }
  double C(int i) const
  {
    if (i < 0) return (*xc)[i + n] - len;
    if (i >= n) return (*xc)[i - n] + len;
    return (*xc)[i];
  }
};

const double *upload(GeomHost &gh, Exec &ex, const std::vector<double> &v)
{
  double *d = (double *)dev_alloc(sizeof(double) * v.size());
  copy_h2d(ex, d, v.data(), sizeof(double) * v.size());
  ex.sync(); // v may be a temporary
  gh.owned.push_back(d);
  return d;
}

// zero-gradient extrapolation to the lower wall from cells (0, 1): cartdiscret.c:388-405
void ext_lo(const Dir &D, double w[2])
{
  double h1 = D.C(0) - D.F(0), h2 = D.C(1) - D.F(0), den = (h1 + h2) * (h1 - h2);
  w[0] = -(h2 * h2) / den;
  w[1] = (h1 * h1) / den;
}
// to the upper wall from cells (n-2, n-1): cartdiscret.c:406-423
void ext_hi(const Dir &D, double w[2])
{
  int    n  = D.n;
  double h1 = D.F(n) - D.C(n - 1), h2 = D.F(n) - D.C(n - 2), den = (h1 + h2) * (h1 - h2);
  w[0] = (h1 * h1) / den;
  w[1] = -(h2 * h2) / den;
}

void build_dir(GeomHost &gh, Exec &ex, Tab &T, const Dir &D, int off, int nloc, bool quirk_t_outlet)
{
  const int n = D.n;
  memset(&T, 0, sizeof(Tab));
  T.n = n, T.off = off, T.per = D.per, T.bc_lo = D.bc_lo, T.bc_hi = D.bc_hi;
  T.wall_lo = (!D.per && off == 0);
  T.wall_hi = (!D.per && off + nloc == n);
  if (!D.per && n < 3) throw Error(FL_ERR_ARG, "a non-periodic direction needs at least 3 cells (one-sided wall stencils)");

  std::vector<double> hinv(n), h(n), lapw(2 * 3 * (size_t)n, 0.), grw(3 * (size_t)n, 0.), itw(2 * (size_t)(n + 1), 0.), gstw((size_t)n + 1, 0.);
  for (int i = 0; i < n; ++i) {
    h[i]    = D.F(i + 1) - D.F(i);
    hinv[i] = 1.0 / h[i];
  }
  {
    // uniform direction (what MeshCartSetUniformCoordinates produces, cart.c:458-465): the interior rows of every
    // table are the same numbers up to the rounding of the coordinates, and kernels may use constants instead
    const double hbar = D.len / n;
    double       dev  = 0.;
    for (int i = 0; i < n; ++i) dev = std::fmax(dev, std::fabs(h[i] - hbar));
    T.uni = dev <= 1e-12 * hbar ? 1 : 0;
    T.uh  = hbar;
  }
  // interior rows: cartdiscret.c:210-232 (second derivative), :64-77 (first derivative)
  for (int i = 0; i < n; ++i) {
    bool lo = (!D.per && i == 0), hi = (!D.per && i == n - 1);
    if (lo || hi) continue;
    double h1 = D.C(i) - D.C(i - 1), h2 = D.C(i + 1) - D.C(i), h3 = D.F(i + 1) - D.F(i);
    for (int w = 0; w < 2; ++w) {
      double *r = &lapw[((size_t)w * n + i) * 3];
      r[0]      = 1. / (h1 * h3);
      r[1]      = -(1. / (h1 * h3) + 1. / (h2 * h3));
      r[2]      = 1. / (h2 * h3);
    }
    grw[(size_t)i * 3 + 0] = -1. / (D.C(i + 1) - D.C(i - 1));
    grw[(size_t)i * 3 + 2] = 1. / (D.C(i + 1) - D.C(i - 1));
  }
  // interior faces: cartdiscret.c:373-386 (interpolation), :444-457 (face-normal derivative)
  for (int f = 0; f <= n; ++f) {
    if (!D.per && (f == 0 || f == n)) continue;
    double den            = D.C(f) - D.C(f - 1);
    itw[(size_t)f * 2 + 0] = (D.C(f) - D.F(f)) / den;
    itw[(size_t)f * 2 + 1] = (D.F(f) - D.C(f - 1)) / den;
    gstw[f]               = 1. / den;
  }

  if (!D.per) {
    double e_lo[2], e_hi[2];
    ext_lo(D, e_lo);
    ext_hi(D, e_hi);
    for (int side = 0; side < 2; ++side) {
      const int bc = side ? D.bc_hi : D.bc_lo;
      if (bc != BC_VELOCITY && bc != BC_PRESSURE_OUTLET && bc != BC_SYMMETRY) throw Error(FL_ERR_ARG, "unsupported boundary condition type on a non-periodic boundary");
      for (int w = 0; w < 2; ++w) {
        // Laplacian row of the wall cell, cnlinearcart2d.c:326-380: velocity wall -> Dirichlet one-sided
        // (cartdiscret.c:167-189 / :262-284); outlet -> zero gradient (:191-208 / :286-303);
        // symmetry -> Dirichlet-0 for the normal component, zero gradient for tangential ones
        bool    dirichlet = (bc == BC_VELOCITY) || (bc == BC_SYMMETRY && w == 1);
        double *r         = &lapw[((size_t)w * n + (side ? n - 1 : 0)) * 3];
        if (dirichlet) {
          double h1, h2, h3;
          if (!side) h1 = D.C(0) - D.F(0), h2 = D.C(1) - D.C(0), h3 = D.C(2) - D.C(0);
          else h1 = D.F(n) - D.C(n - 1), h2 = D.C(n - 1) - D.C(n - 2), h3 = D.C(n - 1) - D.C(n - 3);
          double wn = 2. * (h1 - h2 - h3) / (h1 * h2 * h3);               // wall cell
          double w1 = 2. * (h1 - h3) / (h2 * (h1 + h2) * (h2 - h3));      // next
          double w2 = 2. * (h2 - h1) / (h3 * (h1 + h3) * (h2 - h3));      // next but one
          double wb = 2. * (h2 + h3) / (h1 * (h1 + h2) * (h1 + h3));      // wall value, cnlinearcart2d.c:494-498
          if (!side) r[0] = 0., r[1] = wn, r[2] = w1, T.lap_lo2[w] = w2, T.lap_bc_lo[w] = (bc == BC_VELOCITY ? wb : 0.);
          else r[0] = w1, r[1] = wn, r[2] = 0., T.lap_hi2[w] = w2, T.lap_bc_hi[w] = (bc == BC_VELOCITY ? wb : 0.);
        } else {
          double h1 = side ? D.C(n - 1) - D.C(n - 2) : D.C(1) - D.C(0);
          double h2 = side ? D.F(n) - D.F(n - 1) : D.F(1) - D.F(0);
          if (!side) r[0] = 0., r[1] = -1. / (h1 * h2), r[2] = 1. / (h1 * h2);
          else r[0] = 1. / (h1 * h2), r[1] = -1. / (h1 * h2), r[2] = 0.;
        }
        // wall-face value (operators T and B, cnlinearcart2d.c:1078-1123, :1364-1399)
        bool extrap = (bc == BC_PRESSURE_OUTLET) || (bc == BC_SYMMETRY && w == 0);
        for (int q = 0; q < 2; ++q) {
          (side ? T.it_hi : T.it_lo)[w][q] = extrap ? (side ? e_hi[q] : e_lo[q]) : 0.;
          // convection operator at the wall face, cnlinearcart2d.c:655-766.  Lower face: the reference
          // adds the extrapolated flux with a PLUS sign (cartdiscret.c:335-352); our kernels subtract
          // lower-face fluxes, so the table carries the negated weights.  Kept for parity.
          (side ? T.cv1_hi : T.cv1_lo)[w][q] = extrap ? (side ? e_hi[q] : -e_lo[q]) : 0.;
        }
      }
      for (int q = 0; q < 2; ++q) (side ? T.cv2_hi : T.cv2_lo)[q] = (bc == BC_PRESSURE_OUTLET) ? (side ? e_hi[q] : -e_lo[q]) : 0.;
      // operator T at the wall face.  QUIRK of the reference's 3-D file (kept for parity, fluca_b200_desc.no_t_outlet_quirk): at an
      // UPPER pressure outlet it extrapolates with the arguments (centre n-1, face n, slot of the partial element n) where the 2-D
      // file and operator B pass (centre n-2, centre n-1, face n): cnlinearcart3d.c:1996,2055,2114 against cnlinearcart2d.c:1391.
      // That slot holds x_max + h/2 on a mesh from MeshCartSetUniformCoordinates, so the cells (n-2, n-1) get (-1/3, 4/3) in
      // cartdiscret.c:406-423 instead of (-1/8, 9/8).  Found by running the reference's compiled sources (oracle/ref_model).
      for (int q = 0; q < 2; ++q) (side ? T.tn_hi : T.tn_lo)[q] = (side ? T.it_hi : T.it_lo)[1][q];
      if (quirk_t_outlet && side && bc == BC_PRESSURE_OUTLET) T.tn_hi[0] = -1. / 3., T.tn_hi[1] = 4. / 3.;
      (side ? T.it_hi_bc : T.it_lo_bc) = (bc == BC_VELOCITY) ? 1. : 0.;

      // cell-centred pressure gradient of the wall cell, cnlinearcart2d.c:38-79
      double *gr = &grw[(size_t)(side ? n - 1 : 0) * 3];
      if (bc == BC_VELOCITY) { // cartdiscret.c:3-24 / :79-100
        if (!side) {
          double h1 = D.C(1) - D.C(0), h2 = D.C(2) - D.C(0);
          gr[0] = 0., gr[1] = -(h1 + h2) / (h1 * h2), gr[2] = -h2 / (h1 * (h1 - h2)), T.gr_lo2 = h1 / (h2 * (h1 - h2));
        } else {
          double h1 = D.C(n - 1) - D.C(n - 2), h2 = D.C(n - 1) - D.C(n - 3);
          T.gr_hi2 = -h1 / (h2 * (h1 - h2)), gr[0] = h2 / (h1 * (h1 - h2)), gr[1] = (h1 + h2) / (h1 * h2), gr[2] = 0.;
        }
      } else if (bc == BC_PRESSURE_OUTLET) { // cartdiscret.c:26-43 / :102-119, BC weight cnlinearcart2d.c:194-197,219-222
        if (!side) {
          double h1 = D.C(0) - D.F(0), h2 = D.C(1) - D.C(0);
          gr[0] = 0., gr[1] = (h2 - h1) / (h1 * h2), gr[2] = h1 / (h2 * (h1 + h2)), T.gr_bc_lo = -h2 / (h1 * (h1 + h2));
        } else {
          double h1 = D.F(n) - D.C(n - 1), h2 = D.C(n - 1) - D.C(n - 2);
          gr[0] = -h1 / (h2 * (h1 + h2)), gr[1] = (h1 - h2) / (h1 * h2), gr[2] = 0., T.gr_bc_hi = h2 / (h1 * (h1 + h2));
        }
      } else { // symmetry: cartdiscret.c:45-62 / :120-137
        double h1 = side ? D.F(n) - D.C(n - 1) : D.C(0) - D.F(0);
        double h2 = side ? D.C(n - 1) - D.C(n - 2) : D.C(1) - D.C(0);
        double a  = 2. * h1 / (h2 * (2. * h1 + h2));
        if (!side) gr[0] = 0., gr[1] = -a, gr[2] = a;
        else gr[0] = -a, gr[1] = a, gr[2] = 0.;
      }
      // face-normal pressure derivative at an outlet face, cartdiscret.c:425-442 / :459-476, BC weight
      // cnlinearcart2d.c:1835-1838,1860-1863
      if (bc == BC_PRESSURE_OUTLET) {
        if (!side) {
          double h1 = D.C(0) - D.F(0), h2 = D.C(1) - D.F(0);
          T.gst_lo[0] = -h2 / (h1 * (h1 - h2)), T.gst_lo[1] = h1 / (h2 * (h1 - h2)), T.gst_bc_lo = -(h1 + h2) / (h1 * h2);
        } else {
          double h1 = D.F(n) - D.C(n - 1), h2 = D.F(n) - D.C(n - 2);
          T.gst_hi[0] = -h1 / (h2 * (h1 - h2)), T.gst_hi[1] = h2 / (h1 * (h1 - h2)), T.gst_bc_hi = (h1 + h2) / (h1 * h2);
        }
      }
    }
  }
  T.hinv = upload(gh, ex, hinv);
  T.h    = upload(gh, ex, h);
  T.lapw = upload(gh, ex, lapw);
  T.grw  = upload(gh, ex, grw);
  T.itw  = upload(gh, ex, itw);
  T.gstw = upload(gh, ex, gstw);
}

} // namespace

void geom_build(GeomHost &gh, Exec &ex, int dim, const int n[3], const double *const xf[3], const int bc[6], int rank, int nranks, int k0, int nzl, int quirk_t_outlet)
{
  if (dim != 2 && dim != 3) throw Error(FL_ERR_ARG, "dim must be 2 or 3");
  Geom &g = gh.g;
  memset(&g, 0, sizeof(Geom));
  g.dim = dim;
  g.nx = n[0], g.ny = n[1];
  g.nzg    = dim == 3 ? n[2] : 1;
  g.k0     = dim == 3 ? k0 : 0;
  g.nzl    = dim == 3 ? nzl : 1;
  g.px     = ((g.nx + 1 + 7) / 8) * 8;
  g.py     = g.ny + 1;
  g.plane  = (long)g.px * g.py;
  g.nalloc = g.plane * (g.nzl + 2);
  if ((long)g.px * g.py * (long)(g.nzl + 2) >= 2147483647L) throw Error(FL_ERR_ARG, "slab too large for 32-bit element indices: use more z-slabs (GPUs)");
  g.rank = rank, g.nranks = nranks;
  if (dim == 2 && nranks != 1) throw Error(FL_ERR_ARG, "2-D meshes run on one rank (the slab partition is along z)");
  if (g.k0 < 0 || g.nzl < 1 || g.k0 + g.nzl > g.nzg) throw Error(FL_ERR_ARG, "slab outside the mesh");
  for (int d = 0; d < dim; ++d) {
    if (n[d] < 1) throw Error(FL_ERR_ARG, "empty mesh direction");
    gh.xf[d].assign(xf[d], xf[d] + n[d] + 1);
    gh.xc[d].resize(n[d]);
    for (int i = 0; i < n[d]; ++i) {
      if (!(gh.xf[d][i + 1] > gh.xf[d][i])) throw Error(FL_ERR_ARG, "face coordinates must increase");
      gh.xc[d][i] = (gh.xf[d][i] + gh.xf[d][i + 1]) / 2.0; // cart.c:497
    }
    Dir D;
    D.n = n[d], D.bc_lo = bc[2 * d], D.bc_hi = bc[2 * d + 1];
    D.per = (D.bc_lo == BC_PERIODIC);
    if ((D.bc_hi == BC_PERIODIC) != (D.bc_lo == BC_PERIODIC)) throw Error(FL_ERR_ARG, "periodic boundary conditions must be set on both sides of a direction");
    D.xf = &gh.xf[d], D.xc = &gh.xc[d];
    D.len = gh.xf[d][n[d]] - gh.xf[d][0];
    build_dir(gh, ex, g.t[d], D, d == 2 ? g.k0 : 0, d == 2 ? g.nzl : n[d], dim == 3 && quirk_t_outlet);
  }
  if (dim == 3 && nranks > 1 && (g.t[2].wall_lo || g.t[2].wall_hi) && g.nzl < 3) throw Error(FL_ERR_ARG, "a slab that touches a z wall needs at least 3 planes");
}

} // namespace fluca
