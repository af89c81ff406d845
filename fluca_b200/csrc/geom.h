// geom.h -- mesh geometry, HBM data layout and the 1-D coefficient tables of the matrix-free
// operators.
//
// DATA LAYOUT (every field, cell- or face-located, uses the same padded box):
//   index(i, j, kl) = i + px * (j + py * (kl + 1)),   i in [0, px), j in [0, py), kl in [-1, nzl]
//   px = round_up(nx + 1, 8), py = ny + 1: room for the extra RIGHT / UP face of a non-periodic
//   direction (the faces DMStag appends on the last rank); plane kl = -1 and
//   kl = nzl are the ghost planes of the z-slab partition (neighbour rank's data, periodic wrap,
//   or -- for the z-face field on the last rank -- the extra FRONT faces).
//   Cell (i,j,k) owns its LEFT/DOWN/BACK faces, as DMStag does (cart.c:85-120).
//   Padding entries are zero and stay zero (kernels write valid entries only), so flat vector
//   kernels (axpy, dot) can stream whole planes with 16-byte vector accesses.
//
// COEFFICIENT TABLES: all boundary-condition dependence of the stencils (SURVEY.md Appendix A
// table) is folded on the host into small per-direction tables, so kernels branch only on
// "is this the first / last cell of a non-periodic direction".  The formulas are those of
// fluca/src/ns/utils/cartdiscret.c, re-derived here in h1/h2/h3 form; which formula goes where
// follows the BC switches of fluca/src/ns/impl/linearcn/cnlinearcart{2,3}d.c (cited per entry
// in geom.cu).
#pragma once
#include "exec.h"

namespace fluca {

enum { BC_NONE = 0, BC_VELOCITY = 1, BC_PRESSURE_OUTLET = 2, BC_PERIODIC = 3, BC_SYMMETRY = 4 };

// which = 0: component tangential to the direction (c != d); which = 1: normal component (c == d)
struct Tab {
  int n;   // global number of cells in this direction
  int off; // global index of local index 0 (k0 of the slab for z; 0 for x, y)
  int per; // periodic direction
  int bc_lo, bc_hi;
  int wall_lo, wall_hi; // this rank holds the physical lower / upper boundary of the direction
  int uni;              // every cell of the direction has the same width uh (to round-off): interior rows are constants
  double uh;
  // global-index tables (device pointers)
  const double *hinv; // [n]         1 / cell width
  const double *h;    // [n]         cell width
  const double *lapw; // [2][n][3]   second derivative weights on (i-1, i, i+1)
  const double *grw;  // [n][3]      cell-centred first derivative weights on (i-1, i, i+1)
  const double *itw;  // [n+1][2]    linear interpolation to face f from cells (f-1, f); 0 at wall faces
  const double *gstw; // [n+1]       face-normal derivative 1/(xc[f]-xc[f-1]); 0 at wall faces
  // one-sided extras of the first / last cell or face
  double lap_lo2[2], lap_hi2[2];     // weight on cell 2 (row 0) / cell n-3 (row n-1)
  double lap_bc_lo[2], lap_bc_hi[2]; // weight of the wall value (Dirichlet velocity walls)
  double gr_lo2, gr_hi2;             // weight on cell 2 / n-3 (3-cell one-sided gradient at velocity walls)
  double gr_bc_lo, gr_bc_hi;         // weight of the outlet pressure
  double it_lo[2][2], it_hi[2][2];   // wall-face value from cells (0,1) / (n-2,n-1)
  double it_lo_bc, it_hi_bc;         // 1 if the wall face takes the prescribed velocity
  double tn_lo[2], tn_hi[2];         // the same for operator T (normal component): it_*[1] except at a 3-D upper outlet (geom.cu)
  double cv1_lo[2][2], cv1_hi[2][2]; // convection, advected-component interpolation at the wall face
  double cv2_lo[2], cv2_hi[2];       // convection, normal-component interpolation at the wall face
  double gst_lo[2], gst_hi[2];       // outlet face-normal derivative from cells (0,1) / (n-2,n-1)
  double gst_bc_lo, gst_bc_hi;       // weight of the outlet pressure in it
};

struct Geom {
  int  dim;
  int  nx, ny, nzl; // local extents (nx, ny global; nzl planes of the slab)
  int  nzg, k0;     // global z cells, first global plane of the slab
  int  px, py;
  long plane;  // px * py
  long nalloc; // plane * (nzl + 2)
  int  rank, nranks;
  Tab  t[3];

  // 32-bit element index (geom_build refuses slabs with nalloc >= 2^31)
  FL_HD int idx(int i, int j, int kl) const { return i + px * (j + py * (kl + 1)); }
  // neighbour cell index helpers along x / y: periodic wrap, or clamp (the clamped value always
  // meets a zero weight)
  FL_HD int im(int i) const { return i > 0 ? i - 1 : (t[0].per ? nx - 1 : 0); }
  FL_HD int ip(int i) const { return i < nx - 1 ? i + 1 : (t[0].per ? 0 : nx - 1); }
  FL_HD int jm(int j) const { return j > 0 ? j - 1 : (t[1].per ? ny - 1 : 0); }
  FL_HD int jp(int j) const { return j < ny - 1 ? j + 1 : (t[1].per ? 0 : ny - 1); }
  // upper face of cell i along x / y: index i+1, which wraps to face 0 in a periodic direction
  FL_HD int fxp(int i) const { return (t[0].per && i == nx - 1) ? 0 : i + 1; }
  FL_HD int fyp(int j) const { return (t[1].per && j == ny - 1) ? 0 : j + 1; }
};

// host-side owner of the tables
struct GeomHost {
  Geom                 g;
  std::vector<double>  xf[3], xc[3]; // global face / centre coordinates
  std::vector<void *>  owned;        // device allocations
  ~GeomHost();
};

// builds tables for a mesh of n[] cells with faces xf[d][0..n[d]], BC types bc[6] (LEFT, RIGHT,
// DOWN, UP, BACK, FRONT), on slab [k0, k0 + nzl) of rank / nranks
void geom_build(GeomHost &gh, Exec &ex, int dim, const int n[3], const double *const xf[3], const int bc[6], int rank, int nranks, int k0, int nzl, int quirk_t_outlet = 1);

} // namespace fluca
