// ibm.h -- immersed-boundary Lagrangian-Eulerian coupling (discrete-delta interpolation and spreading).
//
// The reference advertises the method (README.md:14) but holds no code for it (THEORY_GUIDE.md:130-132 is
// a TODO), so there is no reference call site to cite: the algorithm is
// defined in oracle/src/ns.c ("immersed boundary" section) and DESIGN.md, and this file is its B200 form.
//
//   markers      X_m (position), Ud_m (prescribed velocity), dV_m (volume weight), replicated on every rank
//   locate       per direction: cell that holds the marker, first support cell, 3 or 4 delta weights
//   sort         markers ordered by the linear index of their first support cell (z-major): neighbours in
//                the list touch the same cache lines; equal keys form SEGMENTS that share all support cells
//   interpolate  U_m = sum_cells w_m(cell) v(cell)         one warp per marker, lanes over the support,
//                                                          warp-shuffle reduction, planes of this rank only,
//                                                          then one allreduce over the ranks
//   spread       f(cell) += sum_m w_m(cell) F_m dV_m/vol   one warp per segment: lanes accumulate their support
//                                                          cells over all markers of the segment in registers,
//                                                          then ONE atomicAdd per support cell and segment
#pragma once
#include "geom.h"

namespace fluca {

struct Solver;
struct V3;

struct IbmDev { // by-value kernel argument
  int           dim, npts;
  long          n;
  int           nc[3]; // global cells per direction (nc[2] = 1 in 2-D)
  int           per[3];
  int           k0, nzl, px, py;
  const double *xf[3], *xc[3]; // global face / centre coordinates (device)
  double        x0[3], len[3];
  const double *X[3], *Ud[3], *dV;
  const int    *perm; // sorted position -> marker
  const int    *seg;  // [nseg + 1] first sorted position of every segment
  int           nseg;
};

struct Ibm {
  long    n = 0;
  int     npts = 4;
  int     iters = 1; // multi-direct forcing passes per step
  double *coord[3][2] = {}; // device xf / xc per direction
  double *X[3] = {}, *Ud[3] = {}, *Um[3] = {}, *Dl[3] = {}, *F[3] = {}, *dV = nullptr;
  double *Umbuf = nullptr; // the three Um arrays are one allocation (one allreduce)
  int    *perm = nullptr, *seg = nullptr;
  int     nseg = 0;
  long    cap = 0;
  std::vector<void *> owned;
};

void ibm_set_markers(Solver &s, long n, const double *X, const double *Ud, const double *dV, int npts);
void ibm_destroy(Solver &s);
IbmDev ibm_dev(const Solver &s);
// Um = interpolation of the cell field v (summed over the ranks)
void ibm_interpolate(Solver &s, const V3 &v);
// f += spreading of Fm (dim arrays of n doubles, device); f2 (optional) receives the same increment
void ibm_spread(Solver &s, double *const Fm[3], const V3 &f, const V3 *f2 = nullptr);
// steps 2-4 of the coupling: predictor solve, interpolation, forcing added to the momentum RHS
void ibm_force_rhs(Solver &s);

} // namespace fluca
