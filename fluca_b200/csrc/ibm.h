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
//   ownership    a rank works on the markers whose support touches its z-slab (a contiguous stretch of the sorted list plus
//                the periodic wrap); the markers whose support crosses a slab face are SHARED with that neighbour
//   interpolate  U_m = sum_cells w_m(cell) v(cell)         one warp per marker, lanes over the support,
//                                                          warp-shuffle reduction, planes of this rank only; the partial sums
//                                                          of the shared markers are exchanged with the two neighbour slabs
//                                                          (one grouped send/recv of a few hundred doubles) and added
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
  const int    *lq;   // [nl] sorted positions of the markers this rank works on
  const int    *lseg; // [nlseg + 1] first entry of lq of every segment
  long          nl;
  int           nlseg;
};

struct Ibm {
  long    n = 0;
  int     npts = 4;
  int     iters = 1; // multi-direct forcing passes per step
  double *coord[3][2] = {}; // device xf / xc per direction
  double *X[3] = {}, *Ud[3] = {}, *Um[3] = {}, *Dl[3] = {}, *F[3] = {}, *dV = nullptr;
  double *Umbuf = nullptr; // the three Um arrays are one allocation (one allreduce)
  int    *perm = nullptr;
  long    cap = 0;
  // slab ownership
  int    *lq = nullptr, *lseg = nullptr; // local list (sorted positions) and its segments
  long    nl = 0;
  int     nlseg = 0;
  bool    sparse = false;          // neighbour exchange of the shared markers (else: every rank walks every marker + one allreduce)
  int    *sh[2] = {nullptr, nullptr}; // marker ids shared with the lower / upper neighbour slab, in sorted order
  long    nsh[2] = {0, 0}, shcap = 0; // shcap: exchange count (the largest shared list of any rank)
  double *xbuf = nullptr;          // [4][shcap * dim]: send down, recv down, send up, recv up
  int    *own = nullptr;           // [n] 1 where this rank reports the marker (assembly of global marker arrays)
  double *gbuf = nullptr;          // [n * dim] scratch of that assembly
  std::vector<void *> owned;
};

void ibm_set_markers(Solver &s, long n, const double *X, const double *Ud, const double *dV, int npts);
void ibm_destroy(Solver &s);
IbmDev ibm_dev(const Solver &s);
// Um = interpolation of the cell field v, complete for every marker this rank works on
void ibm_interpolate(Solver &s, const V3 &v);
// host copy of a per-marker device array set (dim arrays of n doubles) with every marker's value taken from the rank that
// reports it: what the C ABI returns for marker forces / velocities
void ibm_gather_global(Solver &s, double *const src[3], double *host);
// f += spreading of Fm (dim arrays of n doubles, device); f2 (optional) receives the same increment
void ibm_spread(Solver &s, double *const Fm[3], const V3 &f, const V3 *f2 = nullptr);
// steps 2-4 of the coupling: predictor solve, interpolation, forcing added to the momentum RHS
void ibm_force_rhs(Solver &s);

} // namespace fluca
