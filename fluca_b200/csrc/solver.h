// solver.h -- device-resident state of one rank's slab and the solver entry points.
#pragma once
#include "geom.h"
#include "stencil.h"
#include "ibm.h"
#include <memory>
#include <string>

namespace fluca {

// ------------------------------------------------------------------ communication (z-slab partition)
// One process per GPU.  The data path needs exactly two collectives:
//   halo   : ghost-plane exchange with the two z neighbours (ncclSend/ncclRecv, grouped)
//   allsum : Krylov dot products / means (ncclAllReduce on a tiny device buffer)
struct Comm {
  int rank = 0, nranks = 1;
  virtual ~Comm() { }
  // fields[f] points at the start of an array laid out as (nzl + 2) planes of `plane` doubles.
  // Sends plane 0 down / plane nzl-1 up, receives into plane nzl / plane -1.
  virtual void halo(Exec &ex, double *const *fields, int nf, long plane, int nzl, bool periodic) = 0;
  // in-place sum over ranks of n doubles in device memory
  virtual void allsum(Exec &ex, double *dev, int n) = 0;
  // gathers `count` doubles from every rank (device), rank-major
  virtual void allgather(Exec &ex, const double *send, double *recv, long count) = 0;
  // exchanges `count` doubles with both z neighbours (device buffers): send_down / recv_down with rank - 1, send_up / recv_up
  // with rank + 1; the ends talk to each other when periodic, else their outer buffers are left alone
  virtual void sendrecv(Exec &ex, const double *send_down, double *recv_down, const double *send_up, double *recv_up, long count, bool periodic) = 0;
  virtual void barrier(Exec &ex) { (void)ex; }
  // every operation of this communicator is an asynchronous enqueue on ex.stream (no host round trip): its calls may be
  // captured into a CUDA graph
  virtual bool capturable() const { return false; }
};

struct LocalComm : public Comm {
  void halo(Exec &ex, double *const *fields, int nf, long plane, int nzl, bool periodic) override;
  void allsum(Exec &, double *, int) override { }
  bool capturable() const override { return true; }
  void allgather(Exec &ex, const double *send, double *recv, long count) override { copy_d2d(ex, recv, send, sizeof(double) * count); }
  void sendrecv(Exec &ex, const double *send_down, double *recv_down, const double *send_up, double *recv_up, long count, bool periodic) override
  {
    if (!periodic) return; // one rank: it is its own neighbour on both sides
    copy_d2d(ex, recv_down, send_up, sizeof(double) * count);
    copy_d2d(ex, recv_up, send_down, sizeof(double) * count);
  }
};

// host-callback communicator: used by the CPU (gloo) tests of the multi-rank logic, and available
// to an MPI host (the PETSc glue) when NCCL is not bootstrapped.  Buffers are host pointers.
typedef int (*fl_halo_cb)(void *ctx, const double *send_down, double *recv_down, const double *send_up, double *recv_up, long count, int periodic);
typedef int (*fl_allsum_cb)(void *ctx, double *vals, int n);
typedef int (*fl_allgather_cb)(void *ctx, const double *send, double *recv, long count);
struct CallbackComm : public Comm {
  fl_halo_cb      halo_cb      = nullptr;
  fl_allsum_cb    allsum_cb    = nullptr;
  fl_allgather_cb allgather_cb = nullptr;
  void           *ctx          = nullptr;
  std::vector<double> hs0, hs1, hr0, hr1; // host staging (CUDA build)
  void halo(Exec &ex, double *const *fields, int nf, long plane, int nzl, bool periodic) override;
  void allsum(Exec &ex, double *dev, int n) override;
  void allgather(Exec &ex, const double *send, double *recv, long count) override;
  void sendrecv(Exec &ex, const double *send_down, double *recv_down, const double *send_up, double *recv_up, long count, bool periodic) override;
};

#ifndef FLUCA_HOSTEMU
Comm *make_nccl_comm(const void *unique_id, int id_bytes, int rank, int nranks);
int   nccl_unique_id(void *out, int bytes); // returns bytes written
#endif

// ------------------------------------------------------------------ multigrid level (cell-centred, flux form)
struct MGLevel {
  int  n[3];          // global cells per direction at this level (n[2] = 1 in 2-D)
  int  nzl, k0;       // local slab
  int  px, py;
  long plane, nalloc;
  int  per[3];
  int  cf[3];         // coarsening factor towards the next level (1 or 2); 0,0,0 on the coarsest
  const double *h[3]; // [n[d]]     cell widths
  const double *kf[3]; // [n[d]+1]  face conductance 1/dist, 0 at Neumann walls, 1/(xc-xw) at outlet walls
  int  wall_lo_z, wall_hi_z;
  int  replicated;    // every rank holds the whole level (agglomerated coarse levels): ghost planes need no exchange
  int  uni;           // every direction has constant cell width: interior rows are cd[d] (2 x - x- - x+), diagonal 1 / idg
  double cd[3], idg;
  double *x, *b, *t;  // solution, right-hand side, scratch (Jacobi double buffer)
  bool own_x, own_b;
  FL_HD long idx(int i, int j, int kl) const { return (long)i + (long)px * ((long)j + (long)py * (long)(kl + 1)); }
};

struct Options {
  int    mode          = 0;     // 0: coupled solve (outer GMRES, PC = ABF)   1: one ABF application
  double outer_rtol    = 1e-5;  // nssol.c:24
  int    outer_maxit   = 100;
  int    outer_restart = 30;
  double mom_rtol      = 1e-5;
  double schur_rtol    = 1e-5;
  int    inner_maxit   = 500;
  int    mg_nu1 = 2, mg_nu2 = 2, mg_coarse_sweeps = 40;
  int    quirk_bcg_scale = 1;
  int    quirk_t_outlet  = 1; // 3-D: operator T at an upper pressure outlet as cnlinearcart3d.c:1996 forms it (geom.cu)
  // approximation of A^-1 inside the ABF factors: 0 ID (default, abfpc.c:328-329), 1 DIAG, 2 ROWSUM (PCABFAinvType, flucans.h:99-103)
  int    schur_ainv = 0, upper_ainv = 0;
};

struct Stats {
  int    outer_its = 0, mom_its = 0, schur_its = 0, abf_applies = 0, converged = 0;
  double outer_rnorm0 = 0., outer_rnorm = 0.;
  int    nhist = 0;
  double hist[128];
  long   launches = 0;
  double mom_last_rel = 0., schur_last_rel = 0.;
  int    inner_unconverged = 0;
};

struct Field {
  double *d = nullptr;
};

// Asynchronous solution view (monitors, NSViewSolution, CGNS output -- nssol.c:130-174, cartvec.c:4-25):
// a copy of the state in library-owned PINNED host memory (compact C-ABI layout), made by a second stream that is ordered
// after the step that produced the state.  The solver stream waits for the copy only before it first overwrites what the
// copy reads (p and p-half at the end of the next step; the velocity buffers are not reused before that), so the download
// of step n overlaps the compute of step n + 1.
struct StateView {
  double *v = nullptr, *U[3] = {nullptr, nullptr, nullptr}, *p = nullptr, *phalf = nullptr;
  bool    pending = false; // a copy is in flight on the view stream
  bool    valid   = false; // the buffers hold (or will hold, once waited for) the state of step_index
  int     step_index = 0;
  double  t = 0.;
#ifndef FLUCA_HOSTEMU
  cudaStream_t stream = nullptr;
  cudaEvent_t  ready = nullptr, done = nullptr;
#endif
};

struct Solver {
  Exec      ex;
  GeomHost  gh;
  Options   opt;
  StepParams sp;
  std::unique_ptr<Comm> comm;
  void (*inner_monitor)(void *, int, int, double) = nullptr; // fluca_b200_set_inner_monitor: host callback per inner residual norm
  void *inner_monitor_ctx = nullptr;
  void  monitor(int which, int it, double rnorm) const { if (inner_monitor) inner_monitor(inner_monitor_ctx, which, it, rnorm); }
  bool      has_outlet = false;
  int       dim = 3;
  // boundary values (device) + descriptor
  BcDev     bc;
  double   *bc_store[6][2][2] = {}; // [b][kind 0 vel / 1 prs][slot]

  std::vector<double *> pool; // every field allocation (freed in destroy)
  // state
  V3      v, U;         // live state (time n; n + 1 once do_step has rotated the buffers)
  V3      v0, U0;       // time-n fields of the prepared step: aliases of v, U (prepare_step)
  V3      vprev, Uprev; // buffers of the previous state / spare of the rotation
  double *p = nullptr, *phalf = nullptr;
  // right-hand side b = (rm, ri, rc) and solve vector x = (xv, xU, xp)
  V3      rm, ri, xv, xU;
  double *rc = nullptr, *xp = nullptr;
  // ABF temporaries
  V3      vstar, Ustar;
  double *srhs = nullptr;
  // DIAG / ROWSUM variants of the ABF factors: 1 / diag(A) or 1 / rowsum(A), rebuilt with A every step (PCSetUp_ABF)
  V3      ainv_store[2] = {};
  V3      ainv_s = {}, ainv_u = {}; // Schur complement / upper triangular factor (the same fields when the types agree)
  // momentum BiCGStab
  V3      kr, krh, kp, kv, ks, kt;
  // Poisson Krylov
  double *pr = nullptr, *pp = nullptr, *pq = nullptr, *ps = nullptr, *pt = nullptr, *prh = nullptr;
  std::vector<MGLevel> mg;     // distributed levels (z-slabs); the last one is gathered when mg_agg is in use
  std::vector<MGLevel> mg_agg; // agglomerated coarse levels: the global grid on every rank, solved redundantly
  std::unique_ptr<Comm> local_comm; // single-rank halo (periodic wrap) of the replicated levels
  std::vector<void *>  mg_owned;
  // CUDA graphs of the V-cycle (mg.cu): one per (input field, dot flag, buffer roles of the Jacobi double buffers)
  struct VGraph {
    std::vector<double *> pre, post; // x / t of every level before and after the cycle
    double               *r;
    bool                  dot;
    void                 *exec; // cudaGraphExec_t
    long                  launches;
  };
  std::vector<VGraph> vgraphs;
  int                 vgraph_state = 0; // 0 untried, 1 in use, -1 disabled (capture failed or not applicable)
  long                vcycles = 0;
  // outer GMRES basis: (restart + 1) vectors of 7 fields, plus work vectors
  std::vector<std::vector<double *>> basis, zbasis; // zbasis[k] = ABF(basis[k]) (flexible GMRES: no final application)
  V3      wv, wU, zv, zU, tw;
  double *wp = nullptr, *zp = nullptr;
  int     basis_size = 0;

  StateView view;
  Ibm    ibm; // immersed-boundary markers (empty unless fluca_b200_set_markers was called)

  Stats  stats;
  int    step_index = 0;
  double t          = 0.;
  bool   prepared   = false;
  bool   rhs_valid  = false;  // (rm, ri, rc) hold the right-hand side of the prepared / last step (with the IBM forcing once do_step ran)
  double tol_floor  = 0.;     // lower bound of the inner relative tolerances, set per ABF application by the outer solver
  bool   allow_guess = true;  // FLUCA_B200_NO_GUESS unsets it (A/B timing)
  bool   have_guess = false; // s.vstar holds a guess of the first momentum solve of the step (unscaled)

  double *alloc_field();
  V3      alloc_v3();
};

// construction
void solver_setup(Solver &s, int dim, const int n[3], const double *const xf[3], const int bc[6], double rho, double mu, double dt, const Options &opt, Comm *comm, int k0, int nzl);
void solver_destroy(Solver &s);

// orders the solver stream after a pending view copy; call before anything overwrites the live state
void view_fence(Solver &s);

// halo helpers
void halo_cells(Solver &s, const V3 &v);
void halo_scalar(Solver &s, double *f);
void halo_faces(Solver &s, const V3 &U);

// reductions: finish the NR sums left in ex.d_result (sum over ranks, copy to host)
void reduce_finish(Solver &s, int n, double *out);

// flat vector extents
long interior_len(const Solver &s);                 // plane * nzl
long interior_off(const Solver &s);                 // plane (skip the lower ghost plane)
long face_len(const Solver &s, int d);              // interior_len (+ one plane for z faces on the last wall rank)

// operators (step.cu)
void prepare_step(Solver &s, double t, int step_index);
void a_apply(Solver &s, const V3 &x, const V3 &y);
// guess: s.vstar holds an initial guess of the momentum solve
// in_scale: the application acts on in_scale * (bm, bi, bcn) (the outer Krylov basis is stored unnormalised)
void abf_apply(Solver &s, const V3 &bm, const V3 &bi, const double *bcn, const V3 &ov, const V3 &oU, double *op, bool guess = false, double in_scale = 1.);
void coupled_apply(Solver &s, const V3 &xv, const V3 &xU, double *xp, const V3 &yv, const V3 &yU, double *yp);
void schur_apply_reference_scaling(Solver &s, double *pin, double *out);
void set_ainv_types(Solver &s, int schur_type, int upper_type); // PCABFSetSchurComplementAinvType / ...UpperTriangular...
// out = vol (rho/dt) S' p for the DIAG / ROWSUM Schur complement, returns <a, out> (uses s.tw)
double schur_variant_apply_dot(Solver &s, double *pin, double *out, const double *a);
int  do_step(Solver &s, double t, int step_index);
double bench_kernel(Solver &s, const std::string &name); // one launch of a named kernel group (tools/kernel_bench.py)

#ifndef FLUCA_HOSTEMU
// TMA-staged versions of the hot 3-D operators (tiles.cu); tma_usable() says whether the mesh qualifies
bool tma_usable(const Solver &s);
void tensor_map_forget(const double *field); // drops the cached tensor maps of a field that is about to be freed
void a_apply_dots_tma(Solver &s, const V3 &x, const V3 &y, const V3 &a, bool with_dots);
void poisson_apply_dot_tma(Solver &s, const double *pin, double *out, const double *a);
// w carries (dt/rho) G p into the tile kernel; keep_w: it leaves as x + (dt/rho) G p (the operand of the face block)
void coupled_cells_tma(Solver &s, const V3 &x, const double *p, const V3 &y, const V3 &w, bool keep_w);
#endif

// Krylov / multigrid (krylov.cu, mg.cu)
int  momentum_solve(Solver &s, const V3 &b, const V3 &x, bool guess = false, double bscale = 1.);
int  poisson_solve(Solver &s, double *b, double *x);
void mg_setup(Solver &s);
void mg_destroy(Solver &s);
double *mg_vcycle(Solver &s, double *r, bool want_dot = false);
void poisson_apply(Solver &s, double *pin, double *out);

} // namespace fluca
