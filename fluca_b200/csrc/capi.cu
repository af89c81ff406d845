// capi.cu -- extern "C" boundary (include/fluca_b200.h).  No torch types, no C++ types, no
// exceptions cross this file: every entry point catches and converts to an error code.
#include "../../include/fluca_b200.h"
#include "solver.h"
#include <string>

using namespace fluca;

struct fluca_b200_comm {
  Comm *c;
};

struct fluca_b200_solver {
  Solver s;
  // snapshot of the state for device-resident benchmarking
  bool    have_snap = false;
  V3      sv, sU;
  double *sp = nullptr, *sph = nullptr;
  int     snap_step = 0;
  // outputs of the operator-level test entry points
  bool    have_api = false;
  V3      av, aU;
  double *ap = nullptr;
};

static thread_local std::string g_err;

#define API_BEGIN try {
#define API_END \
  } \
  catch (const fluca::Error &e) \
  { \
    g_err = e.what(); \
    return e.code; \
  } \
  catch (const std::exception &e) \
  { \
    g_err = e.what(); \
    return FLUCA_B200_ERR_INTERNAL; \
  } \
  return FLUCA_B200_OK;

extern "C" const char *fluca_b200_last_error(void) { return g_err.c_str(); }

extern "C" int fluca_b200_is_host_emulation(void)
{
#ifdef FLUCA_HOSTEMU
  return 1;
#else
  return 0;
#endif
}

extern "C" int fluca_b200_comm_unique_id(void *out, int capacity, int *bytes)
{
  API_BEGIN
#ifndef FLUCA_HOSTEMU
  *bytes = nccl_unique_id(out, capacity);
#else
  (void)out, (void)capacity, (void)bytes;
  throw Error(FL_ERR_NCCL, "NCCL is not part of the host-emulation test build");
#endif
  API_END
}

extern "C" int fluca_b200_comm_create_nccl(const void *unique_id, int bytes, int rank, int nranks, fluca_b200_comm **out)
{
  API_BEGIN
#ifndef FLUCA_HOSTEMU
  fluca_b200_comm *c = new fluca_b200_comm;
  c->c               = make_nccl_comm(unique_id, bytes, rank, nranks);
  *out               = c;
#else
  (void)unique_id, (void)bytes, (void)rank, (void)nranks, (void)out;
  throw Error(FL_ERR_NCCL, "NCCL is not part of the host-emulation test build");
#endif
  API_END
}

extern "C" int fluca_b200_comm_create_callbacks(int rank, int nranks, fluca_b200_halo_fn h, fluca_b200_allsum_fn a, fluca_b200_allgather_fn g, void *ctx, fluca_b200_comm **out)
{
  API_BEGIN
  if (!h || !a) throw Error(FL_ERR_ARG, "halo and allsum callbacks are required");
  // multigrid gathers its coarse levels on every rank (mg.cu): without the callback a multi-rank run would call a null pointer
  if (nranks > 1 && !g) throw Error(FL_ERR_ARG, "the allgather callback is required with more than one rank");
  CallbackComm *cc = new CallbackComm;
  cc->rank = rank, cc->nranks = nranks, cc->halo_cb = h, cc->allsum_cb = a, cc->allgather_cb = g, cc->ctx = ctx;
  fluca_b200_comm *c = new fluca_b200_comm;
  c->c               = cc;
  *out               = c;
  API_END
}

extern "C" int fluca_b200_create(const fluca_b200_desc *d, fluca_b200_comm *comm, fluca_b200_solver **out)
{
  API_BEGIN
  if (!d || !out) throw Error(FL_ERR_ARG, "null argument");
#ifndef FLUCA_HOSTEMU
  {
    int         ndev = 0;
    cudaError_t e    = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev < 1) throw Error(FL_ERR_NODEVICE, "fluca_b200 needs a CUDA device (sm_100a); there is no CPU fallback");
  }
#endif
  Options o;
  o.mode = d->mode;
  if (d->outer_rtol > 0.) o.outer_rtol = d->outer_rtol;
  if (d->outer_maxit > 0) o.outer_maxit = d->outer_maxit;
  if (d->outer_restart > 0) o.outer_restart = d->outer_restart;
  if (d->mom_rtol > 0.) o.mom_rtol = d->mom_rtol;
  if (d->schur_rtol > 0.) o.schur_rtol = d->schur_rtol;
  if (d->inner_maxit > 0) o.inner_maxit = d->inner_maxit;
  if (d->mg_nu1 > 0) o.mg_nu1 = d->mg_nu1;
  if (d->mg_nu2 > 0) o.mg_nu2 = d->mg_nu2;
  if (d->mg_coarse_sweeps > 0) o.mg_coarse_sweeps = d->mg_coarse_sweeps;
  o.quirk_bcg_scale = d->no_bcg_quirk ? 0 : 1;
  o.quirk_t_outlet  = d->no_t_outlet_quirk ? 0 : 1;
  if (o.mode != FLUCA_B200_MODE_COUPLED && o.mode != FLUCA_B200_MODE_FRACTIONAL) throw Error(FL_ERR_ARG, "unknown solve mode");
  fluca_b200_solver *h = new fluca_b200_solver;
  Comm              *c = comm ? comm->c : nullptr;
  delete comm;
  try {
    solver_setup(h->s, d->dim, d->n, d->xf, d->bc_type, d->rho, d->mu, d->dt, o, c, d->k0, d->dim == 3 ? d->nzl : 1);
  } catch (...) {
    try {
      solver_destroy(h->s);
    } catch (...) {
    }
    delete h;
    throw;
  }
  *out = h;
  API_END
}

extern "C" int fluca_b200_destroy(fluca_b200_solver *h)
{
  API_BEGIN
  if (!h) return FLUCA_B200_OK;
  solver_destroy(h->s);
  delete h;
  API_END
}

// ------------------------------------------------------------------ compact <-> padded copies
namespace {
struct Ext {
  int w, hgt, planes;
};
Ext cell_ext(const Geom &g) { return Ext{g.nx, g.ny, g.nzl}; }
Ext face_ext(const Geom &g, int d)
{
  Ext e = cell_ext(g);
  if (d == 0 && !g.t[0].per) e.w += 1;
  if (d == 1 && !g.t[1].per) e.hgt += 1;
  if (d == 2 && g.t[2].wall_hi) e.planes += 1;
  return e;
}
void put(Solver &s, double *dev, const double *host, Ext e)
{
  const Geom &g = s.gh.g;
  copy3d(s.ex, nullptr, 0, dev + g.idx(0, 0, 0), sizeof(double) * g.px, g.py, host, sizeof(double) * e.w, e.hgt, sizeof(double) * e.w, e.hgt, e.planes);
}
void get(Solver &s, double *host, const double *dev, Ext e)
{
  const Geom &g = s.gh.g;
  copy3d(s.ex, nullptr, 1, host, sizeof(double) * e.w, e.hgt, dev + g.idx(0, 0, 0), sizeof(double) * g.px, g.py, sizeof(double) * e.w, e.hgt, e.planes);
}
size_t ext_count(Ext e) { return (size_t)e.w * e.hgt * e.planes; }

void put_cells(Solver &s, const V3 &dst, const double *src)
{
  const Ext e = cell_ext(s.gh.g);
  for (int c = 0; c < s.dim; ++c) put(s, dst.c[c], src + c * ext_count(e), e);
}
void get_cells(Solver &s, double *dst, const V3 &src)
{
  const Ext e = cell_ext(s.gh.g);
  for (int c = 0; c < s.dim; ++c) get(s, dst + c * ext_count(e), src.c[c], e);
}
void fill_stats(const Solver &s, fluca_b200_stats *st)
{
  if (!st) return;
  memset(st, 0, sizeof(*st));
  st->outer_its = s.stats.outer_its, st->mom_its = s.stats.mom_its, st->schur_its = s.stats.schur_its;
  st->abf_applies = s.stats.abf_applies, st->converged = s.stats.converged;
  st->outer_rnorm0 = s.stats.outer_rnorm0, st->outer_rnorm = s.stats.outer_rnorm;
  st->nhist = s.stats.nhist;
  for (int i = 0; i < s.stats.nhist && i < 128; ++i) st->hist[i] = s.stats.hist[i];
  st->launches     = s.stats.launches;
  st->mom_last_rel = s.stats.mom_last_rel, st->schur_last_rel = s.stats.schur_last_rel;
  st->inner_unconverged = s.stats.inner_unconverged;
}
} // namespace

extern "C" int fluca_b200_set_state(fluca_b200_solver *h, const double *v, const double *const U[3], const double *p, const double *phalf)
{
  API_BEGIN
  Solver &s = h->s;
  view_fence(s);
  if (v) put_cells(s, s.v, v);
  if (U)
    for (int d = 0; d < s.dim; ++d)
      if (U[d]) put(s, s.U.c[d], U[d], face_ext(s.gh.g, d));
  if (p) put(s, s.p, p, cell_ext(s.gh.g));
  if (phalf) put(s, s.phalf, phalf, cell_ext(s.gh.g));
  s.prepared = false, s.rhs_valid = false; // a right-hand side formed from the old state no longer belongs to it
  s.ex.sync();
  API_END
}

extern "C" int fluca_b200_get_state(fluca_b200_solver *h, double *v, double *const U[3], double *p, double *phalf)
{
  API_BEGIN
  Solver &s = h->s;
  if (v) get_cells(s, v, s.v);
  if (U)
    for (int d = 0; d < s.dim; ++d)
      if (U[d]) get(s, U[d], s.U.c[d], face_ext(s.gh.g, d));
  if (p) get(s, p, s.p, cell_ext(s.gh.g));
  if (phalf) get(s, phalf, s.phalf, cell_ext(s.gh.g));
  s.ex.sync();
  API_END
}

// ------------------------------------------------------------------ asynchronous solution view (StateView, solver.h)
namespace {
void *pinned_alloc(size_t bytes)
{
  void *p = nullptr;
  if (bytes == 0) bytes = 8;
#ifndef FLUCA_HOSTEMU
  FL_CUDA(cudaMallocHost(&p, bytes));
#else
  p = malloc(bytes);
  if (!p) throw Error(FL_ERR_INTERNAL, "host allocation failed");
#endif
  return p;
}
// compact copy of one padded field to the host on the view stream
void view_get(Solver &s, double *host, const double *dev, Ext e)
{
  const Geom &g = s.gh.g;
#ifndef FLUCA_HOSTEMU
  void *stream = (void *)s.view.stream;
#else
  void *stream = nullptr;
#endif
  copy3d(s.ex, stream, 1, host, sizeof(double) * e.w, e.hgt, dev + g.idx(0, 0, 0), sizeof(double) * g.px, g.py, sizeof(double) * e.w, e.hgt, e.planes);
}
} // namespace

extern "C" int fluca_b200_host_alloc(size_t bytes, void **ptr)
{
  API_BEGIN
  if (!ptr) throw Error(FL_ERR_ARG, "null argument");
  *ptr = pinned_alloc(bytes);
  API_END
}

extern "C" int fluca_b200_host_free(void *ptr)
{
  API_BEGIN
  if (ptr) {
#ifndef FLUCA_HOSTEMU
    FL_CUDA(cudaFreeHost(ptr));
#else
    free(ptr);
#endif
  }
  API_END
}

extern "C" int fluca_b200_stage_state(fluca_b200_solver *h)
{
  API_BEGIN
  if (!h) throw Error(FL_ERR_ARG, "null solver");
  Solver    &s = h->s;
  StateView &w = s.view;
  const Geom &g = s.gh.g;
  // first use: pinned buffers, the view stream and its two events; every piece is guarded on its own so that a failed
  // allocation (8.6 GB of pinned memory at 512^3) leaves a state the next call can complete instead of half a view
  if (!w.v) w.v = (double *)pinned_alloc(sizeof(double) * ext_count(cell_ext(g)) * s.dim);
  for (int d = 0; d < s.dim; ++d)
    if (!w.U[d]) w.U[d] = (double *)pinned_alloc(sizeof(double) * ext_count(face_ext(g, d)));
  if (!w.p) w.p = (double *)pinned_alloc(sizeof(double) * ext_count(cell_ext(g)));
  if (!w.phalf) w.phalf = (double *)pinned_alloc(sizeof(double) * ext_count(cell_ext(g)));
#ifndef FLUCA_HOSTEMU
  if (!w.stream) FL_CUDA(cudaStreamCreateWithFlags(&w.stream, cudaStreamNonBlocking));
  if (!w.ready) FL_CUDA(cudaEventCreateWithFlags(&w.ready, cudaEventDisableTiming));
  if (!w.done) FL_CUDA(cudaEventCreateWithFlags(&w.done, cudaEventDisableTiming));
#endif
#ifndef FLUCA_HOSTEMU
  // after everything already submitted on the solver stream (the step that produced this state) ...
  FL_CUDA(cudaEventRecord(w.ready, s.ex.stream));
  FL_CUDA(cudaStreamWaitEvent(w.stream, w.ready, 0));
#endif
  const Ext ce = cell_ext(g);
  for (int c = 0; c < s.dim; ++c) view_get(s, w.v + c * ext_count(ce), s.v.c[c], ce);
  for (int d = 0; d < s.dim; ++d) view_get(s, w.U[d], s.U.c[d], face_ext(g, d));
  view_get(s, w.p, s.p, ce);
  view_get(s, w.phalf, s.phalf, ce);
#ifndef FLUCA_HOSTEMU
  // ... and before the solver next overwrites what the copy reads (view_fence)
  FL_CUDA(cudaEventRecord(w.done, w.stream));
  w.pending = true;
#endif
  w.valid      = true;
  w.step_index = s.step_index;
  w.t          = s.t;
  API_END
}

extern "C" int fluca_b200_staged_state(fluca_b200_solver *h, const double **v, const double *U[3], const double **p, const double **phalf)
{
  API_BEGIN
  if (!h) throw Error(FL_ERR_ARG, "null solver");
  StateView &w = h->s.view;
  if (!w.valid) throw Error(FL_ERR_ARG, "no staged state: call fluca_b200_stage_state first");
#ifndef FLUCA_HOSTEMU
  FL_CUDA(cudaEventSynchronize(w.done)); // the host waits for the copy only; the solver stream keeps running
#endif
  if (v) *v = w.v;
  if (U)
    for (int d = 0; d < 3; ++d) U[d] = w.U[d];
  if (p) *p = w.p;
  if (phalf) *phalf = w.phalf;
  API_END
}

extern "C" int fluca_b200_set_boundary_velocity(fluca_b200_solver *h, int b, int slot, const double *values)
{
  API_BEGIN
  Solver &s = h->s;
  if (b < 0 || b >= 2 * s.dim || slot < 0 || slot > 1 || !values) throw Error(FL_ERR_ARG, "bad boundary / slot");
  copy_h2d(s.ex, s.bc_store[b][0][slot], values, sizeof(double) * s.bc.npts[b] * s.dim);
  s.ex.sync();
  API_END
}

extern "C" int fluca_b200_set_boundary_pressure(fluca_b200_solver *h, int b, int slot, const double *values)
{
  API_BEGIN
  Solver &s = h->s;
  if (b < 0 || b >= 2 * s.dim || slot < 0 || slot > 1 || !values) throw Error(FL_ERR_ARG, "bad boundary / slot");
  copy_h2d(s.ex, s.bc_store[b][1][slot], values, sizeof(double) * s.bc.npts[b]);
  s.ex.sync();
  API_END
}

extern "C" int fluca_b200_step(fluca_b200_solver *h, double t, int step_index, fluca_b200_stats *st)
{
  API_BEGIN
  int rc = do_step(h->s, t, step_index);
  fill_stats(h->s, st);
  if (rc) throw Error(FL_ERR_DIVERGED, "the linear solve of the time step did not reach its tolerance within the iteration limit");
  API_END
}

extern "C" int fluca_b200_prepare_step(fluca_b200_solver *h, double t, int step_index)
{
  API_BEGIN
  prepare_step(h->s, t, step_index);
  h->s.ex.sync();
  API_END
}

extern "C" int fluca_b200_get_rhs(fluca_b200_solver *h, double *rmom, double *const rint[3], double *rcon)
{
  API_BEGIN
  Solver &s = h->s;
  // valid after fluca_b200_prepare_step, and after fluca_b200_step (then with the immersed-boundary forcing the step added
  // and, without a pressure outlet, the mean of r_con removed: the b the outer solve saw)
  if (!s.rhs_valid) throw Error(FL_ERR_ARG, "call fluca_b200_prepare_step or fluca_b200_step first");
  if (rmom) get_cells(s, rmom, s.rm);
  if (rint)
    for (int d = 0; d < s.dim; ++d)
      if (rint[d]) get(s, rint[d], s.ri.c[d], face_ext(s.gh.g, d));
  if (rcon) get(s, rcon, s.rc, cell_ext(s.gh.g));
  s.ex.sync();
  API_END
}

extern "C" int fluca_b200_apply_momentum(fluca_b200_solver *h, const double *x, double *y)
{
  API_BEGIN
  Solver &s = h->s;
  if (!s.prepared) throw Error(FL_ERR_ARG, "call fluca_b200_prepare_step first");
  put_cells(s, s.kp, x);
  a_apply(s, s.kp, s.kv);
  get_cells(s, y, s.kv);
  s.ex.sync();
  API_END
}

extern "C" int fluca_b200_apply_schur(fluca_b200_solver *h, const double *p, double *y)
{
  API_BEGIN
  Solver &s = h->s;
  put(s, s.pp, p, cell_ext(s.gh.g));
  schur_apply_reference_scaling(s, s.pp, s.pq);
  get(s, y, s.pq, cell_ext(s.gh.g));
  s.ex.sync();
  API_END
}

extern "C" int fluca_b200_apply_vcycle(fluca_b200_solver *h, const double *r, double *z)
{
  API_BEGIN
  Solver &s = h->s;
  put(s, s.pr, r, cell_ext(s.gh.g));
  double *zf = mg_vcycle(s, s.pr);
  get(s, z, zf, cell_ext(s.gh.g));
  s.ex.sync();
  API_END
}

extern "C" int fluca_b200_apply_coupled(fluca_b200_solver *h, const double *xv, const double *const xU[3], const double *xp, double *yv, double *const yU[3], double *yp)
{
  API_BEGIN
  Solver &s = h->s;
  if (!s.prepared) throw Error(FL_ERR_ARG, "call fluca_b200_prepare_step first");
  put_cells(s, s.vstar, xv);
  for (int d = 0; d < s.dim; ++d) put(s, s.Ustar.c[d], xU[d], face_ext(s.gh.g, d));
  put(s, s.srhs, xp, cell_ext(s.gh.g));
  // outputs go to dedicated fields: a cell-located work vector must never receive face data (its padding
  // entries have to stay zero for the flat vector kernels)
  if (!h->have_api) {
    h->av = s.alloc_v3(), h->aU = s.alloc_v3(), h->ap = s.alloc_field();
    h->have_api = true;
  }
  coupled_apply(s, s.vstar, s.Ustar, s.srhs, h->av, h->aU, h->ap);
  get_cells(s, yv, h->av);
  for (int d = 0; d < s.dim; ++d) get(s, yU[d], h->aU.c[d], face_ext(s.gh.g, d));
  get(s, yp, h->ap, cell_ext(s.gh.g));
  s.ex.sync();
  API_END
}

extern "C" int fluca_b200_apply_abf(fluca_b200_solver *h, const double *bv, const double *const bU[3], const double *bp, double *xv, double *const xU[3], double *xp, fluca_b200_stats *st)
{
  API_BEGIN
  Solver &s = h->s;
  if (!s.prepared) throw Error(FL_ERR_ARG, "call fluca_b200_prepare_step first");
  s.stats = Stats();
  // stage the input in the solve-vector slots, which abf_apply does not touch as inputs
  put_cells(s, s.xv, bv);
  for (int d = 0; d < s.dim; ++d) put(s, s.xU.c[d], bU[d], face_ext(s.gh.g, d));
  put(s, s.xp, bp, cell_ext(s.gh.g));
  if (!h->have_api) {
    h->av = s.alloc_v3(), h->aU = s.alloc_v3(), h->ap = s.alloc_field();
    h->have_api = true;
  }
  abf_apply(s, s.xv, s.xU, s.xp, h->av, h->aU, h->ap);
  get_cells(s, xv, h->av);
  for (int d = 0; d < s.dim; ++d) get(s, xU[d], h->aU.c[d], face_ext(s.gh.g, d));
  get(s, xp, h->ap, cell_ext(s.gh.g));
  s.ex.sync();
  fill_stats(s, st);
  API_END
}

// ------------------------------------------------------------------ device-resident helpers
extern "C" int fluca_b200_snapshot_save(fluca_b200_solver *h)
{
  API_BEGIN
  Solver &s = h->s;
  if (!h->have_snap) {
    h->sv = s.alloc_v3(), h->sU = s.alloc_v3(), h->sp = s.alloc_field(), h->sph = s.alloc_field();
    h->have_snap = true;
  }
  const size_t nb = sizeof(double) * (size_t)s.gh.g.nalloc;
  for (int c = 0; c < s.dim; ++c) copy_d2d(s.ex, h->sv.c[c], s.v.c[c], nb), copy_d2d(s.ex, h->sU.c[c], s.U.c[c], nb);
  copy_d2d(s.ex, h->sp, s.p, nb), copy_d2d(s.ex, h->sph, s.phalf, nb);
  s.ex.sync();
  API_END
}

extern "C" int fluca_b200_snapshot_restore(fluca_b200_solver *h)
{
  API_BEGIN
  Solver &s = h->s;
  if (!h->have_snap) throw Error(FL_ERR_ARG, "no snapshot saved");
  view_fence(s);
  const size_t nb = sizeof(double) * (size_t)s.gh.g.nalloc;
  for (int c = 0; c < s.dim; ++c) copy_d2d(s.ex, s.v.c[c], h->sv.c[c], nb), copy_d2d(s.ex, s.U.c[c], h->sU.c[c], nb);
  copy_d2d(s.ex, s.p, h->sp, nb), copy_d2d(s.ex, s.phalf, h->sph, nb);
  s.prepared = false, s.rhs_valid = false;
  s.ex.sync();
  API_END
}

extern "C" int fluca_b200_device_layout(fluca_b200_solver *h, int *px, int *py, long *plane, long *nalloc)
{
  API_BEGIN
  const Geom &g = h->s.gh.g;
  if (px) *px = g.px;
  if (py) *py = g.py;
  if (plane) *plane = g.plane;
  if (nalloc) *nalloc = g.nalloc;
  API_END
}

extern "C" int fluca_b200_device_field(fluca_b200_solver *h, const char *name, void **ptr)
{
  API_BEGIN
  Solver     &s = h->s;
  std::string n(name ? name : "");
  double     *p = nullptr;
  if (n == "p") p = s.p;
  else if (n == "phalf") p = s.phalf;
  else if (n.size() == 2 && (n[0] == 'v' || n[0] == 'U') && n[1] >= '0' && n[1] < '0' + s.dim) p = (n[0] == 'v' ? s.v : s.U).c[n[1] - '0'];
  if (!p) throw Error(FL_ERR_ARG, "unknown field name");
  *ptr = p;
  API_END
}

extern "C" int fluca_b200_stream(fluca_b200_solver *h, void **stream)
{
  API_BEGIN
  *stream = (void *)h->s.ex.stream;
  API_END
}

extern "C" long fluca_b200_launch_count(fluca_b200_solver *h) { return h ? h->s.ex.stats.launches : 0; }

// algorithmic bytes per cell (each distinct array element read once, written once per kernel), fp64, split by the
// kernel classes of the live timing.  3-D: rhs 56, BiCGStab iteration 552 = 2 applies x 96 + 15 vector passes x 24,
// poisson_rhs 32, PCG + V(2,2) iteration 227 = apply 16 + vector updates 72 + V-cycle 139 (smoothing 88 + transfers 34, x 8/7),
// project 104, coupled residual 176, orthogonalisation (2k+3) x 56 at GMRES index k.
static void model_bytes_split(const Solver &s, const fluca_b200_stats *st, double out[KT_NCLASS])
{
  for (int c = 0; c < KT_NCLASS; ++c) out[c] = 0.;
  const Geom   &g   = s.gh.g;
  const double  N   = (double)g.nx * g.ny * g.nzl;
  const bool    d3  = (s.dim == 3);
  const double  rhs = d3 ? 56 : 40, prhs = d3 ? 32 : 24, proj = d3 ? 104 : 80;
  const double  mapply = d3 ? 192 : 128, mvec = d3 ? 360 : 240;
  const double  coup = d3 ? 176 : 128, vec7 = d3 ? 56 : 40;
  const double  lev = d3 ? 8. / 7. : 4. / 3.; // all levels over the fine one
  const double  pvec = 72., smooth = 88. * lev, transfer = 34. * lev;
  // with a pressure outlet the pressure Krylov method is BiCGStab: two applies + two V-cycles per iteration; the DIAG /
  // ROWSUM Schur complement is BiCGStab too and its apply is the two-launch form (3-D: 56 + 48 B instead of 16)
  const double  papply = s.opt.schur_ainv != 0 ? (d3 ? 104. : 80.) : 16.;
  const double  twice  = (s.has_outlet || s.opt.schur_ainv != 0) ? 2. : 1.;
  out[KT_RHS_PROJECT]    = N * (rhs + st->abf_applies * (prhs + proj));
  out[KT_MOMENTUM_APPLY] = N * st->mom_its * mapply;
  out[KT_MOMENTUM_VEC]   = N * st->mom_its * mvec;
  out[KT_POISSON_APPLY]  = N * st->schur_its * twice * papply;
  out[KT_POISSON_VEC]    = N * st->schur_its * twice * pvec;
  out[KT_MG_SMOOTH]      = N * st->schur_its * twice * smooth;
  out[KT_MG_TRANSFER]    = N * st->schur_its * twice * transfer;
  if (s.opt.mode == 0)
    for (int k = 0; k < st->outer_its; ++k) out[KT_OUTER] += N * (coup + (2. * (k % s.opt.outer_restart) + 3.) * vec7);
}

extern "C" double fluca_b200_step_model_bytes(fluca_b200_solver *h, const fluca_b200_stats *st)
{
  if (!h || !st) return 0.;
  double by[KT_NCLASS], B = 0.;
  model_bytes_split(h->s, st, by);
  for (int c = 0; c < KT_NCLASS; ++c) B += by[c];
  return B;
}

extern "C" int fluca_b200_step_model_bytes_split(fluca_b200_solver *h, const fluca_b200_stats *st, double bytes[FLUCA_B200_KT_NCLASS])
{
  API_BEGIN
  if (!h || !st || !bytes) throw Error(FL_ERR_ARG, "null argument");
  model_bytes_split(h->s, st, bytes);
  API_END
}

extern "C" int fluca_b200_kernel_timing(fluca_b200_solver *h, int enable)
{
  API_BEGIN
  h->s.ex.sync();
  h->s.ex.ktime_on = enable != 0;
  API_END
}

extern "C" int fluca_b200_kernel_times(fluca_b200_solver *h, double ms[FLUCA_B200_KT_NCLASS], long counts[FLUCA_B200_KT_NCLASS], int reset)
{
  API_BEGIN
  Exec &ex = h->s.ex;
  ex.ktime_collect();
  for (int c = 0; c < KT_NCLASS; ++c) {
    if (ms) ms[c] = ex.kt_ms[c];
    if (counts) counts[c] = ex.kt_count[c];
    if (reset) ex.kt_ms[c] = 0., ex.kt_count[c] = 0;
  }
  API_END
}

extern "C" int fluca_b200_time_kernel(fluca_b200_solver *h, const char *name, int reps, double *ms, double *bytes)
{
  API_BEGIN
#ifdef FLUCA_HOSTEMU
  (void)h, (void)name, (void)reps, (void)ms, (void)bytes;
  throw Error(FL_ERR_NODEVICE, "kernel timing needs the CUDA build");
#else
  Solver     &s = h->s;
  const Geom &g = s.gh.g;
  std::string n(name ? name : "");
  const double N = (double)g.nx * g.ny * g.nzl;
  if (reps < 1) reps = 1;
  if (!s.prepared) throw Error(FL_ERR_ARG, "kernel timing needs fluca_b200_prepare_step first");
  double per_cell = 0.;
  cudaEvent_t e0, e1;
  FL_CUDA(cudaEventCreate(&e0));
  FL_CUDA(cudaEventCreate(&e1));
  auto run = [&](int k) {
    for (int r = 0; r < k; ++r) {
      if (n == "momentum_apply") a_apply(s, s.kp, s.kv);
      else if (n == "poisson_apply") poisson_apply(s, s.pp, s.pq);
      else if (n == "mg_vcycle") (void)mg_vcycle(s, s.pr);
      else per_cell = bench_kernel(s, n);
    }
  };
  run(2);
  s.ex.sync();
  FL_CUDA(cudaEventRecord(e0, s.ex.stream));
  run(reps);
  FL_CUDA(cudaEventRecord(e1, s.ex.stream));
  FL_CUDA(cudaEventSynchronize(e1));
  float t = 0.f;
  FL_CUDA(cudaEventElapsedTime(&t, e0, e1));
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  *ms = (double)t / reps;
  const bool d3 = (s.dim == 3);
  if (n == "momentum_apply") *bytes = (d3 ? 96. : 64.) * N;
  else if (n == "poisson_apply") *bytes = 16. * N;
  else if (n == "mg_vcycle") *bytes = (d3 ? 122. * 8. / 7. : 122. * 4. / 3.) * N;
  else *bytes = per_cell * N;
#endif
  API_END
}

// ------------------------------------------------------------------ immersed boundary
extern "C" int fluca_b200_set_markers(fluca_b200_solver *h, long n, const double *X, const double *Ud, const double *dV, int delta_points)
{
  API_BEGIN
  ibm_set_markers(h->s, n, X, Ud, dV, delta_points);
  API_END
}

extern "C" int fluca_b200_get_marker_forces(fluca_b200_solver *h, double *F, double *Um)
{
  API_BEGIN
  Solver &s = h->s;
  Ibm    &b = s.ibm;
  if (b.n > 0) {
    // collective with several ranks: every marker's value comes from the rank that reports it
    if (F) ibm_gather_global(s, b.F, F);
    if (Um) ibm_gather_global(s, b.Um, Um);
  }
  API_END
}

extern "C" int fluca_b200_ibm_interpolate(fluca_b200_solver *h, const double *v, double *Um)
{
  API_BEGIN
  Solver &s = h->s;
  if (s.ibm.n <= 0) throw Error(FL_ERR_ARG, "no markers set");
  put_cells(s, s.vstar, v);
  ibm_interpolate(s, s.vstar);
  ibm_gather_global(s, s.ibm.Um, Um);
  API_END
}

extern "C" int fluca_b200_ibm_info(fluca_b200_solver *h, long info[4])
{
  API_BEGIN
  const Ibm &b = h->s.ibm;
  info[0] = b.nl, info[1] = b.nsh[0], info[2] = b.nsh[1], info[3] = b.sparse ? 1 : 0;
  API_END
}

extern "C" int fluca_b200_ibm_spread(fluca_b200_solver *h, const double *Fm, double *f)
{
  API_BEGIN
  Solver &s = h->s;
  Ibm    &b = s.ibm;
  if (b.n <= 0) throw Error(FL_ERR_ARG, "no markers set");
  for (int d = 0; d < s.dim; ++d) {
    copy_h2d(s.ex, b.Dl[d], Fm + (size_t)b.n * d, sizeof(double) * b.n);
    dev_zero(s.ex, s.vstar.c[d], sizeof(double) * (size_t)s.gh.g.nalloc);
  }
  ibm_spread(s, b.Dl, s.vstar);
  get_cells(s, f, s.vstar);
  s.ex.sync();
  API_END
}

extern "C" int fluca_b200_set_abf_ainv_types(fluca_b200_solver *h, int schur_type, int upper_type)
{
  API_BEGIN
  if (!h) throw Error(FL_ERR_ARG, "null solver");
  set_ainv_types(h->s, schur_type, upper_type);
  API_END
}

extern "C" int fluca_b200_set_inner_monitor(fluca_b200_solver *h, fluca_b200_inner_monitor_fn fn, void *ctx)
{
  API_BEGIN
  if (!h) throw Error(FL_ERR_ARG, "null solver");
  h->s.inner_monitor = fn, h->s.inner_monitor_ctx = ctx;
  API_END
}

extern "C" int fluca_b200_set_ibm_iterations(fluca_b200_solver *h, int passes)
{
  API_BEGIN
  if (passes < 1 || passes > 64) throw Error(FL_ERR_ARG, "forcing passes must be in [1, 64]");
  h->s.ibm.iters = passes;
  API_END
}
