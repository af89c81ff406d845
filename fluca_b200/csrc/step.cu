// step.cu -- the NS time step: right-hand side, ABF (fractional-step) application, coupled operator,
// outer GMRES and the step driver.
//
// Host-side control flow mirrors NSStep_CNLinear_Cart3d_Internal (cnlinearcart3d.c:2807-2863),
// NSFormFunction (:2945-3043), PCApply_ABF (abfpc.c:48-111) and the outer KSP the base class sets
// up (nssol.c:13-30: GMRES, right-preconditioned because the norm is unpreconditioned, zero guess).
#include "solver.h"

namespace fluca {

double *Solver::alloc_field()
{
  double *d = (double *)dev_alloc(sizeof(double) * (size_t)gh.g.nalloc);
  pool.push_back(d);
  return d;
}
V3 Solver::alloc_v3()
{
  V3 r;
  r.c[0] = alloc_field();
  r.c[1] = alloc_field();
  r.c[2] = (dim == 3) ? alloc_field() : nullptr;
  return r;
}

long interior_len(const Solver &s) { return s.gh.g.plane * s.gh.g.nzl; }
long interior_off(const Solver &s) { return s.gh.g.plane; }
long face_len(const Solver &s, int d)
{
  const Geom &g = s.gh.g;
  return g.plane * g.nzl + ((d == 2 && g.t[2].wall_hi) ? g.plane : 0);
}

// ------------------------------------------------------------------ halos
void LocalComm::halo(Exec &ex, double *const *fields, int nf, long plane, int nzl, bool periodic)
{
  if (!periodic) return;
  for (int f = 0; f < nf; ++f) {
    double *a = fields[f];
    copy_d2d(ex, a, a + plane * nzl, sizeof(double) * plane);             // ghost -1 <- plane nzl-1
    copy_d2d(ex, a + plane * (nzl + 1), a + plane, sizeof(double) * plane); // ghost nzl <- plane 0
  }
  ex.stats.launches += 2 * nf;
}

void CallbackComm::halo(Exec &ex, double *const *fields, int nf, long plane, int nzl, bool periodic)
{
  if (nranks == 1 && !periodic) return;
  for (int f = 0; f < nf; ++f) {
    double *a = fields[f];
#ifdef FLUCA_HOSTEMU
    int rc = halo_cb(ctx, a + plane, a, a + plane * nzl, a + plane * (nzl + 1), plane, periodic ? 1 : 0);
#else
    hs0.resize(plane), hs1.resize(plane), hr0.resize(plane), hr1.resize(plane);
    copy_d2h(ex, hs0.data(), a + plane, sizeof(double) * plane);
    copy_d2h(ex, hs1.data(), a + plane * nzl, sizeof(double) * plane);
    copy_d2h(ex, hr0.data(), a, sizeof(double) * plane);
    copy_d2h(ex, hr1.data(), a + plane * (nzl + 1), sizeof(double) * plane);
    ex.sync();
    int rc = halo_cb(ctx, hs0.data(), hr0.data(), hs1.data(), hr1.data(), plane, periodic ? 1 : 0);
    copy_h2d(ex, a, hr0.data(), sizeof(double) * plane);
    copy_h2d(ex, a + plane * (nzl + 1), hr1.data(), sizeof(double) * plane);
    ex.sync();
#endif
    if (rc) throw Error(FL_ERR_INTERNAL, "halo callback failed");
  }
}
void CallbackComm::allsum(Exec &ex, double *dev, int n)
{
  if (nranks == 1) return;
#ifdef FLUCA_HOSTEMU
  if (allsum_cb(ctx, dev, n)) throw Error(FL_ERR_INTERNAL, "allsum callback failed");
#else
  std::vector<double> h(n);
  copy_d2h(ex, h.data(), dev, sizeof(double) * n);
  ex.sync();
  if (allsum_cb(ctx, h.data(), n)) throw Error(FL_ERR_INTERNAL, "allsum callback failed");
  copy_h2d(ex, dev, h.data(), sizeof(double) * n);
  ex.sync();
#endif
}
void CallbackComm::allgather(Exec &ex, const double *send, double *recv, long count)
{
#ifdef FLUCA_HOSTEMU
  if (allgather_cb(ctx, send, recv, count)) throw Error(FL_ERR_INTERNAL, "allgather callback failed");
#else
  std::vector<double> hs(count), hr(count * nranks);
  copy_d2h(ex, hs.data(), send, sizeof(double) * count);
  ex.sync();
  if (allgather_cb(ctx, hs.data(), hr.data(), count)) throw Error(FL_ERR_INTERNAL, "allgather callback failed");
  copy_h2d(ex, recv, hr.data(), sizeof(double) * count * nranks);
  ex.sync();
#endif
}

void CallbackComm::sendrecv(Exec &ex, const double *send_down, double *recv_down, const double *send_up, double *recv_up, long count, bool periodic)
{
  if (count <= 0) return;
  if (nranks == 1 && !periodic) return;
#ifdef FLUCA_HOSTEMU
  int rc = halo_cb(ctx, send_down, recv_down, send_up, recv_up, count, periodic ? 1 : 0);
#else
  hs0.resize(count), hs1.resize(count), hr0.resize(count), hr1.resize(count);
  copy_d2h(ex, hs0.data(), send_down, sizeof(double) * count);
  copy_d2h(ex, hs1.data(), send_up, sizeof(double) * count);
  copy_d2h(ex, hr0.data(), recv_down, sizeof(double) * count);
  copy_d2h(ex, hr1.data(), recv_up, sizeof(double) * count);
  ex.sync();
  int rc = halo_cb(ctx, hs0.data(), hr0.data(), hs1.data(), hr1.data(), count, periodic ? 1 : 0);
  copy_h2d(ex, recv_down, hr0.data(), sizeof(double) * count);
  copy_h2d(ex, recv_up, hr1.data(), sizeof(double) * count);
  ex.sync();
#endif
  if (rc) throw Error(FL_ERR_INTERNAL, "halo callback failed");
}

void halo_cells(Solver &s, const V3 &v)
{
  const Geom &g = s.gh.g;
  if (g.dim != 3) return;
  double *f[3] = {v.c[0], v.c[1], v.c[2]};
  KTimer  kt(s.ex, KT_HALO, s.comm->nranks > 1 || g.t[2].per);
  s.comm->halo(s.ex, f, 3, g.plane, g.nzl, g.t[2].per != 0);
}
void halo_scalar(Solver &s, double *p)
{
  const Geom &g = s.gh.g;
  if (g.dim != 3) return;
  double *f[1] = {p};
  KTimer  kt(s.ex, KT_HALO, s.comm->nranks > 1 || g.t[2].per);
  s.comm->halo(s.ex, f, 1, g.plane, g.nzl, g.t[2].per != 0);
}
void halo_faces(Solver &s, const V3 &U)
{
  // only the z-face field crosses slab boundaries (upper face of the last plane = BACK face of the
  // next rank's first plane).  Its lower ghost plane is never read.
  const Geom &g = s.gh.g;
  if (g.dim != 3) return;
  double *f[1] = {U.c[2]};
  // on the last wall rank plane nzl holds owned FRONT faces: a non-periodic exchange never writes it
  KTimer kt(s.ex, KT_HALO, s.comm->nranks > 1 || g.t[2].per);
  s.comm->halo(s.ex, f, 1, g.plane, g.nzl, g.t[2].per != 0);
}

#ifndef FLUCA_HOSTEMU
__global__ void k_publish_sums(const double *__restrict__ src, double *__restrict__ dst, int n)
{
  if ((int)threadIdx.x < n) dst[threadIdx.x] = src[threadIdx.x];
}
#endif

void reduce_finish(Solver &s, int n, double *out)
{
  {
    KTimer kt(s.ex, KT_HALO, s.comm->nranks > 1); // communication class of the live timing: halos, allreduces, allgathers
    s.comm->allsum(s.ex, s.ex.d_result, n);
  }
  // The sums reach the host through a one-warp kernel that stores them into mapped pinned memory, not through a device-to-host
  // memcpy: a memcpy queues on the copy engine behind whatever another stream has put there -- the 8.6 GB solution view of
  // fluca_b200_stage_state at 512^3 (section 8c of DESIGN.md) -- and ~70 reductions per step would each wait for it.
#ifndef FLUCA_HOSTEMU
  if (s.ex.h_result_dev) {
    k_publish_sums<<<1, 32, 0, s.ex.stream>>>(s.ex.d_result, s.ex.h_result_dev, n);
    FL_CUDA(cudaGetLastError());
    s.ex.stats.launches++;
  } else
#endif
    copy_d2h(s.ex, s.ex.h_result, s.ex.d_result, sizeof(double) * n);
  s.ex.sync();
  for (int i = 0; i < n; ++i) out[i] = s.ex.h_result[i];
}

void view_fence(Solver &s)
{
  if (!s.view.pending) return;
#ifndef FLUCA_HOSTEMU
  FL_CUDA(cudaStreamWaitEvent(s.ex.stream, s.view.done, 0));
#endif
  s.view.pending = false;
}

// ------------------------------------------------------------------ setup
void solver_setup(Solver &s, int dim, const int n[3], const double *const xf[3], const int bcin[6], double rho, double mu, double dt, const Options &opt, Comm *comm, int k0, int nzl)
{
  s.ex.init();
  s.dim = dim;
  s.opt = opt;
  s.comm.reset(comm ? comm : new LocalComm());
  int bc[6];
  for (int b = 0; b < 6; ++b) bc[b] = (b < 2 * dim) ? bcin[b] : BC_NONE;
  geom_build(s.gh, s.ex, dim, n, xf, bc, s.comm->rank, s.comm->nranks, k0, nzl, opt.quirk_t_outlet);
  const Geom &g = s.gh.g;
  s.has_outlet = false;
  for (int b = 0; b < 2 * dim; ++b)
    if (bc[b] == BC_PRESSURE_OUTLET) s.has_outlet = true; // nsbasic.c:215-244
  if (!(rho > 0.) || !(dt > 0.) || !(mu >= 0.)) throw Error(FL_ERR_ARG, "density and time step must be positive, viscosity non-negative");
  s.sp.dt = dt, s.sp.rho = rho, s.sp.mu = mu;
  s.sp.nu2   = 0.5 * mu * dt / rho;
  s.sp.dtrho = dt / rho;
  // cnlinearcart2d.c:2104,2109 scale by dt/rho; cnlinearcart3d.c:2977,2981 by 1
  s.sp.sG = (dim == 3 && opt.quirk_bcg_scale) ? 1.0 : dt / rho;

  memset(&s.bc, 0, sizeof(BcDev));
  for (int b = 0; b < 2 * dim; ++b) {
    const int d = b / 2;
    long      np = d == 0 ? (long)g.ny * g.nzl : (d == 1 ? (long)g.nx * g.nzl : (long)g.nx * g.ny);
    s.bc.npts[b] = np;
    for (int slot = 0; slot < 2; ++slot) {
      s.bc_store[b][0][slot] = (double *)dev_alloc(sizeof(double) * np * 3);
      s.bc_store[b][1][slot] = (double *)dev_alloc(sizeof(double) * np);
      s.bc.vel[b][slot]      = s.bc_store[b][0][slot];
      s.bc.prs[b][slot]      = s.bc_store[b][1][slot];
    }
  }

  s.v = s.alloc_v3(), s.U = s.alloc_v3(), s.vprev = s.alloc_v3(), s.Uprev = s.alloc_v3();
  s.v0 = s.v, s.U0 = s.U;
  s.p = s.alloc_field(), s.phalf = s.alloc_field();
  s.rm = s.alloc_v3(), s.ri = s.alloc_v3(), s.rc = s.alloc_field();
  s.xv = s.alloc_v3(), s.xU = s.alloc_v3(), s.xp = s.alloc_field();
  s.vstar = s.alloc_v3(), s.Ustar = s.alloc_v3(), s.srhs = s.alloc_field();
  // Krylov work space.  BASELINE config 5 (2.2 GB per field and GPU) only fits the coupled mode if work vectors whose lives
  // never overlap share storage:
  //  * BiCGStab's s overwrites r in place (s = r - alpha v, then r = s - omega t): five vectors, not six;
  //  * the pressure Krylov vectors and the cell scratch tw live only between two momentum solves, when the momentum
  //    vectors r^, p, v, t are dead: they are views of those (all cell-located: their padding stays zero);
  //  * the work vector of the restart residual is the last basis vector, which is free at a cycle start (below).
  s.kr = s.alloc_v3(), s.krh = s.alloc_v3(), s.kp = s.alloc_v3(), s.kv = s.alloc_v3(), s.kt = s.alloc_v3();
  s.ks = s.kr;
  {
    std::vector<double *> scratch;
    for (const V3 *q : {&s.krh, &s.kp, &s.kv, &s.kt})
      for (int c = 0; c < dim; ++c) scratch.push_back(q->c[c]);
    size_t at = 0;
    s.pr = scratch[at++], s.pp = scratch[at++], s.pq = scratch[at++];
    s.ps = scratch[at++], s.pt = scratch[at++], s.prh = scratch[at++]; // BiCGStab forms (pressure outlet, DIAG / ROWSUM Schur complement)
    s.tw = V3{{nullptr, nullptr, nullptr}};
    for (int c = 0; c < dim; ++c) s.tw.c[c] = scratch[at++];
  }
  if (opt.mode == 0) {
    s.basis_size = opt.outer_restart + 1;
    s.basis.resize(s.basis_size);
    for (auto &vec : s.basis) {
      vec.resize(7, nullptr);
      for (int f = 0; f < 7; ++f)
        if (dim == 3 || (f != 2 && f != 5)) vec[f] = s.alloc_field();
    }
    s.zbasis.resize(opt.outer_restart);
    for (auto &vec : s.zbasis) {
      vec.resize(7, nullptr);
      for (int f = 0; f < 7; ++f)
        if (dim == 3 || (f != 2 && f != 5)) vec[f] = s.alloc_field();
    }
    // r = b - M x of a restart is formed in the last basis vector (free until the cycle reaches it)
    for (int c = 0; c < 3; ++c) s.wv.c[c] = s.basis[opt.outer_restart][c], s.wU.c[c] = s.basis[opt.outer_restart][3 + c];
    s.wp = s.basis[opt.outer_restart][6];
    for (int c = 0; c < 3; ++c) s.zv.c[c] = s.zbasis[0][c], s.zU.c[c] = s.zbasis[0][3 + c];
    s.zp = s.zbasis[0][6];
  }
  mg_setup(s);
  s.ex.sync();
}

void solver_destroy(Solver &s)
{
  s.ex.sync();
  ibm_destroy(s);
#ifndef FLUCA_HOSTEMU
  if (s.view.stream) cudaStreamSynchronize(s.view.stream);
  if (s.view.ready) cudaEventDestroy(s.view.ready);
  if (s.view.done) cudaEventDestroy(s.view.done);
  if (s.view.stream) cudaStreamDestroy(s.view.stream);
  for (double *hp : {s.view.v, s.view.U[0], s.view.U[1], s.view.U[2], s.view.p, s.view.phalf})
    if (hp) cudaFreeHost(hp);
#else
  for (double *hp : {s.view.v, s.view.U[0], s.view.U[1], s.view.U[2], s.view.p, s.view.phalf}) free(hp);
#endif
  s.view = StateView();
#ifndef FLUCA_HOSTEMU
  for (double *p : s.pool) tensor_map_forget(p); // the cache is keyed by address: a later solver may get the same one
  for (void *p : s.mg_owned) tensor_map_forget((const double *)p);
#endif
  mg_destroy(s);
  for (double *p : s.pool) dev_free(p);
  s.pool.clear();
  for (int b = 0; b < 6; ++b)
    for (int k = 0; k < 2; ++k)
      for (int sl = 0; sl < 2; ++sl) dev_free(s.bc_store[b][k][sl]), s.bc_store[b][k][sl] = nullptr;
  s.comm.reset();
  s.ex.destroy();
}

// ------------------------------------------------------------------ dispatch helpers
#define DIM_DISPATCH(s, CALL2, CALL3) \
  do { \
    if ((s).dim == 2) { CALL2; } else { CALL3; } \
  } while (0)

static Box cell_box(const Solver &s)
{
  Box b = {s.gh.g.nx, s.gh.g.ny, s.gh.g.nzl};
  return b;
}

// Resident CTAs per SM the streaming box kernels of this file are compiled for (FLUCA_B200_BOX_MINB = 2, 3 or 4: 128 / 80 / 64
// registers per thread).  These kernels keep one plane of loads in flight per thread, so their bandwidth follows occupancy.
static int box_minb()
{
  static const int v = getenv("FLUCA_B200_BOX_MINB") ? atoi(getenv("FLUCA_B200_BOX_MINB")) : 3;
  return v;
}
#define FOR_BOX_MINB(U, box, f) \
  do { \
    switch (box_minb()) { \
    case 2: for_box<U, 2>(s.ex, box, f); break; \
    case 4: for_box<U, 4>(s.ex, box, f); break; \
    default: for_box<U, 3>(s.ex, box, f); break; \
    } \
  } while (0)
#define FOR_BOX_REDUCE_MINB(NR, box, f) \
  do { \
    switch (box_minb()) { \
    case 2: for_box_reduce<NR, 2>(s.ex, box, f); break; \
    case 4: for_box_reduce<NR, 4>(s.ex, box, f); break; \
    default: for_box_reduce<NR, 3>(s.ex, box, f); break; \
    } \
  } while (0)

// f(i, j, kl) over the cells adjacent to boundary b held by this rank
template <class F>
struct BoundaryAdapter {
  F   f;
  int d, fixed;
  FL_HD void operator()(int a, int b, int c) const
  {
    if (d == 0) f(fixed, b, c);
    else if (d == 1) f(a, fixed, c);
    else f(a, b, fixed);
  }
};

template <class F>
static void for_boundary(Solver &s, int b, F f)
{
  const Geom &g = s.gh.g;
  const int   d = b / 2, side = b % 2;
  const Tab  &T = g.t[d];
  if (T.per) return;
  if (d == 2 && !(side ? T.wall_hi : T.wall_lo)) return;
  const int          ext[3] = {g.nx, g.ny, g.nzl};
  Box                box    = {d == 0 ? 1 : g.nx, d == 1 ? 1 : g.ny, d == 2 ? 1 : g.nzl};
  BoundaryAdapter<F> ad     = {f, d, side ? ext[d] - 1 : 0};
  for_box(s.ex, box, ad);
}

// ------------------------------------------------------------------ right-hand side (NSFormFunction)
template <int DIM>
static void build_rhs(Solver &s)
{
  KScope ks(s.ex, KT_RHS_PROJECT);
  const Geom      &g  = s.gh.g;
  const bool       first = (s.step_index == 0);
  double          *q  = first ? s.p : s.phalf; // cnlinearcart3d.c:2972-2982
  halo_cells(s, s.v0);
  halo_faces(s, s.U0);
  halo_scalar(s, q);
  MomentumRhs<DIM> mr;
  mr.g = g, mr.sp = s.sp, mr.bc = s.bc, mr.v0 = CV3(s.v0), mr.q = q, mr.r = s.rm;
  FOR_BOX_MINB(1, cell_box(s), mr);

  // r_int = bcT(t+dt) + (-T) dt/rho (bcG(tq) - bcG(th)) + dt/rho (bcGst(tq) - bcGst(th)),  :2998-3033
  for (int d = 0; d < DIM; ++d) dev_zero(s.ex, s.ri.c[d], sizeof(double) * (size_t)g.nalloc);
  dev_zero(s.ex, s.rc, sizeof(double) * (size_t)g.nalloc); // :3035
  for (int b = 0; b < 2 * DIM; ++b) {
    const int  d = b / 2, side = b % 2;
    const Tab &T  = g.t[d];
    const int  ty = side ? T.bc_hi : T.bc_lo;
    double    *rid = s.ri.c[d];
    const BcDev bc = s.bc;
    const Geom  gg = g;
    if (ty == BC_VELOCITY) {
      for_boundary(s, b, FL_LAMBDA(int i, int j, int kl) {
        Nbr<DIM> nb;
        nbr<DIM>(gg, i, j, kl, nb);
        const long pt = bc_pt(gg, b, i, j, kl);
        const long f  = side ? nb.fu[d] : nb.c;
        rid[f]        = bc.vel[b][1][d * bc.npts[b] + pt];
      });
    } else if (ty == BC_PRESSURE_OUTLET) {
      const double dtrho = s.sp.dtrho;
      for_boundary(s, b, FL_LAMBDA(int i, int j, int kl) {
        Nbr<DIM> nb;
        nbr<DIM>(gg, i, j, kl, nb);
        const Tab   &TT = gg.t[d];
        const long   pt = bc_pt(gg, b, i, j, kl);
        const double dl = bc.prs[b][0][pt] - bc.prs[b][1][pt]; // p_b(t_q) - p_b(t+dt/2)
        const int    n  = TT.n;
        if (!side) {
          const double g0 = dtrho * TT.gr_bc_lo * dl; // the only non-zero entry of dt/rho (bcG_q - bcG_h), at cell 0
          rid[nb.c] += -TT.tn_lo[0] * g0 + dtrho * TT.gst_bc_lo * dl;
          rid[nb.p[d]] += -TT.itw[2 * 1 + 0] * g0; // face 1 interpolates cells (0, 1)
        } else {
          const double g0 = dtrho * TT.gr_bc_hi * dl;
          rid[nb.fu[d]] += -TT.tn_hi[1] * g0 + dtrho * TT.gst_bc_hi * dl;
          rid[nb.c] += -TT.itw[2 * (n - 1) + 1] * g0; // face n-1 interpolates cells (n-2, n-1)
        }
      });
    }
  }
}

// PCABFSetSchurComplementAinvType / PCABFSetUpperTriangularAinvType (abfpc.c:300-318; -ns_pc_abf_{schur,upper}_ainv_type)
void set_ainv_types(Solver &s, int schur_type, int upper_type)
{
  if (schur_type < 0 || schur_type > 2 || upper_type < 0 || upper_type > 2) throw Error(FL_ERR_ARG, "unknown A-inverse type (0 ID, 1 DIAG, 2 ROWSUM)");
  s.opt.schur_ainv = schur_type, s.opt.upper_ainv = upper_type;
  if (schur_type != 0) {
    if (!s.ainv_store[0].c[0]) s.ainv_store[0] = s.alloc_v3();
  }
  if (upper_type != 0 && upper_type != schur_type && !s.ainv_store[1].c[0]) s.ainv_store[1] = s.alloc_v3();
  s.ainv_s = schur_type != 0 ? s.ainv_store[0] : V3{};
  s.ainv_u = upper_type == 0 ? V3{} : (upper_type == schur_type ? s.ainv_store[0] : s.ainv_store[1]);
  s.prepared = false; // the vectors belong to the A of a step: rebuilt by the next prepare_step
}

// 1 / diag(A) or 1 / rowsum(A) of the step's momentum operator (PCSetUp_ABF runs once per step, abfpc.c:155-160)
static void build_ainv(Solver &s)
{
  for (int which = 0; which < 2; ++which) {
    const int type = which == 0 ? s.opt.schur_ainv : s.opt.upper_ainv;
    const V3 &dst  = which == 0 ? s.ainv_s : s.ainv_u;
    if (type == 0 || (which == 1 && type == s.opt.schur_ainv)) continue;
    if (s.dim == 2) {
      AinvCells<2> f;
      f.g = s.gh.g, f.sp = s.sp, f.bc = s.bc, f.v0 = CV3(s.v0), f.U0 = CV3(s.U0), f.type = type, f.ainv = dst;
      for_box(s.ex, cell_box(s), f);
    } else {
      AinvCells<3> f;
      f.g = s.gh.g, f.sp = s.sp, f.bc = s.bc, f.v0 = CV3(s.v0), f.U0 = CV3(s.U0), f.type = type, f.ainv = dst;
      for_box(s.ex, cell_box(s), f);
    }
  }
}

void prepare_step(Solver &s, double t, int step_index)
{
  s.t = t, s.step_index = step_index;
  // sol0 = sol (nsbasic.c:281-282) without a copy: during a step the time-n fields ARE the live state (nothing writes
  // s.v / s.U before the rotation at the end of do_step), so forming the right-hand side has no side effect on the state
  // and may be repeated (NSFormFunction between two steps, as VecCopy(sol, sol0) may)
  s.v0 = s.v, s.U0 = s.U;
  DIM_DISPATCH(s, build_rhs<2>(s), build_rhs<3>(s));
  build_ainv(s); // after build_rhs: it exchanged the ghost planes of v0 and U0
  s.prepared = true, s.rhs_valid = true;
}

// ------------------------------------------------------------------ operators
template <int DIM>
struct AApplyPlain {
  Geom       g;
  StepParams sp;
  BcDev      bc;
  CV3        x, v0, U0;
  V3         y;
  FL_HD void operator()(int i, int j, int kl) const
  {
    double r[DIM];
    a_apply_cell<DIM>(g, sp, bc, x, v0, U0, i, j, kl, r);
    const long c = g.idx(i, j, kl);
#pragma unroll
    for (int q = 0; q < DIM; ++q) y.c[q][c] = r[q];
  }
};

void a_apply(Solver &s, const V3 &x, const V3 &y)
{
  halo_cells(s, x);
#ifndef FLUCA_HOSTEMU
  if (tma_usable(s)) {
    a_apply_dots_tma(s, x, y, x, false);
    return;
  }
#endif
  if (s.dim == 2) {
    AApplyPlain<2> f;
    f.g = s.gh.g, f.sp = s.sp, f.bc = s.bc, f.x = CV3(x), f.v0 = CV3(s.v0), f.U0 = CV3(s.U0), f.y = y;
    for_box(s.ex, cell_box(s), f);
  } else {
    AApplyPlain<3> f;
    f.g = s.gh.g, f.sp = s.sp, f.bc = s.bc, f.x = CV3(x), f.v0 = CV3(s.v0), f.U0 = CV3(s.U0), f.y = y;
    for_box(s.ex, cell_box(s), f);
  }
}

static long global_cells(const Solver &s)
{
  const Geom &g = s.gh.g;
  return (long)g.nx * g.ny * g.nzg;
}

static void remove_mean(Solver &s, double *f)
{
  const long off = interior_off(s), len = interior_len(s);
  double    *a   = f + off;
  for_range_reduce<1>(s.ex, len, FL_LAMBDA(long i, double acc[1]) { acc[0] += a[i]; });
  double sum;
  reduce_finish(s, 1, &sum);
  // padding entries must stay zero: subtract on valid cells only
  const double mean = sum / (double)global_cells(s);
  const Geom   g    = s.gh.g;
  for_box(s.ex, cell_box(s), FL_LAMBDA(int i, int j, int kl) { f[g.idx(i, j, kl)] -= mean; });
}

template <int DIM>
static void abf_apply_t(Solver &s, const V3 &bm, const V3 &bi, const double *bcn, const V3 &ov, const V3 &oU, double *op, bool guess, double in_scale)
{
  KScope ks(s.ex, KT_RHS_PROJECT);
  const Geom &g = s.gh.g;
  // stage 1 (abfpc.c:72-77): momentum solve, then U* = r_int + T v* and the Poisson right-hand side in ONE pass (FaceStarRhs)
  if (momentum_solve(s, bm, s.vstar, guess, in_scale)) ++s.stats.inner_unconverged;
  halo_cells(s, s.vstar);
  halo_faces(s, bi); // the upper z face of the slab's last plane is formed from r_int's ghost face plane and v*'s ghost plane
  FaceStarRhs<DIM> fs;
  fs.g = g, fs.a = in_scale, fs.scale = s.sp.rho / s.sp.dt, fs.rcscale = in_scale, fs.in = CV3(bi), fs.w = CV3(s.vstar), fs.rc = bcn, fs.Us = s.Ustar, fs.out = s.srhs;
  FOR_BOX_REDUCE_MINB(1, cell_box(s), fs);
  if (!s.has_outlet) {
    // constant null space (abfpc.c:173-177): make the right-hand side compatible
    double sum;
    reduce_finish(s, 1, &sum);
    const double mean = sum / (double)global_cells(s);
    double      *sr   = s.srhs;
    const Geom   gg   = g;
    for_box(s.ex, cell_box(s), FL_LAMBDA(int i, int j, int kl) { sr[gg.idx(i, j, kl)] -= mean; });
  }
  if (poisson_solve(s, s.srhs, op)) ++s.stats.inner_unconverged;
  if (!s.has_outlet) remove_mean(s, op);
  // stage 2 (abfpc.c:80-101); the T*G~p terms of V cancel: V = V* - G~st p
  halo_scalar(s, op);
  if (s.opt.upper_ainv == 0) {
    ProjectAll<DIM> pa; // v = v* - G~ p and U = U* - G~st p in one pass
    pa.g = g, pa.dtrho = s.sp.dtrho, pa.vs = CV3(s.vstar), pa.Us = CV3(s.Ustar), pa.p = op, pa.v = ov, pa.U = oU;
    FOR_BOX_MINB(2, cell_box(s), pa);
  } else {
    // DIAG / ROWSUM (abfpc.c:81-99): v = v* - a1 G~ p,  V = V* - T (a1 G~ p) - (-R) p = V* - G~st p + T w,  w = (1 - a1) G~ p
    GradScaleCells<DIM> gs;
    gs.g = g, gs.scale = s.sp.dtrho, gs.ainv = CV3(s.ainv_u), gs.p = op, gs.vs = CV3(s.vstar), gs.v = ov, gs.w = s.tw;
    for_box(s.ex, cell_box(s), gs);
    halo_cells(s, s.tw);
    FaceCombine<DIM> fc;
    fc.g = g, fc.a = 1., fc.b = 1., fc.c = -s.sp.dtrho, fc.in = CV3(s.Ustar), fc.w = CV3(s.tw), fc.p = op, fc.out = oU;
    for_box<2>(s.ex, cell_box(s), fc);
  }
  s.stats.abf_applies++;
}

void abf_apply(Solver &s, const V3 &bm, const V3 &bi, const double *bcn, const V3 &ov, const V3 &oU, double *op, bool guess, double in_scale)
{
  DIM_DISPATCH(s, abf_apply_t<2>(s, bm, bi, bcn, ov, oU, op, guess, in_scale), abf_apply_t<3>(s, bm, bi, bcn, ov, oU, op, guess, in_scale));
}

template <int DIM>
static void coupled_apply_t(Solver &s, const V3 &xv, const V3 &xU, double *xp, const V3 &yv, const V3 &yU, double *yp)
{
  const Geom &g = s.gh.g;
  halo_cells(s, xv);
  halo_scalar(s, xp);
  bool tiled = false;
#ifndef FLUCA_HOSTEMU
  if (DIM == 3 && tma_usable(s)) {
    coupled_cells_tma(s, xv, xp, yv, s.tw, true);
    tiled = true;
  }
#endif
  if (!tiled) {
    CoupledCells<DIM> cc;
    cc.g = g, cc.sp = s.sp, cc.bc = s.bc, cc.x = CV3(xv), cc.v0 = CV3(s.v0), cc.U0 = CV3(s.U0), cc.p = xp, cc.y = yv, cc.w = s.tw;
    for_box(s.ex, cell_box(s), cc);
  }
  halo_cells(s, s.tw);
  FaceCombine<DIM> fc;
  fc.g = g, fc.a = 1., fc.b = -1., fc.c = s.sp.dtrho, fc.in = CV3(xU), fc.w = CV3(s.tw), fc.p = xp, fc.out = yU;
  for_box<2>(s.ex, cell_box(s), fc);
  halo_faces(s, xU);
  DivCell<DIM> dc;
  dc.g = g, dc.U = CV3(xU), dc.out = yp;
  FOR_BOX_MINB(2, cell_box(s), dc);
}

// M z for z = ABF(b): only the momentum and continuity blocks need arithmetic -- y_v = A z_v + (dt/rho) G z_p, y_p = D z_U; the
// interpolation block equals b's (see outer_gmres) and is not formed here.
template <int DIM>
static void coupled_apply_abf_output_t(Solver &s, const V3 &zv, const V3 &zU, double *zp, const V3 &yv, double *yp)
{
  const Geom &g = s.gh.g;
  halo_cells(s, zv);
  halo_scalar(s, zp);
  bool tiled = false;
#ifndef FLUCA_HOSTEMU
  if (DIM == 3 && tma_usable(s)) {
    coupled_cells_tma(s, zv, zp, yv, s.tw, false);
    tiled = true;
  }
#endif
  if (!tiled) {
    CoupledCells<DIM> cc;
    cc.g = g, cc.sp = s.sp, cc.bc = s.bc, cc.x = CV3(zv), cc.v0 = CV3(s.v0), cc.U0 = CV3(s.U0), cc.p = zp, cc.y = yv, cc.w = V3{{nullptr, nullptr, nullptr}};
    for_box(s.ex, cell_box(s), cc);
  }
  halo_faces(s, zU);
  DivCell<DIM> dc;
  dc.g = g, dc.U = CV3(zU), dc.out = yp;
  FOR_BOX_MINB(2, cell_box(s), dc);
}

static void coupled_apply_abf_output(Solver &s, const V3 &zv, const V3 &zU, double *zp, const V3 &yv, double *yp)
{
  DIM_DISPATCH(s, coupled_apply_abf_output_t<2>(s, zv, zU, zp, yv, yp), coupled_apply_abf_output_t<3>(s, zv, zU, zp, yv, yp));
}

void coupled_apply(Solver &s, const V3 &xv, const V3 &xU, double *xp, const V3 &yv, const V3 &yU, double *yp)
{
  DIM_DISPATCH(s, coupled_apply_t<2>(s, xv, xU, xp, yv, yU, yp), coupled_apply_t<3>(s, xv, xU, xp, yv, yU, yp));
}

// S p in the reference's scaling: S = -(dt/rho) D Gst0, or D((-T) a1 G~ - (-R)) for the DIAG / ROWSUM variants (abfpc.c:151-170)
void schur_apply_reference_scaling(Solver &s, double *pin, double *out)
{
  poisson_apply(s, pin, out);
  const Geom   g     = s.gh.g;
  const double dtrho = s.sp.dtrho;
  const int    dim   = s.dim;
  for_box(s.ex, cell_box(s), FL_LAMBDA(int i, int j, int kl) {
    double vol = g.t[0].h[i] * g.t[1].h[j];
    if (dim == 3) vol *= g.t[2].h[g.k0 + kl];
    out[g.idx(i, j, kl)] *= dtrho / vol;
  });
}

// ------------------------------------------------------------------ 7-field vectors of the outer solve
struct MV {
  double *f[7];
  long    len[7];
  long    n; // max len
};

static MV make_mv(const Solver &s, const V3 &v, const V3 &U, double *p)
{
  MV         m;
  const long off = interior_off(s), len = interior_len(s);
  for (int c = 0; c < 3; ++c) {
    m.f[c]       = v.c[c] ? v.c[c] + off : nullptr;
    m.len[c]     = v.c[c] ? len : 0;
    m.f[3 + c]   = U.c[c] ? U.c[c] + off : nullptr;
    m.len[3 + c] = U.c[c] ? face_len(s, c) : 0;
  }
  m.f[6]   = p + off;
  m.len[6] = len;
  m.n      = 0;
  for (int f = 0; f < 7; ++f)
    if (m.len[f] > m.n) m.n = m.len[f];
  return m;
}
static MV make_mv(const Solver &s, const std::vector<double *> &b)
{
  V3 v, U;
  for (int c = 0; c < 3; ++c) v.c[c] = b[c], U.c[c] = b[3 + c];
  return make_mv(s, v, U, b[6]);
}

static double mv_dot(Solver &s, const MV &a, const MV &b)
{
  for_range_reduce<1>(s.ex, a.n, FL_LAMBDA(long i, double acc[1]) {
    double t = 0.;
#pragma unroll
    for (int f = 0; f < 7; ++f)
      if (i < a.len[f]) t += a.f[f][i] * b.f[f][i];
    acc[0] += t;
  });
  double r;
  reduce_finish(s, 1, &r);
  return r;
}
// y = alpha * x + beta * y   (beta == 0 overwrites)
static void mv_axpby(Solver &s, double alpha, const MV &x, double beta, const MV &y)
{
  for_range(s.ex, x.n, FL_LAMBDA(long i) {
#pragma unroll
    for (int f = 0; f < 7; ++f)
      if (i < x.len[f]) y.f[f][i] = alpha * x.f[f][i] + (beta == 0. ? 0. : beta * y.f[f][i]);
  });
}
static void mv_zero(Solver &s, const MV &y)
{
  for_range(s.ex, y.n, FL_LAMBDA(long i) {
#pragma unroll
    for (int f = 0; f < 7; ++f)
      if (i < y.len[f]) y.f[f][i] = 0.;
  });
}


// ---- multi-vector kernels of the outer GMRES.  A 7-field vector is streamed as NA active fields of one common length
// plus the extra FRONT-face plane of the z-face field on the last wall rank (the only field that can be longer); the
// number of basis vectors per launch is a template parameter, so the loops carry no run-time predicate.
struct MVC { // compacted view
  double *f[7];
  int     na;   // active fields (7 in 3-D, 5 in 2-D)
  long    len;  // common length
  double *tail; // field with extra entries, positioned at its entry `len`; nullptr if none
  long    ntail;
};
static MVC compact(const MV &m)
{
  MVC c;
  c.na = 0, c.len = 0, c.tail = nullptr, c.ntail = 0;
  for (int f = 0; f < 7; ++f) c.f[f] = nullptr;
  for (int f = 0; f < 7; ++f)
    if (m.len[f] > 0 && (c.len == 0 || m.len[f] < c.len)) c.len = m.len[f];
  for (int f = 0; f < 7; ++f)
    if (m.len[f] > 0) {
      if (m.len[f] > c.len) c.tail = m.f[f] + c.len, c.ntail = m.len[f] - c.len;
      c.f[c.na++] = m.f[f];
    }
  return c;
}
static const int MVB = 8; // basis vectors per launch
template <int NC>
struct MVCSet {
  MVC v[NC];
};

// red[j] = <V_j, w> (j < NC), red[NC] = <w, w>
template <int NC>
static void mv_dots_t(Solver &s, const MV *V, const MV &w, double *red)
{
  MVCSet<NC> S;
  for (int j = 0; j < NC; ++j) S.v[j] = compact(V[j]);
  const MVC W = compact(w);
  for_range_reduce<NC + 1>(s.ex, W.len, FL_LAMBDA(long i, double acc[NC + 1]) {
    double t = 0., d[NC];
#pragma unroll
    for (int j = 0; j < NC; ++j) d[j] = 0.;
#pragma unroll
    for (int f = 0; f < 7; ++f)
      if (f < W.na) {
        const double wv = W.f[f][i];
        t += wv * wv;
#pragma unroll
        for (int j = 0; j < NC; ++j) d[j] += S.v[j].f[f][i] * wv;
      }
    if (i < W.ntail) {
      const double wv = W.tail[i];
      t += wv * wv;
#pragma unroll
      for (int j = 0; j < NC; ++j) d[j] += S.v[j].tail[i] * wv;
    }
#pragma unroll
    for (int j = 0; j < NC; ++j) acc[j] += d[j];
    acc[NC] += t;
  });
  reduce_finish(s, NC + 1, red);
}

struct MVCoef {
  double c[MVB];
};

// out = beta * w + sum_j c_j V_j ; returns |out|^2.  out may be w itself.
template <int NC>
static double mv_comb_t(Solver &s, const MV *V, const double *c, double beta, const MV &w, const MV &out)
{
  MVCSet<NC> S;
  for (int j = 0; j < NC; ++j) S.v[j] = compact(V[j]);
  const MVC W = compact(w), O = compact(out);
  MVCoef    C;
  for (int j = 0; j < MVB; ++j) C.c[j] = j < NC ? c[j] : 0.;
  for_range_reduce<1>(s.ex, W.len, FL_LAMBDA(long i, double acc[1]) {
    double t = 0.;
#pragma unroll
    for (int f = 0; f < 7; ++f)
      if (f < W.na) {
        double wv = beta == 0. ? 0. : beta * W.f[f][i];
#pragma unroll
        for (int j = 0; j < NC; ++j) wv += C.c[j] * S.v[j].f[f][i];
        O.f[f][i] = wv;
        t += wv * wv;
      }
    if (i < W.ntail) {
      double wv = beta == 0. ? 0. : beta * W.tail[i];
#pragma unroll
      for (int j = 0; j < NC; ++j) wv += C.c[j] * S.v[j].tail[i];
      O.tail[i] = wv;
      t += wv * wv;
    }
    acc[0] += t;
  });
  double r;
  reduce_finish(s, 1, &r);
  return r;
}

#define MV_DISPATCH(nc, CALL) \
  switch (nc) { \
  case 1: CALL(1); break; \
  case 2: CALL(2); break; \
  case 3: CALL(3); break; \
  case 4: CALL(4); break; \
  case 5: CALL(5); break; \
  case 6: CALL(6); break; \
  case 7: CALL(7); break; \
  default: CALL(8); break; \
  }

// out = beta * w + sum_{j < nv} c_j V_j in blocks of MVB vectors; returns |out|^2
static double mv_lincomb(Solver &s, const std::vector<MV> &V, int nv, const double *c, double beta, const MV &w, const MV &out)
{
  double nrm2 = 0.;
  if (nv == 0) {
    mv_axpby(s, beta, w, 0., out);
    return mv_dot(s, out, out);
  }
  for (int j0 = 0; j0 < nv; j0 += MVB) {
    const int nc = nv - j0 < MVB ? nv - j0 : MVB;
#define CALL(N) nrm2 = mv_comb_t<N>(s, &V[j0], c + j0, j0 == 0 ? beta : 1., j0 == 0 ? w : out, out)
    MV_DISPATCH(nc, CALL)
#undef CALL
  }
  return nrm2;
}

// fused classical Gram-Schmidt step against the UNNORMALISED basis V_j (norms vn_j): h_j = <V_j, w> / vn_j (j < nv),
// ww = <w, w>; out = w - sum_j (h_j / vn_j) V_j; nrm2 = |out|^2.  `w` is a read-only VIEW (its face block may alias the face
// block of a basis vector, see outer_gmres); `out` is where the new basis vector is stored (out may be w itself).
// dots_only: stop after the projections of the first block (nv <= MVB) -- the caller decides from h and ww whether the
// vector is needed at all, and calls again with h_known to form it.
static void mv_project(Solver &s, const std::vector<MV> &V, const double *vn, int nv, const MV &w, const MV &out, double *h, double &ww, double &nrm2, bool dots_only = false, bool h_known = false)
{
  if (!h_known) ww = 0.;
  for (int j0 = 0; j0 < nv; j0 += MVB) {
    const int nc = nv - j0 < MVB ? nv - j0 : MVB;
    double    red[MVB + 1], neg[MVB];
    const MV &win = j0 == 0 ? w : out;
    if (!(h_known && j0 == 0)) {
#define CALL(N) mv_dots_t<N>(s, &V[j0], win, red)
      MV_DISPATCH(nc, CALL)
#undef CALL
      if (j0 == 0) ww = red[nc];
      for (int j = 0; j < nc; ++j) h[j0 + j] = red[j] / vn[j0 + j];
    }
    if (dots_only) return;
    for (int j = 0; j < nc; ++j) neg[j] = -h[j0 + j] / vn[j0 + j];
#define CALL(N) nrm2 = mv_comb_t<N>(s, &V[j0], neg, 1., win, out)
    MV_DISPATCH(nc, CALL)
#undef CALL
  }
}

// Right-preconditioned restarted GMRES on M x = b, PC = ABF, zero initial guess, true-residual norm (the outer KSP of
// nssol.c:13-30); x is built in (s.xv, s.xU, s.xp).  Two things differ from a textbook loop, neither changes the iterates:
//  * the preconditioned vectors z_k = ABF(v_k) are kept (flexible form), so the solution update x += sum_k y_k z_k needs no
//    further ABF application -- PETSc's KSPGMRES applies the preconditioner once more per cycle to build the solution;
//  * basis vectors are stored unnormalised with their norms on the host (the first one is b itself, never copied): the
//    scale goes into the ABF application and the Gram-Schmidt coefficients instead of an extra pass over 7 fields.
static int outer_gmres(Solver &s)
{
  KScope ks(s.ex, KT_OUTER);
  const int m = s.opt.outer_restart;
  MV        X = make_mv(s, s.xv, s.xU, s.xp), Bv = make_mv(s, s.rm, s.ri, s.rc), W = make_mv(s, s.wv, s.wU, s.wp);
  std::vector<MV> V, Zb;
  for (auto &b : s.basis) V.push_back(make_mv(s, b));
  for (auto &b : s.zbasis) Zb.push_back(make_mv(s, b));
  std::vector<double> H((size_t)(m + 1) * m, 0.), cs(m), sn(m), gvec(m + 1), y(m), vn(m + 1, 1.);
  const bool   no_lazy = getenv("FLUCA_B200_NO_LAZY_BASIS") != nullptr; // always form the last basis vector (A/B timing, tests)
  const double relax_c = getenv("FLUCA_B200_RELAX") ? atof(getenv("FLUCA_B200_RELAX")) : 0.05; // 0 switches the relaxation off; measured: 0.05 keeps the outer counts, 0.2 adds outer iterations
  int    its = 0;
  double rnorm0 = -1., rnorm = 0.;
  bool   done = false, first_cycle = true, x_zero = true;
  s.stats.nhist = 0;
  while (!done) {
    // r = b - M x; in the first cycle x = 0 and the first basis vector is b in place
    const double one = 1.;
    const bool   alias_b = first_cycle;
    if (first_cycle) rnorm = std::sqrt(mv_dot(s, Bv, Bv));
    else {
      coupled_apply(s, s.xv, s.xU, s.xp, s.wv, s.wU, s.wp);
      rnorm = std::sqrt(mv_comb_t<1>(s, &Bv, &one, -1., W, V[0])); // V0 = b - M x
    }
    first_cycle = false;
    if (rnorm0 < 0.) {
      rnorm0 = rnorm;
      if (s.stats.nhist < 128) s.stats.hist[s.stats.nhist++] = rnorm;
    }
    if (!(rnorm == rnorm)) throw Error(FL_ERR_DIVERGED, "outer residual is NaN");
    if (rnorm <= s.opt.outer_rtol * rnorm0 || rnorm == 0.) {
      s.stats.converged = 1;
      break;
    }
    if (its >= s.opt.outer_maxit) break;
    std::vector<MV> Vl(V);
    if (alias_b) Vl[0] = Bv;
    vn[0] = rnorm;
    std::fill(gvec.begin(), gvec.end(), 0.);
    gvec[0] = rnorm;
    int k = 0;
    for (; k < m && its < s.opt.outer_maxit; ++k) {
      // Z_k = ABF(V_k) with V_k UNNORMALISED (norm vn_k): Z_k = vn_k z_k, W = M Z_k = vn_k w.  Only host scalars carry the
      // factor: no pass over the fields scales anything.
      V3      kv, kU, zv, zU;
      double *kp = (k == 0 && alias_b) ? s.rc : s.basis[k][6], *zp = s.zbasis[k][6];
      for (int c = 0; c < 3; ++c) {
        kv.c[c] = (k == 0 && alias_b) ? s.rm.c[c] : s.basis[k][c];
        kU.c[c] = (k == 0 && alias_b) ? s.ri.c[c] : s.basis[k][3 + c];
        zv.c[c] = s.zbasis[k][c], zU.c[c] = s.zbasis[k][3 + c];
      }
      // the first application of a step acts on b itself: the previous velocity (or the forced IBM predictor) is its guess
      const bool guess = s.have_guess && its == 0;
      // Inexact flexible GMRES: the Arnoldi relation M Z = V H holds whatever the accuracy of z_k = ABF(v_k), so the
      // residual estimate stays the true residual; an inner error of eta_k only needs eta_k |r_k| below the target
      // (Simoncini & Szyld 2003): inner rtol = max(user's, c * outer_rtol * |r_0| / |r_k|), at most 0.1.  The first
      // application (|r_0| = |b|) keeps the user's tolerances.
      s.tol_floor = (relax_c > 0. && rnorm > 0. && rnorm < rnorm0) ? relax_c * s.opt.outer_rtol * rnorm0 / rnorm : 0.;
      abf_apply(s, kv, kU, kp, zv, zU, zp, guess, 1.);
      s.tol_floor = 0.;
      // W = M Z_k.  Its interpolation block needs no arithmetic: the face update of the ABF application is the exact
      // algebraic inverse of that block row (abfpc.c:73-74 against :96-101: -T v + U + (-R) p = U* - T v* = the input), for
      // every A-inverse variant and whatever the accuracy of the two inner solves.  So W's face block IS V_k's face block:
      // the Gram-Schmidt kernels read it in place and no field is written for it.
      V3 nv;
      for (int c = 0; c < 3; ++c) nv.c[c] = s.basis[k + 1][c];
      coupled_apply_abf_output(s, zv, zU, zp, nv, s.basis[k + 1][6]);
      if (!s.has_outlet) remove_mean(s, s.basis[k + 1][6]); // null space of J (nsbasic.c:229-243)
      const MV Win = make_mv(s, nv, kU, s.basis[k + 1][6]);
      // classical Gram-Schmidt, all projections in one pass (PETSc's GMRES default,
      // KSPGMRESClassicalGramSchmidtOrthogonalization; its default refinement type is "never").  w = M z_k is nearly
      // parallel to v_k (ABF is a good preconditioner), so the remainder is small; one more pass removes the cancellation
      // error, which only matters when the requested tolerance comes near it: remainder / |w| against 1e-12 / outer_rtol
      std::vector<double> hcolv(m + 1, 0.), col(m + 1, 0.);
      double             *hcol = hcolv.data(), ww = 0., nrm2 = 0.;
      // column k of H and the residual it implies, for a given norm hn of the remainder (true scale: / vn_k)
      auto residual_with = [&](double hn, bool commit) {
        for (int jx = 0; jx <= k; ++jx) col[jx] = hcol[jx] / vn[k];
        col[k + 1] = hn;
        for (int jx = 0; jx < k; ++jx) {
          const double a = col[jx], b = col[jx + 1];
          col[jx]     = cs[jx] * a + sn[jx] * b;
          col[jx + 1] = -sn[jx] * a + cs[jx] * b;
        }
        const double a = col[k], b = col[k + 1], r = std::hypot(a, b);
        const double ck = r > 0. ? a / r : 1., sk = r > 0. ? b / r : 0., gk = gvec[k];
        if (commit) {
          for (int jx = 0; jx < k; ++jx) H[(size_t)jx * m + k] = col[jx];
          H[(size_t)k * m + k]       = r;
          H[(size_t)(k + 1) * m + k] = 0.;
          cs[k] = ck, sn[k] = sk;
          gvec[k + 1] = -sk * gvec[k];
          gvec[k]     = ck * gvec[k];
        }
        return std::fabs(sk * gk); // the residual norm after this column
      };
      bool lazy_done = false;
      if (k + 1 <= MVB) {
        // The last iteration of a solve never needs its new basis vector.  One pass gives every projection and <w, w>; the
        // remainder norm follows from Pythagoras (the basis is orthonormal to round-off).  If that estimate is safely clear of
        // cancellation and already meets the tolerance with a margin, the vector is not formed (7 fields read + 7 written).
        mv_project(s, Vl, vn.data(), k + 1, Win, V[k + 1], hcol, ww, nrm2, true);
        double sum2 = 0.;
        for (int jx = 0; jx <= k; ++jx) sum2 += hcol[jx] * hcol[jx];
        const double est2 = ww - sum2;
        if (no_lazy == false && est2 > 1e-8 * ww) {
          const double rn_est = residual_with(std::sqrt(est2) / vn[k], false);
          if (rn_est <= 0.9 * s.opt.outer_rtol * rnorm0) {
            rnorm     = residual_with(std::sqrt(est2) / vn[k], true);
            lazy_done = true;
          }
        }
        if (!lazy_done) mv_project(s, Vl, vn.data(), k + 1, Win, V[k + 1], hcol, ww, nrm2, false, true);
      } else {
        mv_project(s, Vl, vn.data(), k + 1, Win, V[k + 1], hcol, ww, nrm2);
      }
      if (lazy_done) {
        its++;
        if (s.stats.nhist < 128) s.stats.hist[s.stats.nhist++] = rnorm;
        ++k;
        done              = true;
        s.stats.converged = 1;
        break;
      }
      const double cancel = 1e-12 / s.opt.outer_rtol; // round-off left in the remainder, relative to the target
      if (nrm2 < 0.5 * ww && nrm2 < cancel * cancel * 1e4 * ww) {
        std::vector<double> h2(m + 1, 0.);
        double              ww2;
        mv_project(s, Vl, vn.data(), k + 1, V[k + 1], V[k + 1], h2.data(), ww2, nrm2);
        for (int jx = 0; jx <= k; ++jx) hcol[jx] += h2[jx];
      }
      const double nrm = std::sqrt(nrm2), hn = nrm / vn[k];
      vn[k + 1]        = nrm > 0. ? nrm : 1.;
      (void)residual_with(hn, true);
      its++;
      rnorm = std::fabs(gvec[k + 1]);
      if (s.stats.nhist < 128) s.stats.hist[s.stats.nhist++] = rnorm;
      if (rnorm <= s.opt.outer_rtol * rnorm0 || hn == 0.) {
        ++k;
        done              = true;
        s.stats.converged = 1;
        break;
      }
    }
    // x += sum_j y_j z_j
    for (int jx = k - 1; jx >= 0; --jx) {
      double sum = gvec[jx];
      for (int l = jx + 1; l < k; ++l) sum -= H[(size_t)jx * m + l] * y[l];
      y[jx] = sum / H[(size_t)jx * m + jx];
    }
    for (int jx = 0; jx < k; ++jx) y[jx] /= vn[jx]; // Z_j = vn_j z_j
    (void)mv_lincomb(s, Zb, k, y.data(), x_zero ? 0. : 1., X, X);
    x_zero = false;
    if (its >= s.opt.outer_maxit) done = true;
  }
  if (x_zero) mv_zero(s, X);
  s.stats.outer_its    = its;
  s.stats.outer_rnorm0 = rnorm0;
  s.stats.outer_rnorm  = rnorm;
  return s.stats.converged ? 0 : 1;
}

// ------------------------------------------------------------------ single-kernel launches for tools/kernel_bench.py
// (fluca_b200_time_kernel): one launch of the named kernel group on scratch fields of a prepared step; returns its algorithmic
// bytes per cell (3-D)
double bench_kernel(Solver &s, const std::string &n)
{
  const Geom &g = s.gh.g;
  if (n == "face_star_rhs") {
    if (s.dim == 3) {
      FaceStarRhs<3> fs;
      fs.g = g, fs.a = 1., fs.scale = s.sp.rho / s.sp.dt, fs.rcscale = 1., fs.in = CV3(s.ri), fs.w = CV3(s.vstar), fs.rc = s.rc, fs.Us = s.Ustar, fs.out = s.srhs;
      FOR_BOX_REDUCE_MINB(1, cell_box(s), fs);
    }
    return 88.;
  }
  if (n == "project_all") {
    if (s.dim == 3) {
      ProjectAll<3> pa;
      pa.g = g, pa.dtrho = s.sp.dtrho, pa.vs = CV3(s.vstar), pa.Us = CV3(s.Ustar), pa.p = s.xp, pa.v = s.xv, pa.U = s.xU;
      FOR_BOX_MINB(2, cell_box(s), pa);
    }
    return 104.;
  }
  if (n == "div_cell") {
    if (s.dim == 3) {
      DivCell<3> dc;
      dc.g = g, dc.U = CV3(s.xU), dc.out = s.srhs;
      FOR_BOX_MINB(2, cell_box(s), dc);
    }
    return 32.;
  }
  if (n == "coupled_abf_output") {
    coupled_apply_abf_output(s, s.xv, s.xU, s.xp, s.vstar, s.srhs);
    return 96. + 8. + 32.; // y_v = A z_v + G~ z_p (x, v0, U0, p -> y) and y_p = D z_U, as SURVEY 8(d) would count them
  }
  if (n == "momentum_rhs") {
    if (s.dim == 3) {
      MomentumRhs<3> mr;
      mr.g = g, mr.sp = s.sp, mr.bc = s.bc, mr.v0 = CV3(s.v0), mr.q = s.phalf, mr.r = s.rm;
      FOR_BOX_MINB(1, cell_box(s), mr);
    }
    return 56.;
  }
  throw Error(FL_ERR_ARG, "unknown kernel name");
}

// ------------------------------------------------------------------ the step
int do_step(Solver &s, double t, int step_index)
{
  const long l0 = s.ex.stats.launches;
  s.stats       = Stats();
  prepare_step(s, t, step_index);
  // initial guess of the first momentum solve: the velocity of the previous step (v* = v^n + O(dt))
  s.allow_guess = getenv("FLUCA_B200_NO_GUESS") == nullptr;
  s.have_guess  = s.allow_guess && step_index > 0;
  if (s.have_guess)
    for (int c = 0; c < s.dim; ++c) copy_d2d(s.ex, s.vstar.c[c], s.v0.c[c], sizeof(double) * (size_t)s.gh.g.nalloc);
  if (s.ibm.n > 0) ibm_force_rhs(s); // immersed-boundary forcing joins the momentum right-hand side (ibm.h)
  if (!s.has_outlet) remove_mean(s, s.rc); // F(0) = -b with the null space removed (nsbasic.c:133-144)
  int rc = 0;
  if (s.opt.mode == 1) {
    abf_apply(s, s.rm, s.ri, s.rc, s.xv, s.xU, s.xp, s.have_guess);
    s.stats.converged = 1;
  } else {
    rc = outer_gmres(s);
  }
  // sol <- x ; pressure extrapolation (cnlinearcart3d.c:2843-2854)
  view_fence(s); // a view of the previous state may still be reading p and p-half (its velocity buffers are not written by this step)
  {
    // three buffers rotate: the solution becomes the state, the old state is kept untouched for one more step (a pending
    // view copy may still read it; it is the solve vector of step n + 2), the spare becomes the next solve vector
    V3 tv = s.vprev, tU = s.Uprev;
    s.vprev = s.v, s.Uprev = s.U;
    s.v = s.xv, s.U = s.xU;
    s.xv = tv, s.xU = tU;
    const long   off = interior_off(s), len = interior_len(s);
    double      *p = s.p + off, *ph = s.phalf + off;
    const double *dp = s.xp + off;
    if (step_index == 0) {
      for_range(s.ex, len, FL_LAMBDA(long i) {
        const double p0 = p[i], d = dp[i];
        p[i]  = p0 + 2. * d;
        ph[i] = p0 + d;
      });
    } else {
      for_range(s.ex, len, FL_LAMBDA(long i) {
        const double h = ph[i], d = dp[i];
        p[i]  = h + 1.5 * d;
        ph[i] = h + d;
      });
    }
  }
  s.ex.sync();
  s.stats.launches = s.ex.stats.launches - l0;
  s.prepared       = false;
  return rc;
}

} // namespace fluca
