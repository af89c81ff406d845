/* oracle/ref_model/ref_driver.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Runs the REFERENCE's own Navier-Stokes sources (compiled from /root/reference by oracle/Makefile `ref`: cartdiscret.c, cnlinear.c,
 * cnlinearcart2d.c, cnlinearcart3d.c, abfpc.c, with the reference's headers) on top of the PETSc model of this directory, behind a
 * plain C interface that oracle/ref.py loads with ctypes.  What is the reference's: every stencil weight, every operator and
 * boundary-condition vector, the right-hand side, the Jacobian update, the ABF factors and their application, v0interp, the solution
 * update and the pressure extrapolation of the step.  What is this file's: the Mesh object (the reference's needs DMStagCreate), the
 * base-class plumbing the reference keeps in nsbasic.c / nssol.c (NSSetUp, NSStep, NSGetField: restated in the order of operations of
 * nsbasic.c:153-299), and SNESSolve, which the reference delegates to PETSc -- here one of
 *     0  the exact solution of J x = b (dense LU; the null space handled as KSP does: removed from b, zero-mean pressure),
 *        i.e. what the reference's SNES converges to with any convergent Krylov method and preconditioner;
 *     1  one application of the reference's PCABF with exact inner solves (-ns_ksp_type preonly), the classical fractional step;
 *     2  right-preconditioned GMRES(30) with the reference's PCABF and exact inner solves (the default KSP of nssol.c:13-30 with the
 *        inner solves taken to convergence), recording the residual history.
 * Fields cross the interface in the oracle's layout: v [component][k][j][i], U_d [k][j][i] with the extra face layer of a
 * non-periodic direction, p and p-half [k][j][i]; solution-sized vectors are v, U_x, U_y, (U_z), p back to back.
 */
#include <fluca/private/nslinearcnimpl.h>
#include <math.h>
#include <stdlib.h>
#include "petsc_model_impl.h"

PetscErrorCode PCCreate_ABF(PC); /* abfpc.c */
PetscErrorCode NSCreate_CNLinear(NS); /* cnlinear.c */

PetscClassId NS_CLASSID = 21, MESH_CLASSID = 22;

/* ------------------------------------------------------------------ Mesh: what flucamesh.h promises to the NS sources */
struct _p_Mesh {
  struct _p_PetscObject hdr;
  int                   dim, N[3], per[3];
  DM                    dm[4];
  struct model_coords  *coords;
};
PetscErrorCode MeshGetDimension(Mesh m, PetscInt *dim) { return *dim = m->dim, PETSC_SUCCESS; }
PetscErrorCode MeshGetDM(Mesh m, MeshDMType t, DM *dm) { return *dm = m->dm[t], PETSC_SUCCESS; }
PetscErrorCode MeshCreateGlobalVector(Mesh m, MeshDMType t, Vec *v) { return DMCreateGlobalVector(m->dm[t], v); }
PetscErrorCode MeshCreateMatrix(Mesh m, MeshDMType rt, MeshDMType ct, Mat *A)
{
  *A         = ModelMatCreateAIJ(m->dm[rt]->nglobal, m->dm[ct]->nglobal);
  (*A)->rl2g = &m->dm[rt]->l2g, (*A)->cl2g = &m->dm[ct]->l2g;
  return PETSC_SUCCESS;
}
PetscErrorCode MeshGetNumberBoundaries(Mesh m, PetscInt *nb) { return *nb = 2 * m->dim, PETSC_SUCCESS; }
PetscErrorCode MeshSetOutputSequenceNumber(Mesh m, PetscInt num, PetscReal val) { return (void)m, (void)num, (void)val, PETSC_SUCCESS; } /* no viewers here */
PetscErrorCode MeshGetOutputSequenceNumber(Mesh m, PetscInt *num, PetscReal *val) { return (void)m, *num = -1, *val = 0., PETSC_SUCCESS; }
static Mesh mesh_create(int dim, const int N[3], const int per[3], const double *const xf[3])
{
  Mesh m = (Mesh)calloc(1, sizeof(*m));
  int  d, li;
  ModelHeaderInit(m, MESH_CLASSID, "Mesh", MESHCART, NULL);
  m->dim = dim;
  for (d = 0; d < 3; ++d) m->N[d] = d < dim ? N[d] : 1, m->per[d] = d < dim ? per[d] : 0;
  m->coords = ModelCoordsCreate(dim, m->N, m->per);
  for (d = 0; d < dim; ++d) { /* 1-D product coordinates: [element][LEFT, ELEMENT], RIGHT = the next element's LEFT */
    const int    gs = m->coords->gs[d], gn = m->coords->gn[d];
    const double L = xf[d][m->N[d]] - xf[d][0];
    for (li = 0; li < gn; ++li) {
      const int g = li + gs;
      double    left, right = NAN;
      if (g < 0) left = xf[d][g + m->N[d]] - L, right = xf[d][g + m->N[d] + 1] - L;
      else if (g < m->N[d]) left = xf[d][g], right = xf[d][g + 1];
      else {
        left = xf[d][m->N[d]];
        if (m->per[d]) right = xf[d][1] + L;
      }
      m->coords->coord[d][2 * li] = left, m->coords->coord[d][2 * li + 1] = (left + right) / 2.;
      /* The partial element at the upper end of a non-periodic direction has no centre, but its slot exists and the reference READS
         it (cnlinearcart3d.c:1996,2055,2114: the face-normal interpolation at an upper pressure outlet passes arrc[N][ielem] as the
         wall coordinate).  PETSc's DMStagSetUniformCoordinatesProduct fills that slot like any other element, x_max + h/2 -- what a
         mesh built with MeshCartSetUniformCoordinates holds; on loaded non-uniform coordinates the reference leaves a stale value
         there (cart.c:137-143 sets centres for i < N only).  The model stores the mirror image of the last centre, which is
         x_max + h/2 on a uniform mesh. */
      if (g == m->N[d] && !m->per[d]) m->coords->coord[d][2 * li + 1] = 2. * xf[d][m->N[d]] - (xf[d][m->N[d] - 1] + xf[d][m->N[d]]) / 2.;
    }
  }
  /* cart.c:88-120: scalar, vector (dim dof per element), staggered scalar (1 dof per face), staggered vector (dim dof per face) */
  m->dm[MESH_DM_SCALAR]      = dim == 2 ? ModelDMStagCreate(2, m->N, m->per, 0, 0, 1, 0, m->coords) : ModelDMStagCreate(3, m->N, m->per, 0, 0, 0, 1, m->coords);
  m->dm[MESH_DM_VECTOR]      = dim == 2 ? ModelDMStagCreate(2, m->N, m->per, 0, 0, 2, 0, m->coords) : ModelDMStagCreate(3, m->N, m->per, 0, 0, 0, 3, m->coords);
  m->dm[MESH_DM_STAG_SCALAR] = dim == 2 ? ModelDMStagCreate(2, m->N, m->per, 0, 1, 0, 0, m->coords) : ModelDMStagCreate(3, m->N, m->per, 0, 0, 1, 0, m->coords);
  m->dm[MESH_DM_STAG_VECTOR] = dim == 2 ? ModelDMStagCreate(2, m->N, m->per, 0, 2, 0, 0, m->coords) : ModelDMStagCreate(3, m->N, m->per, 0, 0, 3, 0, m->coords);
  return m;
}
static void mesh_destroy(Mesh m)
{
  int d;
  for (d = 0; d < 4; ++d) ModelDMDestroy(m->dm[d]);
  ModelCoordsDestroy(m->coords);
  ModelHeaderFree(m);
  free(m);
}

/* ------------------------------------------------------------------ the handle */
typedef int (*ref_bc_cb)(int bnd, int kind, int dim, double t, const double *x, double *val);
typedef struct Ref Ref;
struct bc_ctx {
  Ref *h;
  int  bnd;
};
struct Ref {
  Mesh          mesh;
  NS            ns;
  PC            pc;
  ref_bc_cb     cb;
  struct bc_ctx bctx[6];
  struct _p_IS  is[3];
  int           dim, n[3], per[3];
  long          ncell, nface[3], nsol;
  int          *map_v, *map_U[3], *map_p; /* canonical position -> global entry of vdm / Sdm / sdm */
  /* SNESSolve */
  int    mode, maxit, nhist, its;
  double rtol, hist[256];
  double *last_b; /* right-hand side of the last solve, canonical */
  int    has_nullspace;
};
static PetscErrorCode bc_velocity(PetscInt dim, PetscReal t, const PetscReal x[], PetscScalar val[], void *ctx)
{
  struct bc_ctx *c = (struct bc_ctx *)ctx;
  PetscCheck(!c->h->cb(c->bnd, 0, (int)dim, t, x, val), 0, PETSC_ERR_LIB, "velocity callback of boundary %d failed", c->bnd);
  return PETSC_SUCCESS;
}
static PetscErrorCode bc_pressure(PetscInt dim, PetscReal t, const PetscReal x[], PetscScalar val[], void *ctx)
{
  struct bc_ctx *c = (struct bc_ctx *)ctx;
  PetscCheck(!c->h->cb(c->bnd, 1, (int)dim, t, x, val), 0, PETSC_ERR_LIB, "pressure callback of boundary %d failed", c->bnd);
  return PETSC_SUCCESS;
}

/* ------------------------------------------------------------------ what the NS sources call in nssol.c / nsbasic.c / flucaviewer */
static const char *const field_name[3] = {NS_FIELD_VELOCITY, NS_FIELD_FACE_NORMAL_VELOCITY, NS_FIELD_PRESSURE};
static const MeshDMType  field_dm[3]   = {MESH_DM_VECTOR, MESH_DM_STAG_SCALAR, MESH_DM_SCALAR};
static Ref              *ref_of_ns(NS ns) { return (Ref *)ns->mon_ctxs[0]; } /* the model keeps its handle in an unused monitor slot */
PetscErrorCode NSGetField(NS ns, const char name[], PetscInt *idx, MeshDMType *dmtype, IS *is)
{
  int f;
  for (f = 0; f < 3; ++f)
    if (!strcmp(field_name[f], name)) break;
  PetscCheck(f < 3, 0, PETSC_ERR_ARG_OUTOFRANGE, "Field \"%s\" not found", name);
  if (idx) *idx = f;
  if (dmtype) *dmtype = field_dm[f];
  if (is) *is = &ref_of_ns(ns)->is[f];
  return PETSC_SUCCESS;
}
PetscErrorCode NSCheckDiverged(NS ns) { return (void)ns, PETSC_SUCCESS; } /* nsbasic.c:425-436 reads the SNES reason: the exact solves do not diverge */

/* ------------------------------------------------------------------ canonical layout <-> DMStag global entries */
static PetscErrorCode build_map(Ref *h, DM dm, DMStagStencilLocation loc, int comp, int facedir, int *map, long *count)
{
  PetscInt x, y, z, m, n, p, ex, ey, ez, i, j, k;
  long     c = 0;
  PetscCall(DMStagGetCorners(dm, &x, &y, &z, &m, &n, &p, &ex, &ey, &ez));
  if (h->dim == 2) z = 0, p = 1, ez = 0;
  m += facedir == 0 ? ex : 0, n += facedir == 1 ? ey : 0, p += facedir == 2 ? ez : 0;
  for (k = z; k < z + p; ++k)
    for (j = y; j < y + n; ++j)
      for (i = x; i < x + m; ++i) {
        DMStagStencil s = {loc, i, j, k, comp};
        PetscInt      ix;
        PetscCall(DMStagStencilToIndexLocal(dm, h->dim, 1, &s, &ix));
        PetscCheck(dm->l2g.idx[ix] >= 0, 0, PETSC_ERR_PLIB, "canonical point (%d,%d,%d) has no DMStag entry", (int)i, (int)j, (int)k);
        if (map) map[c] = dm->l2g.idx[ix];
        ++c;
      }
  *count = c;
  return PETSC_SUCCESS;
}
static const DMStagStencilLocation face_loc[3] = {DMSTAG_LEFT, DMSTAG_DOWN, DMSTAG_BACK};
static PetscErrorCode build_maps(Ref *h)
{
  DM   sdm = h->mesh->dm[MESH_DM_SCALAR], vdm = h->mesh->dm[MESH_DM_VECTOR], Sdm = h->mesh->dm[MESH_DM_STAG_SCALAR];
  long cnt;
  int  d;
  PetscCall(build_map(h, sdm, DMSTAG_ELEMENT, 0, -1, NULL, &h->ncell));
  h->map_p = (int *)malloc(sizeof(int) * (size_t)h->ncell);
  h->map_v = (int *)malloc(sizeof(int) * (size_t)h->ncell * h->dim);
  PetscCall(build_map(h, sdm, DMSTAG_ELEMENT, 0, -1, h->map_p, &cnt));
  for (d = 0; d < h->dim; ++d) PetscCall(build_map(h, vdm, DMSTAG_ELEMENT, d, -1, h->map_v + h->ncell * d, &cnt));
  h->nsol = h->ncell * (h->dim + 1);
  for (d = 0; d < h->dim; ++d) {
    PetscCall(build_map(h, Sdm, face_loc[d], 0, d, NULL, &h->nface[d]));
    h->map_U[d] = (int *)malloc(sizeof(int) * (size_t)h->nface[d]);
    PetscCall(build_map(h, Sdm, face_loc[d], 0, d, h->map_U[d], &cnt));
    h->nsol += h->nface[d];
  }
  return PETSC_SUCCESS;
}
/* canonical solution-sized array <-> a nest (v, U, p) */
static void nest_put(Ref *h, Vec nest, const double *x)
{
  long i, o = 0;
  int  d;
  for (i = 0; i < h->ncell * h->dim; ++i) nest->sub[0]->a[h->map_v[i]] = x[o++];
  for (d = 0; d < h->dim; ++d)
    for (i = 0; i < h->nface[d]; ++i) nest->sub[1]->a[h->map_U[d][i]] = x[o++];
  for (i = 0; i < h->ncell; ++i) nest->sub[2]->a[h->map_p[i]] = x[o++];
}
static void nest_get(Ref *h, Vec nest, double *x)
{
  long i, o = 0;
  int  d;
  for (i = 0; i < h->ncell * h->dim; ++i) x[o++] = nest->sub[0]->a[h->map_v[i]];
  for (d = 0; d < h->dim; ++d)
    for (i = 0; i < h->nface[d]; ++i) x[o++] = nest->sub[1]->a[h->map_U[d][i]];
  for (i = 0; i < h->ncell; ++i) x[o++] = nest->sub[2]->a[h->map_p[i]];
}

/* ------------------------------------------------------------------ SNESSolve */
static PetscErrorCode solve_exact(Ref *h, NS ns, Vec b, Vec x)
{
  const int nb[3] = {(int)b->sub[0]->n, (int)b->sub[1]->n, (int)b->sub[2]->n}, off[3] = {0, nb[0], nb[0] + nb[1]}, n = nb[0] + nb[1] + nb[2], N = n + (h->has_nullspace ? 1 : 0);
  double   *a = (double *)calloc((size_t)N * N, sizeof(double)), *r = (double *)calloc((size_t)N, sizeof(double)), piv;
  int       bi, bj, i, q;
  for (bi = 0; bi < 3; ++bi)
    for (bj = 0; bj < 3; ++bj) {
      Mat B = ns->J->blk[bi][bj];
      if (!B) continue;
      for (i = 0; i < B->m; ++i)
        for (q = 0; q < B->rn[i]; ++q) a[(size_t)(off[bi] + i) * N + off[bj] + B->rc[i][q]] += B->rv[i][q];
    }
  ModelVecGather(b, r);
  if (h->has_nullspace) /* zero-mean pressure correction; the constant taken out of the continuity residual (KSP with a MatNullSpace) */
    for (i = 0; i < nb[2]; ++i) a[(size_t)(off[2] + i) * N + n] = 1., a[(size_t)n * N + off[2] + i] = 1.;
  piv = ModelDenseSolve(N, a, r);
  free(a);
  if (!(piv > 1e-15)) {
    free(r);
    SETERRQ(0, PETSC_ERR_LIB, "coupled operator singular to working precision (pivot ratio %g)", piv);
  }
  ModelVecScatter(x, r);
  free(r);
  h->its = 1;
  return PETSC_SUCCESS;
}
static PetscErrorCode pc_apply(Ref *h, Vec b, Vec x) { return h->pc->ops->apply(h->pc, b, x); }
static double nest_dot(Vec x, Vec y)
{
  double s = 0.;
  int    f, i;
  for (f = 0; f < 3; ++f)
    for (i = 0; i < x->sub[f]->n; ++i) s += x->sub[f]->a[i] * y->sub[f]->a[i];
  return s;
}
/* right-preconditioned GMRES(m), zero guess, modified Gram-Schmidt; hist = the residual norms (true residuals of x = M^-1 V y) */
static PetscErrorCode solve_gmres(Ref *h, NS ns, Vec b, Vec x)
{
  enum { M = 30 };
  Vec    V[M + 1], z, w, r;
  double H[M + 1][M], cs[M], sn[M], g[M + 1], y[M], rnorm, rnorm0;
  int    k, j, its = 0, done = 0;
  PetscCall(VecDuplicate(b, &z));
  PetscCall(VecDuplicate(b, &w));
  PetscCall(VecDuplicate(b, &r));
  for (k = 0; k <= M; ++k) PetscCall(VecDuplicate(b, &V[k]));
  PetscCall(VecSet(x, 0.));
  PetscCall(VecCopy(b, r));
  rnorm0 = rnorm = sqrt(nest_dot(r, r));
  h->nhist = 0, h->hist[h->nhist++] = rnorm;
  while (!done && rnorm > h->rtol * rnorm0 && its < h->maxit) {
    PetscCall(VecCopy(r, V[0]));
    PetscCall(VecScale(V[0], 1. / rnorm));
    memset(g, 0, sizeof(g));
    g[0] = rnorm;
    for (k = 0; k < M && its < h->maxit; ++k) {
      PetscCall(pc_apply(h, V[k], z));
      PetscCall(MatMult(ns->J, z, w));
      if (ns->nullspace) PetscCall(MatNullSpaceRemove(ns->nullspace, w));
      for (j = 0; j <= k; ++j) {
        H[j][k] = nest_dot(w, V[j]);
        PetscCall(VecAXPY(w, -H[j][k], V[j]));
      }
      H[k + 1][k] = sqrt(nest_dot(w, w));
      PetscCall(VecCopy(w, V[k + 1]));
      if (H[k + 1][k] > 0.) PetscCall(VecScale(V[k + 1], 1. / H[k + 1][k]));
      for (j = 0; j < k; ++j) {
        const double a = H[j][k], c = H[j + 1][k];
        H[j][k] = cs[j] * a + sn[j] * c, H[j + 1][k] = -sn[j] * a + cs[j] * c;
      }
      {
        const double a = H[k][k], c = H[k + 1][k], d = hypot(a, c);
        cs[k] = d > 0. ? a / d : 1., sn[k] = d > 0. ? c / d : 0.;
        H[k][k] = d, H[k + 1][k] = 0.;
        g[k + 1] = -sn[k] * g[k], g[k] = cs[k] * g[k];
      }
      ++its;
      rnorm = fabs(g[k + 1]);
      if (h->nhist < 256) h->hist[h->nhist++] = rnorm;
      if (rnorm <= h->rtol * rnorm0) {
        ++k;
        done = 1;
        break;
      }
    }
    for (j = k - 1; j >= 0; --j) {
      double s = g[j];
      int    l;
      for (l = j + 1; l < k; ++l) s -= H[j][l] * y[l];
      y[j] = s / H[j][j];
    }
    PetscCall(VecSet(w, 0.));
    for (j = 0; j < k; ++j) PetscCall(VecAXPY(w, y[j], V[j]));
    PetscCall(pc_apply(h, w, z)); /* KSPGMRES builds the solution with one more application of the preconditioner per cycle */
    PetscCall(VecAXPY(x, 1., z));
    if (!done) { /* restart: the true residual */
      PetscCall(MatMult(ns->J, x, w));
      PetscCall(VecWAXPY(r, -1., w, b));
      if (ns->nullspace) PetscCall(MatNullSpaceRemove(ns->nullspace, r));
      rnorm = sqrt(nest_dot(r, r));
    }
  }
  h->its = its;
  for (k = 0; k <= M; ++k) PetscCall(VecDestroy(&V[k]));
  PetscCall(VecDestroy(&z));
  PetscCall(VecDestroy(&w));
  PetscCall(VecDestroy(&r));
  return PETSC_SUCCESS;
}
static PetscErrorCode snes_solve(SNES snes, Vec bunused, Vec x)
{
  Ref *h  = (Ref *)snes->ctx;
  NS   ns = h->ns;
  (void)bunused;
  PetscCall(VecZeroEntries(x));                            /* FormInitialGuess_Private, nsbasic.c:146-151 */
  PetscCall(ns->ops->formfunction(ns, x, ns->r));          /* the b(x) of SNESSetPicard (nsbasic.c:249): the type fills f with b */
  if (ns->nullspace) PetscCall(MatNullSpaceRemove(ns->nullspace, ns->r)); /* PicardComputeFunction_Private, nsbasic.c:133-144 */
  PetscCall(ns->ops->formjacobian(ns, x, ns->J, NS_UPDATE_JACOBIAN));     /* FormJacobian_Private, nsbasic.c:113-120 */
  nest_get(h, ns->r, h->last_b);
  h->nhist = 0;
  if (h->mode == 0) return solve_exact(h, ns, ns->r, x);
  h->pc->mat = h->pc->pmat = ns->J;
  PetscCall(h->pc->ops->setup(h->pc)); /* PCSetUp_ABF runs once per solve: the operators changed */
  if (h->mode == 1) {
    h->its = 1;
    return pc_apply(h, ns->r, x);
  }
  return solve_gmres(h, ns, ns->r, x);
}

/* ------------------------------------------------------------------ the C interface (oracle/ref.py) */
const char *ref_last_error(void) { return ModelLastError(); }

static PetscErrorCode setup(Ref *h, const int *bctype, double rho, double mu, double dt)
{
  NS   ns;
  Vec  sub[3];
  IS   is[3];
  int  f, b;
  SNES snes;
  /* NSCreate + NSSetType(cnlinear) + NSSetMesh + parameters (nsbasic.c:15-79, nsopts.c) */
  ns = (NS)calloc(1, sizeof(*ns));
  ModelHeaderInit(ns, NS_CLASSID, "NS", NSCNLINEAR, NULL);
  h->ns           = ns;
  ns->mon_ctxs[0] = h;
  ns->rho = rho, ns->mu = mu, ns->dt = dt, ns->mesh = h->mesh;
  ns->reason = NS_CONVERGED_ITERATING;
  PetscCall(PetscCalloc1(2 * h->dim, &ns->bcs));
  for (b = 0; b < 2 * h->dim; ++b) {
    h->bctx[b].h = h, h->bctx[b].bnd = b;
    ns->bcs[b].type = (NSBoundaryConditionType)bctype[b];
    ns->bcs[b].velocity = bc_velocity, ns->bcs[b].ctx_velocity = &h->bctx[b];
    ns->bcs[b].pressure = bc_pressure, ns->bcs[b].ctx_pressure = &h->bctx[b];
    if (bctype[b] == NS_BC_PRESSURE_OUTLET) h->has_nullspace = 0;
  }
  PetscCall(NSCreate_CNLinear(ns));
  /* NSSetUp (nsbasic.c:153-274): fields, solution nest, Jacobian nest + formjacobian(INIT), work vectors, null space, SNES + PCABF,
     then the type's setup */
  for (f = 0; f < 3; ++f) {
    h->is[f].field = f, h->is[f].n = h->mesh->dm[field_dm[f]]->nglobal, is[f] = &h->is[f];
    PetscCall(MeshCreateGlobalVector(h->mesh, field_dm[f], &sub[f]));
    PetscCall(PetscObjectSetName((PetscObject)sub[f], field_name[f]));
  }
  ns->sol = ModelVecCreateNest(3, sub);
  for (f = 0; f < 3; ++f) PetscCall(VecDestroy(&sub[f]));
  PetscCall(MatCreateNest(0, 3, is, 3, is, NULL, &ns->J));
  PetscCall(ns->ops->formjacobian(ns, ns->x, ns->J, NS_INIT_JACOBIAN));
  PetscCall(MatCreateVecs(ns->J, &ns->x, &ns->r));
  if (h->has_nullspace) {
    Vec nv;
    PetscCall(MatCreateVecs(ns->J, NULL, &nv));
    PetscCall(VecSet(nv, 0.));
    PetscCall(VecSet(nv->sub[2], 1. / sqrt((double)nv->sub[2]->n)));
    PetscCall(MatNullSpaceCreate(0, PETSC_FALSE, 1, &nv, &ns->nullspace));
    PetscCall(VecDestroy(&nv));
    PetscCall(MatSetNullSpace(ns->J, ns->nullspace));
  }
  snes = (SNES)calloc(1, sizeof(*snes));
  ModelHeaderInit(snes, 23, "SNES", "picard-model", NULL);
  snes->ctx = h, snes->solve = snes_solve;
  ns->snes  = snes;
  h->pc     = (PC)calloc(1, sizeof(*h->pc));
  ModelHeaderInit(h->pc, PC_CLASSID, "PC", PCABF, NULL);
  PetscCall(PCCreate_ABF(h->pc));
  PetscCall(PCABFSetFields(h->pc, 0, 1, 2));
  PetscCall(ns->ops->setup(ns));
  ns->setupcalled = PETSC_TRUE;
  return PETSC_SUCCESS;
}

void *ref_create(int dim, const int *n, const double *xf0, const double *xf1, const double *xf2, const int *bctype, double rho, double mu, double dt, ref_bc_cb cb)
{
  Ref                *h = (Ref *)calloc(1, sizeof(*h));
  const double *const xf[3] = {xf0, xf1, xf2};
  int                 d;
  h->dim = dim, h->cb = cb, h->has_nullspace = 1, h->rtol = 1e-5, h->maxit = 10000;
  for (d = 0; d < 3; ++d) h->n[d] = d < dim ? n[d] : 1, h->per[d] = d < dim && bctype[2 * d] == NS_BC_PERIODIC;
  h->mesh = mesh_create(dim, h->n, h->per, xf);
  if (build_maps(h) || setup(h, bctype, rho, mu, dt)) return NULL;
  h->last_b = (double *)calloc((size_t)h->nsol, sizeof(double));
  return h;
}
void ref_sizes(void *vh, long *ncell, long nface[3])
{
  Ref *h = (Ref *)vh;
  *ncell = h->ncell;
  nface[0] = h->nface[0], nface[1] = h->nface[1], nface[2] = h->dim == 3 ? h->nface[2] : 0;
}
/* x: v, U_x, U_y, (U_z), p back to back; phalf separate (cnl->phalf, cnlinear.c:54) */
int ref_set_state(void *vh, const double *x, const double *phalf, int step, double t)
{
  Ref         *h   = (Ref *)vh;
  NS_CNLinear *cnl = (NS_CNLinear *)h->ns->data;
  long         i;
  nest_put(h, h->ns->sol, x);
  if (phalf)
    for (i = 0; i < h->ncell; ++i) cnl->phalf->a[h->map_p[i]] = phalf[i];
  h->ns->step = step, h->ns->t = t;
  return 0;
}
int ref_get_state(void *vh, double *x, double *phalf, int *step, double *t)
{
  Ref         *h   = (Ref *)vh;
  NS_CNLinear *cnl = (NS_CNLinear *)h->ns->data;
  long         i;
  nest_get(h, h->ns->sol, x);
  for (i = 0; i < h->ncell; ++i) phalf[i] = cnl->phalf->a[h->map_p[i]];
  *step = (int)h->ns->step, *t = h->ns->t;
  return 0;
}
/* NSStep (nsbasic.c:276-299) with SNESSolve in the given mode; ABF factor types as -ns_pc_abf_{schur,upper}_ainv_type */
int ref_step(void *vh, int mode, int schur_ainv, int upper_ainv, double rtol, int maxit, int *its, int *nhist, double *hist)
{
  Ref *h  = (Ref *)vh;
  NS   ns = h->ns;
  int  i;
  h->mode = mode, h->rtol = rtol, h->maxit = maxit;
  if (PCABFSetSchurComplementAinvType(h->pc, (PCABFAinvType)schur_ainv) || PCABFSetUpperTriangularAinvType(h->pc, (PCABFAinvType)upper_ainv)) return 1;
  if (!ns->sol0 && VecDuplicate(ns->sol, &ns->sol0)) return 1;
  if (VecCopy(ns->sol, ns->sol0)) return 1;
  if (ns->ops->step(ns)) return 1;
  if (ns->reason >= 0) ++ns->step, ns->t += ns->dt;
  if (its) *its = h->its;
  if (nhist) *nhist = h->nhist;
  if (hist)
    for (i = 0; i < h->nhist; ++i) hist[i] = h->hist[i];
  return 0;
}
/* inner KSPs of PCABF: 0 = the model's exact solves (the checker's default), 1 = GMRES(30) + ILU(0), rtol 1e-5: what serial PETSc
   runs when nobody configures them (petsc_model_ksp.c); rtol <= 0: that default.  Process-wide; counts = Krylov iterations since creation. */
void ref_set_inner_solvers(int iterative, double rtol) { ModelKSPSetDefaults(iterative), ModelKSPSetDefaultRtol(rtol); }
int  ref_inner_iterations(void *vh, long *mom, long *schur)
{
  Ref *h = (Ref *)vh;
  KSP  ka = NULL, ks = NULL;
  if (PCABFGetSubKSPs(h->pc, &ka, &ks) || !ka || !ks) return 1;
  *mom = ka->total_its, *schur = ks->total_its;
  return 0;
}
/* the right-hand side of the solve inside the last ref_step (null space removed, as SNES sees it) */
int ref_last_rhs(void *vh, double *b)
{
  Ref *h = (Ref *)vh;
  memcpy(b, h->last_b, sizeof(double) * (size_t)h->nsol);
  return 0;
}
/* NSFormFunction on the current state (sol0 <- sol first, as NSStep does): f = b, no null-space removal */
int ref_form_function(void *vh, double *b)
{
  Ref *h  = (Ref *)vh;
  NS   ns = h->ns;
  if (!ns->sol0 && VecDuplicate(ns->sol, &ns->sol0)) return 1;
  if (VecCopy(ns->sol, ns->sol0)) return 1;
  if (ns->ops->formfunction(ns, ns->x, ns->r)) return 1;
  nest_get(h, ns->r, b);
  return 0;
}
/* one application of the reference's PCABF (exact inner solves) to a canonical vector, with the operators of the last ref_step */
int ref_abf_apply(void *vh, int schur_ainv, int upper_ainv, const double *b, double *x)
{
  Ref *h = (Ref *)vh;
  Vec  vb, vx;
  int  rc;
  if (PCABFSetSchurComplementAinvType(h->pc, (PCABFAinvType)schur_ainv) || PCABFSetUpperTriangularAinvType(h->pc, (PCABFAinvType)upper_ainv)) return 1;
  if (VecDuplicate(h->ns->r, &vb) || VecDuplicate(h->ns->r, &vx)) return 1;
  nest_put(h, vb, b);
  h->pc->mat = h->pc->pmat = h->ns->J;
  rc = h->pc->ops->setup(h->pc) || h->pc->ops->apply(h->pc, vb, vx);
  if (!rc) nest_get(h, vx, x);
  VecDestroy(&vb), VecDestroy(&vx);
  return rc;
}
/* y = J x with the Jacobian of the last ref_step (canonical vectors) */
int ref_apply_jacobian(void *vh, const double *x, double *y)
{
  Ref *h = (Ref *)vh;
  Vec  vx, vy;
  int  rc;
  if (VecDuplicate(h->ns->r, &vx) || VecDuplicate(h->ns->r, &vy)) return 1;
  nest_put(h, vx, x);
  rc = MatMult(h->ns->J, vx, vy);
  if (!rc) nest_get(h, vy, y);
  VecDestroy(&vx), VecDestroy(&vy);
  return rc;
}
/* one operator as coordinate triplets in canonical numbering (row / column spaces: 0 = v, 1 = U, 2 = p).  Names: the blocks of J
 * "A" "G" "negT" "I" "negR" "D" and the matrices composed on it, "L" ("Laplacian") and "Gst" ("StaggeredGradient").  Call with
 * rows = NULL for the count.  Explicitly stored zeros are returned as such. */
static int *inverse_map(Ref *h, int space, int *nout)
{
  DM   dm = h->mesh->dm[field_dm[space]];
  int *inv = (int *)malloc(sizeof(int) * (size_t)dm->nglobal), d;
  long i, o = 0;
  for (i = 0; i < dm->nglobal; ++i) inv[i] = -1;
  if (space == 0)
    for (i = 0; i < h->ncell * h->dim; ++i) inv[h->map_v[i]] = (int)i;
  else if (space == 2)
    for (i = 0; i < h->ncell; ++i) inv[h->map_p[i]] = (int)i;
  else
    for (d = 0; d < h->dim; ++d)
      for (i = 0; i < h->nface[d]; ++i) inv[h->map_U[d][i]] = (int)o++;
  *nout = dm->nglobal;
  return inv;
}
long ref_matrix(void *vh, const char *name, int *rows, int *cols, double *vals)
{
  Ref *h = (Ref *)vh;
  Mat  J = h->ns->J, A = NULL;
  int  rs = 0, cs = 0, *ri, *ci, nr, nc, i, q;
  long cnt = 0;
  if (!strcmp(name, "A")) A = J->blk[0][0], rs = 0, cs = 0;
  else if (!strcmp(name, "G")) A = J->blk[0][2], rs = 0, cs = 2;
  else if (!strcmp(name, "negT")) A = J->blk[1][0], rs = 1, cs = 0;
  else if (!strcmp(name, "I")) A = J->blk[1][1], rs = 1, cs = 1;
  else if (!strcmp(name, "negR")) A = J->blk[1][2], rs = 1, cs = 2;
  else if (!strcmp(name, "D")) A = J->blk[2][1], rs = 2, cs = 1;
  else if (!strcmp(name, "L")) PetscObjectQuery((PetscObject)J, "Laplacian", (PetscObject *)&A), rs = 0, cs = 0;
  else if (!strcmp(name, "Gst")) PetscObjectQuery((PetscObject)J, "StaggeredGradient", (PetscObject *)&A), rs = 1, cs = 2;
  if (!A) return -1;
  ri = inverse_map(h, rs, &nr), ci = inverse_map(h, cs, &nc);
  if (A->m != nr || A->n != nc) return -2;
  for (i = 0; i < A->m; ++i)
    for (q = 0; q < A->rn[i]; ++q) {
      if (rows) {
        if (ri[i] < 0 || ci[A->rc[i][q]] < 0) return -3; /* an entry outside the canonical fields */
        rows[cnt] = ri[i], cols[cnt] = ci[A->rc[i][q]], vals[cnt] = A->rv[i][q];
      }
      ++cnt;
    }
  free(ri), free(ci);
  return cnt;
}
void ref_destroy(void *vh)
{
  Ref *h = (Ref *)vh;
  NS   ns;
  int  d;
  if (!h) return;
  ns = h->ns;
  if (h->pc) {
    h->pc->ops->destroy(h->pc);
    ModelHeaderFree(h->pc);
    free(h->pc);
  }
  if (ns) {
    if (ns->ops->destroy) ns->ops->destroy(ns);
    VecDestroy(&ns->sol), VecDestroy(&ns->sol0), VecDestroy(&ns->x), VecDestroy(&ns->r);
    MatDestroy(&ns->J);
    MatNullSpaceDestroy(&ns->nullspace);
    if (ns->snes) ModelHeaderFree(ns->snes), free(ns->snes);
    free(ns->bcs);
    ModelHeaderFree(ns);
    free(ns);
  }
  mesh_destroy(h->mesh);
  free(h->map_v), free(h->map_p), free(h->last_b);
  for (d = 0; d < 3; ++d) free(h->map_U[d]);
  free(h);
}
