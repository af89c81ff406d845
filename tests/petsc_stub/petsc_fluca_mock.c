/* tests/petsc_stub/petsc_fluca_mock.c -- TEST INFRASTRUCTURE ONLY.
 *
 * A single-rank functional model of the PETSc / Fluca API subset declared in petsc_fluca_stub.h, so that glue/nsb200.c --
 * which cannot meet a real PETSc in this image -- is at least COMPILED, LINKED and RUN: tests/test_glue_mock.py drives it through
 * tests/c/ns_b200_glue_driver.c against the host-emulation build of the library (CPU) and the CUDA library (B200) and compares
 * the contents of ns->sol with the oracle.
 *
 * What is modelled, from the PETSc manual pages (nothing here is PETSc or Fluca source):
 *   * DMStag's element-wise storage: per element the strata in the order vertex, edges, faces, element (2-D: down-left, down,
 *     left, element; 3-D: back-down-left, back-down, back-left, back, down-left, down, left, element), one extra PARTIAL element
 *     at the upper end of every non-periodic direction that carries only the entries lying on its lower boundary, ghost elements
 *     (stencil width 1) at both ends of a periodic direction; local arrays indexed with GLOBAL element numbers.
 *     Entries that do not exist in a partial element read as NaN, so a glue that touches them fails the comparison.
 *   * 1-D product coordinates with the slots LEFT = 0, ELEMENT = 1, RIGHT = 2 (the next element's LEFT).
 *   * the object state counter: every write access to a Vec increases it; restoring a sub-vector of a VecNest increases the
 *     state of the nest (switchable, MockSetNestRestoreBumpsState).
 *   * the NS base class: the order of operations of NSSetUp / NSStep / NSViewSolution / NSLoadSolution
 *     (fluca/src/ns/interface/nsbasic.c:153-299, nssol.c:130-204), with and without the matrix-free hooks of glue/patches/.
 * What is not: more than one rank (MPI_Comm_size is 1), matrices (only the placeholders NSFormJacobian(INIT) sets), CGNS.
 * Arguments PETSc documents as "ignored in lower dimensions" come back as garbage here on purpose.
 */
#include "petsc_fluca_mock.h"
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#define GARBAGE (-777)

/* ------------------------------------------------------------------ memory, errors, options */
static long live_allocations = 0;
long        mock_ndm_global_to_local = 0;
static int  nest_restore_bumps = 1;

static void *xcalloc(size_t n)
{
  void *p = calloc(1, n ? n : 1);
  if (!p) abort();
  ++live_allocations;
  return p;
}
static void xfree(void *p)
{
  if (p) {
    --live_allocations;
    free(p);
  }
}
long MockLiveAllocations(void) { return live_allocations; }
void MockSetNestRestoreBumpsState(int on) { nest_restore_bumps = on; }

PetscErrorCode PetscMallocStub(size_t n, void *pp)
{
  void *p = xcalloc(n);
  memset(p, 0xA5, n); /* PetscMalloc1 does not clear */
  *(void **)pp = p;
  return PETSC_SUCCESS;
}
PetscErrorCode PetscCallocStub(size_t n, void *pp)
{
  *(void **)pp = xcalloc(n);
  return PETSC_SUCCESS;
}
PetscErrorCode PetscFreeStub(void *p)
{
  xfree(p);
  return PETSC_SUCCESS;
}
PetscErrorCode PetscMemzero(void *p, size_t n)
{
  memset(p, 0, n);
  return PETSC_SUCCESS;
}
PetscErrorCode PetscArraycmpStub(const void *a, const void *b, size_t n, PetscBool *e)
{
  *e = memcmp(a, b, n) ? PETSC_FALSE : PETSC_TRUE;
  return PETSC_SUCCESS;
}
PetscErrorCode PetscArraycpyStub(void *a, const void *b, size_t n)
{
  memcpy(a, b, n);
  return PETSC_SUCCESS;
}
PetscErrorCode PetscErrorStub(MPI_Comm comm, int err, const char *fmt, ...)
{
  va_list ap;
  (void)comm;
  fprintf(stderr, "[mock PETSc] error %d: ", err);
  va_start(ap, fmt);
  vfprintf(stderr, fmt, ap);
  va_end(ap);
  fputc('\n', stderr);
  return err ? err : 1;
}
#define MockCheck(cond, ...) \
  do { \
    if (!(cond)) return PetscErrorStub(0, 99, __VA_ARGS__); \
  } while (0)

static int mock_info = 0;
PetscErrorCode PetscInfoStub(void *obj, const char *fmt, ...)
{
  (void)obj;
  if (mock_info) {
    va_list ap;
    va_start(ap, fmt);
    vfprintf(stdout, fmt, ap);
    va_end(ap);
  }
  return PETSC_SUCCESS;
}
PetscErrorCode PetscPrintf(MPI_Comm comm, const char fmt[], ...)
{
  va_list ap;
  (void)comm;
  va_start(ap, fmt);
  vfprintf(stdout, fmt, ap);
  va_end(ap);
  return PETSC_SUCCESS;
}

#define MAXOPT 64
static struct {
  char name[64], value[64];
  int  used;
} options[MAXOPT];
static int noptions = 0;

PetscErrorCode MockOptionsSetValue(const char name[], const char value[])
{
  int i;
  for (i = 0; i < noptions; ++i)
    if (!strcmp(options[i].name, name)) break;
  MockCheck(i < MAXOPT, "too many options");
  if (i == noptions) ++noptions;
  snprintf(options[i].name, sizeof(options[i].name), "%s", name);
  snprintf(options[i].value, sizeof(options[i].value), "%s", value ? value : "");
  options[i].used = 0;
  if (!strcmp(name, "-info")) mock_info = 1;
  return PETSC_SUCCESS;
}
PetscErrorCode MockOptionsClear(void)
{
  noptions  = 0;
  mock_info = 0;
  return PETSC_SUCCESS;
}
static const char *option_find(const char *name)
{
  int i;
  for (i = 0; i < noptions; ++i)
    if (!strcmp(options[i].name, name)) {
      options[i].used = 1;
      return options[i].value;
    }
  return NULL;
}
PetscErrorCode PetscOptionsIntStub(PetscOptionItems o, const char *name, const char *text, const char *man, PetscInt cur, PetscInt *val, PetscBool *set)
{
  const char *v = option_find(name);
  (void)o, (void)text, (void)man, (void)cur;
  if (v) *val = (PetscInt)atol(v);
  if (set) *set = v ? PETSC_TRUE : PETSC_FALSE;
  return PETSC_SUCCESS;
}
PetscErrorCode PetscOptionsRealStub(PetscOptionItems o, const char *name, const char *text, const char *man, PetscReal cur, PetscReal *val, PetscBool *set)
{
  const char *v = option_find(name);
  (void)o, (void)text, (void)man, (void)cur;
  if (v) *val = atof(v);
  if (set) *set = v ? PETSC_TRUE : PETSC_FALSE;
  return PETSC_SUCCESS;
}
PetscErrorCode PetscOptionsBoolStub(PetscOptionItems o, const char *name, const char *text, const char *man, PetscBool cur, PetscBool *val, PetscBool *set)
{
  const char *v = option_find(name);
  (void)o, (void)text, (void)man, (void)cur;
  if (v) *val = (!v[0] || !strcmp(v, "1") || !strcmp(v, "true") || !strcmp(v, "yes")) ? PETSC_TRUE : PETSC_FALSE; /* a bare flag is true */
  if (set) *set = v ? PETSC_TRUE : PETSC_FALSE;
  return PETSC_SUCCESS;
}
static int same_nocase(const char *a, const char *b)
{
  for (; *a && *b; ++a, ++b)
    if ((*a | 32) != (*b | 32)) return 0;
  return !*a && !*b;
}
PetscErrorCode PetscOptionsEnumStub(PetscOptionItems o, const char *name, const char *text, const char *man, const char *const *list, PetscEnum cur, PetscEnum *val, PetscBool *set)
{
  const char *v = option_find(name);
  int         n = 0, i;
  (void)o, (void)text, (void)man, (void)cur;
  while (list[n]) ++n;
  n -= 2; /* the value names are followed by the enum's type name and prefix */
  if (v) {
    for (i = 0; i < n; ++i)
      if (same_nocase(v, list[i])) break;
    MockCheck(i < n, "unknown value %s for option %s", v, name);
    *val = (PetscEnum)i;
  }
  if (set) *set = v ? PETSC_TRUE : PETSC_FALSE;
  return PETSC_SUCCESS;
}
const char *const PCABFAinvTypes[] = {"ID", "DIAG", "ROWSUM", "PCABFAinvType", "", NULL}; /* the values of flucans.h:99-104 */

/* ------------------------------------------------------------------ MPI: one rank */
int MPI_Comm_rank(MPI_Comm c, int *r) { return (void)c, *r = 0, 0; }
int MPI_Comm_size(MPI_Comm c, int *s) { return (void)c, *s = 1, 0; }
int MPI_Bcast(void *b, int n, MPI_Datatype t, int root, MPI_Comm c) { return (void)b, (void)n, (void)t, (void)root, (void)c, 0; }
int MPI_Allreduce(const void *s, void *r, int n, MPI_Datatype t, MPI_Op op, MPI_Comm c)
{
  (void)op, (void)c;
  if (s != MPI_IN_PLACE) memcpy(r, s, (size_t)n * (t == MPI_DOUBLE ? sizeof(double) : t == MPI_INT ? sizeof(int) : 1));
  return 0;
}

/* ------------------------------------------------------------------ objects */
struct _p_IS {
  struct _p_PetscObject hdr;
  int                   field;
};
struct _p_Mat {
  struct _p_PetscObject hdr;
  PetscInt              rows, cols;
  Mat                   block[3][3];
  int                   nest, assembled;
};
struct _p_Vec {
  struct _p_PetscObject hdr;
  DM                    dm;
  int                   local;
  size_t                n;
  double               *a;
  int                   nsub;
  Vec                   sub[4];
  void                 *table[3]; /* pointer tables handed out by DMStagVecGetArray */
  int                   array_out;
};
#define MAXLOC 8
struct _p_DM {
  struct _p_PetscObject hdr;
  Mesh                  mesh;
  int                   dim, N[3], per[3];
  int                   dof[4]; /* per stratum: vertices, edges, faces, elements (2-D: vertices, faces, elements) */
  int                   epe;    /* entries per element */
  int                   gs[3], gn[3], on[3]; /* ghost start / count (local), owned count incl. the partial element (global) */
  int                   nloc;
  DMStagStencilLocation loc[MAXLOC];
  int                   locmask[MAXLOC], locoff[MAXLOC], locdof[MAXLOC];
};
struct _p_Mesh {
  struct _p_PetscObject hdr;
  int                   dim, N[3], per[3];
  DM                    dm[3];
  double               *coord[3];
  double              **ctab[3];
};
struct _p_PetscViewer {
  struct _p_PetscObject hdr;
  FILE                 *f;
  struct stored {
    char           name[64];
    size_t         n;
    double        *a;
    struct stored *next;
  } *store;
  PetscInt  step;
  PetscReal time;
};

static void header(void *obj, const char *cls, const char *type)
{
  struct _p_PetscObject *h = (struct _p_PetscObject *)obj;
  h->class_name = cls;
  snprintf(h->type_name, sizeof(h->type_name), "%s", type);
  h->refct = 1;
}
MPI_Comm       PetscObjectComm(PetscObject o) { return o->comm; }
PetscErrorCode PetscObjectGetComm(PetscObject o, MPI_Comm *c) { return *c = o->comm, PETSC_SUCCESS; }
PetscErrorCode PetscObjectTypeCompare(PetscObject o, const char type[], PetscBool *m) { return *m = strcmp(o->type_name, type) ? PETSC_FALSE : PETSC_TRUE, PETSC_SUCCESS; }
PetscErrorCode PetscObjectStateGet(PetscObject o, PetscObjectState *s) { return *s = o->state, PETSC_SUCCESS; }
PetscErrorCode PetscObjectSetName(PetscObject o, const char name[]) { return snprintf(o->name, sizeof(o->name), "%s", name), PETSC_SUCCESS; }

/* ------------------------------------------------------------------ Vec */
static Vec vec_new(DM dm, int local, size_t n)
{
  Vec v = (Vec)xcalloc(sizeof(*v));
  header(v, "Vec", "seq");
  v->dm = dm, v->local = local, v->n = n;
  v->a = (double *)xcalloc(sizeof(double) * n);
  return v;
}
static Vec nest_new(int nsub, Vec sub[])
{
  Vec v = (Vec)xcalloc(sizeof(*v));
  int i;
  header(v, "Vec", "nest");
  v->nsub = nsub;
  for (i = 0; i < nsub; ++i) v->sub[i] = sub[i], ++sub[i]->hdr.refct;
  return v;
}
PetscErrorCode VecDestroy(Vec *pv)
{
  Vec v = *pv;
  int i;
  if (!v) return PETSC_SUCCESS;
  *pv = NULL;
  if (--v->hdr.refct > 0) return PETSC_SUCCESS;
  MockCheck(!v->array_out, "VecDestroy with an array checked out");
  for (i = 0; i < v->nsub; ++i) PetscCall(VecDestroy(&v->sub[i]));
  xfree(v->a);
  xfree(v);
  return PETSC_SUCCESS;
}
PetscErrorCode VecSet(Vec v, PetscScalar a)
{
  size_t i;
  int    s;
  for (s = 0; s < v->nsub; ++s) PetscCall(VecSet(v->sub[s], a));
  for (i = 0; i < v->n; ++i) v->a[i] = a;
  ++v->hdr.state;
  return PETSC_SUCCESS;
}
PetscErrorCode VecZeroEntries(Vec v) { return VecSet(v, 0.); }
PetscErrorCode VecScale(Vec v, PetscScalar a)
{
  size_t i;
  int    s;
  for (s = 0; s < v->nsub; ++s) PetscCall(VecScale(v->sub[s], a));
  for (i = 0; i < v->n; ++i) v->a[i] *= a;
  ++v->hdr.state;
  return PETSC_SUCCESS;
}
PetscErrorCode VecCopy(Vec x, Vec y)
{
  int s;
  MockCheck(x->nsub == y->nsub && x->n == y->n, "VecCopy: layouts differ");
  for (s = 0; s < x->nsub; ++s) PetscCall(VecCopy(x->sub[s], y->sub[s]));
  if (x->n) memcpy(y->a, x->a, sizeof(double) * x->n);
  ++y->hdr.state;
  return PETSC_SUCCESS;
}
PetscErrorCode VecDuplicate(Vec x, Vec *y)
{
  if (x->nsub) {
    Vec sub[4];
    int s;
    for (s = 0; s < x->nsub; ++s) PetscCall(VecDuplicate(x->sub[s], &sub[s]));
    *y = nest_new(x->nsub, sub);
    for (s = 0; s < x->nsub; ++s) PetscCall(VecDestroy(&sub[s]));
  } else {
    *y = vec_new(x->dm, x->local, x->n);
    snprintf((*y)->hdr.name, sizeof((*y)->hdr.name), "%s", x->hdr.name);
  }
  return PETSC_SUCCESS;
}
/* a borrowed sub-vector of a nest; its state at check-out is remembered in `array_out` of nobody: the nest bumps on restore */
static PetscObjectState sub_state_at_get[4];
PetscErrorCode VecGetSubVector(Vec v, IS is, Vec *sub)
{
  MockCheck(v->nsub > 0, "VecGetSubVector: the mock only splits nests");
  MockCheck(is && is->field >= 0 && is->field < v->nsub, "VecGetSubVector: bad IS");
  *sub                       = v->sub[is->field];
  sub_state_at_get[is->field] = (*sub)->hdr.state;
  return PETSC_SUCCESS;
}
PetscErrorCode VecRestoreSubVector(Vec v, IS is, Vec *sub)
{
  MockCheck(v->nsub > 0 && is && *sub == v->sub[is->field], "VecRestoreSubVector: not the vector that was handed out");
  if (nest_restore_bumps || (*sub)->hdr.state != sub_state_at_get[is->field]) ++v->hdr.state;
  *sub = NULL;
  return PETSC_SUCCESS;
}

/* ------------------------------------------------------------------ Mat: only the placeholders of NSFormJacobian(INIT) */
PetscErrorCode MatCreateConstantDiagonal(MPI_Comm c, PetscInt m, PetscInt n, PetscInt M, PetscInt N, PetscScalar d, Mat *A)
{
  (void)c, (void)M, (void)N, (void)d;
  *A = (Mat)xcalloc(sizeof(**A));
  header(*A, "Mat", "constantdiagonal");
  (*A)->rows = m, (*A)->cols = n;
  return PETSC_SUCCESS;
}
PetscErrorCode MatDestroy(Mat *pA)
{
  Mat A = *pA;
  int i, j;
  if (!A) return PETSC_SUCCESS;
  *pA = NULL;
  if (--A->hdr.refct > 0) return PETSC_SUCCESS;
  for (i = 0; i < 3; ++i)
    for (j = 0; j < 3; ++j) PetscCall(MatDestroy(&A->block[i][j]));
  xfree(A);
  return PETSC_SUCCESS;
}
PetscErrorCode MatNestSetSubMat(Mat J, PetscInt i, PetscInt j, Mat B)
{
  MockCheck(J->nest && i >= 0 && i < 3 && j >= 0 && j < 3, "MatNestSetSubMat: not a 3x3 nest / bad block");
  PetscCall(MatDestroy(&J->block[i][j]));
  J->block[i][j] = B, ++B->hdr.refct;
  return PETSC_SUCCESS;
}
PetscErrorCode MatAssemblyBegin(Mat A, MatAssemblyType t) { return (void)A, (void)t, PETSC_SUCCESS; }
PetscErrorCode MatAssemblyEnd(Mat A, MatAssemblyType t) { return (void)t, A->assembled = 1, PETSC_SUCCESS; }

/* ------------------------------------------------------------------ DMStag */
static const struct {
  DMStagStencilLocation loc;
  int                   mask, stratum;
} loc2d[4] = {{DMSTAG_DOWN_LEFT, 3, 0}, {DMSTAG_DOWN, 2, 1}, {DMSTAG_LEFT, 1, 1}, {DMSTAG_ELEMENT, 0, 2}},
  loc3d[8] = {{DMSTAG_BACK_DOWN_LEFT, 7, 0}, {DMSTAG_BACK_DOWN, 6, 1}, {DMSTAG_BACK_LEFT, 5, 1}, {DMSTAG_BACK, 4, 2}, {DMSTAG_DOWN_LEFT, 3, 1}, {DMSTAG_DOWN, 2, 2}, {DMSTAG_LEFT, 1, 2}, {DMSTAG_ELEMENT, 0, 3}};

static DM dm_new(Mesh mesh, int d0, int d1, int d2, int d3)
{
  DM  dm = (DM)xcalloc(sizeof(*dm));
  int d, l;
  header(dm, "DM", "stag");
  dm->mesh = mesh, dm->dim = mesh->dim;
  dm->dof[0] = d0, dm->dof[1] = d1, dm->dof[2] = d2, dm->dof[3] = d3;
  for (d = 0; d < 3; ++d) {
    dm->N[d]   = d < dm->dim ? mesh->N[d] : 1;
    dm->per[d] = d < dm->dim ? mesh->per[d] : 1; /* an unused direction has no partial element and no ghosts */
    dm->gs[d]  = d < dm->dim && dm->per[d] ? -1 : 0;
    dm->gn[d]  = d < dm->dim ? dm->N[d] + (dm->per[d] ? 2 : 1) : 1;
    dm->on[d]  = d < dm->dim ? dm->N[d] + (dm->per[d] ? 0 : 1) : 1;
  }
  dm->nloc = dm->dim == 2 ? 4 : 8;
  for (l = 0; l < dm->nloc; ++l) {
    dm->loc[l]     = dm->dim == 2 ? loc2d[l].loc : loc3d[l].loc;
    dm->locmask[l] = dm->dim == 2 ? loc2d[l].mask : loc3d[l].mask;
    dm->locdof[l]  = dm->dof[dm->dim == 2 ? loc2d[l].stratum : loc3d[l].stratum];
    dm->locoff[l]  = dm->epe;
    dm->epe += dm->locdof[l];
  }
  return dm;
}
/* does location l exist in the element with global index g? (a partial element holds only what lies on its lower boundary) */
static int entry_exists(DM dm, int l, const int g[3])
{
  int d;
  for (d = 0; d < dm->dim; ++d)
    if (!dm->per[d] && g[d] == dm->N[d] && !(dm->locmask[l] >> d & 1)) return 0;
  return 1;
}
PetscErrorCode DMGetDimension(DM dm, PetscInt *dim) { return *dim = dm->dim, PETSC_SUCCESS; }
PetscErrorCode DMStagGetCorners(DM dm, PetscInt *x, PetscInt *y, PetscInt *z, PetscInt *m, PetscInt *n, PetscInt *p, PetscInt *ex, PetscInt *ey, PetscInt *ez)
{
  if (x) *x = 0;
  if (y) *y = 0;
  if (z) *z = dm->dim == 3 ? 0 : GARBAGE;
  if (m) *m = dm->N[0];
  if (n) *n = dm->N[1];
  if (p) *p = dm->dim == 3 ? dm->N[2] : GARBAGE;
  if (ex) *ex = dm->per[0] ? 0 : 1;
  if (ey) *ey = dm->per[1] ? 0 : 1;
  if (ez) *ez = dm->dim == 3 ? (dm->per[2] ? 0 : 1) : GARBAGE;
  return PETSC_SUCCESS;
}
PetscErrorCode DMStagGetLocationSlot(DM dm, DMStagStencilLocation loc, PetscInt c, PetscInt *slot)
{
  int l;
  for (l = 0; l < dm->nloc; ++l)
    if (dm->loc[l] == loc) break;
  MockCheck(l < dm->nloc, "DMStagGetLocationSlot: location %d is not stored by an element of a %d-D DMStag (ask the neighbour for RIGHT/UP/FRONT)", (int)loc, dm->dim);
  MockCheck(c >= 0 && c < dm->locdof[l], "DMStagGetLocationSlot: component %d but the stratum has %d dof", (int)c, dm->locdof[l]);
  *slot = dm->locoff[l] + c;
  return PETSC_SUCCESS;
}
PetscErrorCode DMStagGetEntries(DM dm, PetscInt *entries)
{
  int g[3], l, n = 0;
  for (g[2] = 0; g[2] < dm->on[2]; ++g[2])
    for (g[1] = 0; g[1] < dm->on[1]; ++g[1])
      for (g[0] = 0; g[0] < dm->on[0]; ++g[0])
        for (l = 0; l < dm->nloc; ++l)
          if (entry_exists(dm, l, g)) n += dm->locdof[l];
  *entries = n;
  return PETSC_SUCCESS;
}
static size_t global_size(DM dm) { return (size_t)dm->on[0] * dm->on[1] * dm->on[2] * dm->epe; }
static size_t local_size(DM dm) { return (size_t)dm->gn[0] * dm->gn[1] * dm->gn[2] * dm->epe; }
PetscErrorCode DMGetLocalVector(DM dm, Vec *l) { return *l = vec_new(dm, 1, local_size(dm)), PETSC_SUCCESS; }
PetscErrorCode DMRestoreLocalVector(DM dm, Vec *l)
{
  MockCheck(*l && (*l)->dm == dm && (*l)->local, "DMRestoreLocalVector: not a local vector of this DM");
  return VecDestroy(l);
}
/* loop over the local (ghosted) elements: li = local index, g = global index, w = the owned element a ghost is a copy of (-1: none) */
#define FOR_LOCAL_ELEMENTS(dm) \
  for (li[2] = 0; li[2] < (dm)->gn[2]; ++li[2]) \
    for (li[1] = 0; li[1] < (dm)->gn[1]; ++li[1]) \
      for (li[0] = 0; li[0] < (dm)->gn[0]; ++li[0])
static void element_indices(DM dm, const int li[3], int g[3], int w[3], int *ghost)
{
  int d;
  *ghost = 0;
  for (d = 0; d < 3; ++d) {
    g[d] = li[d] + dm->gs[d];
    w[d] = g[d];
    if (d < dm->dim && dm->per[d] && (g[d] < 0 || g[d] >= dm->N[d])) w[d] = (g[d] + dm->N[d]) % dm->N[d], *ghost = 1;
  }
}
PetscErrorCode DMGlobalToLocal(DM dm, Vec g, InsertMode mode, Vec l)
{
  int li[3], gi[3], w[3], ghost, lc, c;
  MockCheck(g && l && g->dm == dm && l->dm == dm, "DMGlobalToLocal: vector does not belong to this DM");
  MockCheck(!g->local && l->local && mode == INSERT_VALUES, "DMGlobalToLocal: wrong kind of vector / mode");
  ++mock_ndm_global_to_local;
  FOR_LOCAL_ELEMENTS(dm)
  {
    element_indices(dm, li, gi, w, &ghost);
    {
      double       *dst = l->a + (((size_t)li[2] * dm->gn[1] + li[1]) * dm->gn[0] + li[0]) * dm->epe;
      const double *src = g->a + (((size_t)w[2] * dm->on[1] + w[1]) * dm->on[0] + w[0]) * dm->epe;
      for (lc = 0; lc < dm->nloc; ++lc)
        for (c = 0; c < dm->locdof[lc]; ++c) dst[dm->locoff[lc] + c] = entry_exists(dm, lc, gi) ? src[dm->locoff[lc] + c] : NAN;
    }
  }
  ++l->hdr.state;
  return PETSC_SUCCESS;
}
PetscErrorCode DMLocalToGlobal(DM dm, Vec l, InsertMode mode, Vec g)
{
  int li[3], gi[3], w[3], ghost, lc, c;
  MockCheck(g && l && g->dm == dm && l->dm == dm, "DMLocalToGlobal: vector does not belong to this DM");
  MockCheck(!g->local && l->local, "DMLocalToGlobal: wrong kind of vector");
  FOR_LOCAL_ELEMENTS(dm)
  {
    element_indices(dm, li, gi, w, &ghost);
    if (ghost && mode == INSERT_VALUES) continue; /* ghost values are dropped on insertion, summed into their owner on ADD_VALUES */
    {
      const double *src = l->a + (((size_t)li[2] * dm->gn[1] + li[1]) * dm->gn[0] + li[0]) * dm->epe;
      double       *dst = g->a + (((size_t)w[2] * dm->on[1] + w[1]) * dm->on[0] + w[0]) * dm->epe;
      for (lc = 0; lc < dm->nloc; ++lc)
        for (c = 0; c < dm->locdof[lc]; ++c) {
          if (!entry_exists(dm, lc, gi)) {
            MockCheck(src[dm->locoff[lc] + c] == 0. || isnan(src[dm->locoff[lc] + c]), "DMLocalToGlobal: a value was written to an entry that does not exist in a partial element");
            continue;
          }
          if (mode == ADD_VALUES) dst[dm->locoff[lc] + c] += src[dm->locoff[lc] + c];
          else dst[dm->locoff[lc] + c] = src[dm->locoff[lc] + c];
        }
    }
  }
  ++g->hdr.state;
  return PETSC_SUCCESS;
}
static PetscErrorCode stag_array(DM dm, Vec v, void *out)
{
  const size_t n0 = (size_t)dm->gn[0], n1 = (size_t)dm->gn[1], n2 = (size_t)dm->gn[2];
  size_t       i, j, k;
  MockCheck(v && v->dm == dm && v->local, "DMStagVecGetArray: needs a LOCAL vector of this DM");
  MockCheck(!v->array_out, "DMStagVecGetArray: array already checked out");
  {
    double **cells = (double **)xcalloc(sizeof(double *) * n0 * n1 * n2);
    for (i = 0; i < n0 * n1 * n2; ++i) cells[i] = v->a + i * dm->epe;
    v->table[0] = cells;
    if (dm->dim == 2) {
      double ***rows = (double ***)xcalloc(sizeof(double **) * n1);
      for (j = 0; j < n1; ++j) rows[j] = cells + j * n0 - dm->gs[0];
      v->table[1]      = rows;
      *(double ****)out = rows - dm->gs[1];
    } else {
      double  ***rows   = (double ***)xcalloc(sizeof(double **) * n1 * n2);
      double ****planes = (double ****)xcalloc(sizeof(double ***) * n2);
      for (j = 0; j < n1 * n2; ++j) rows[j] = cells + j * n0 - dm->gs[0];
      for (k = 0; k < n2; ++k) planes[k] = rows + k * n1 - dm->gs[1];
      v->table[1] = rows, v->table[2] = planes;
      *(double *****)out = planes - dm->gs[2];
    }
  }
  v->array_out = 1;
  return PETSC_SUCCESS;
}
static PetscErrorCode stag_array_restore(DM dm, Vec v, void *out, int wrote)
{
  int t;
  MockCheck(v && v->dm == dm && v->array_out, "DMStagVecRestoreArray: no array checked out");
  for (t = 0; t < 3; ++t) xfree(v->table[t]), v->table[t] = NULL;
  v->array_out   = 0;
  *(void **)out = NULL;
  if (wrote) ++v->hdr.state;
  return PETSC_SUCCESS;
}
PetscErrorCode DMStagVecGetArray(DM dm, Vec v, void *a) { return stag_array(dm, v, a); }
PetscErrorCode DMStagVecGetArrayRead(DM dm, Vec v, void *a) { return stag_array(dm, v, a); }
PetscErrorCode DMStagVecRestoreArray(DM dm, Vec v, void *a) { return stag_array_restore(dm, v, a, 1); }
PetscErrorCode DMStagVecRestoreArrayRead(DM dm, Vec v, void *a) { return stag_array_restore(dm, v, a, 0); }

PetscErrorCode DMStagGetProductCoordinateArraysRead(DM dm, void *ax, void *ay, void *az)
{
  Mesh m = dm->mesh;
  if (ax) *(double ***)ax = m->ctab[0] - dm->gs[0];
  if (ay) *(double ***)ay = m->ctab[1] - dm->gs[1];
  if (az && m->dim == 3) *(double ***)az = m->ctab[2] - dm->gs[2]; /* untouched in 2-D */
  return PETSC_SUCCESS;
}
PetscErrorCode DMStagRestoreProductCoordinateArraysRead(DM dm, void *ax, void *ay, void *az)
{
  if (ax) *(void **)ax = NULL;
  if (ay) *(void **)ay = NULL;
  if (az && dm->dim == 3) *(void **)az = NULL;
  return PETSC_SUCCESS;
}
PetscErrorCode DMStagGetProductCoordinateLocationSlot(DM dm, DMStagStencilLocation loc, PetscInt *slot)
{
  (void)dm;
  MockCheck(loc == DMSTAG_LEFT || loc == DMSTAG_ELEMENT || loc == DMSTAG_RIGHT, "DMStagGetProductCoordinateLocationSlot: 1-D coordinates have LEFT, ELEMENT and RIGHT only");
  *slot = loc == DMSTAG_LEFT ? 0 : (loc == DMSTAG_ELEMENT ? 1 : 2);
  return PETSC_SUCCESS;
}

/* ------------------------------------------------------------------ Mesh (cart) */
PetscErrorCode MockMeshCartCreate(PetscInt dim, const PetscInt N[], const PetscBool periodic[], const double *const xf[], Mesh *mesh)
{
  Mesh m = (Mesh)xcalloc(sizeof(*m));
  int  d, li;
  MockCheck(dim == 2 || dim == 3, "MockMeshCartCreate: dim 2 or 3");
  header(m, "Mesh", MESHCART);
  m->dim = (int)dim;
  for (d = 0; d < 3; ++d) m->N[d] = d < dim ? (int)N[d] : 1, m->per[d] = d < dim ? (int)periodic[d] : 0;
  /* cart.c:88-120: scalar (1 dof per element), vector (dim dof per element), staggered scalar (1 dof per face) */
  m->dm[MESH_DM_SCALAR]      = dim == 2 ? dm_new(m, 0, 0, 1, 0) : dm_new(m, 0, 0, 0, 1);
  m->dm[MESH_DM_VECTOR]      = dim == 2 ? dm_new(m, 0, 0, (int)dim, 0) : dm_new(m, 0, 0, 0, (int)dim);
  m->dm[MESH_DM_STAG_SCALAR] = dim == 2 ? dm_new(m, 0, 1, 0, 0) : dm_new(m, 0, 0, 1, 0);
  for (d = 0; d < dim; ++d) {
    const DM     dm = m->dm[0];
    const double L  = xf[d][m->N[d]] - xf[d][0];
    m->coord[d] = (double *)xcalloc(sizeof(double) * (2 * (size_t)dm->gn[d] + 1));
    m->ctab[d]  = (double **)xcalloc(sizeof(double *) * (size_t)dm->gn[d]);
    for (li = 0; li < dm->gn[d]; ++li) {
      const int g = li + dm->gs[d];
      double    left, right = NAN;
      if (g < 0) left = xf[d][g + m->N[d]] - L, right = xf[d][g + m->N[d] + 1] - L;
      else if (g < m->N[d]) left = xf[d][g], right = xf[d][g + 1];
      else {
        left = xf[d][m->N[d]];
        if (m->per[d]) right = xf[d][1] + L;
      }
      m->coord[d][2 * li]     = left;
      m->coord[d][2 * li + 1] = (left + right) / 2.; /* the centre of a partial element does not exist: NaN */
      m->ctab[d][li]          = m->coord[d] + 2 * li;
    }
    m->coord[d][2 * dm->gn[d]] = NAN;
  }
  *mesh = m;
  return PETSC_SUCCESS;
}
PetscErrorCode MeshDestroy(Mesh *pm)
{
  Mesh m = *pm;
  int  d;
  if (!m) return PETSC_SUCCESS;
  *pm = NULL;
  if (--m->hdr.refct > 0) return PETSC_SUCCESS;
  for (d = 0; d < 3; ++d) xfree(m->dm[d]), xfree(m->coord[d]), xfree(m->ctab[d]);
  xfree(m);
  return PETSC_SUCCESS;
}
PetscErrorCode MeshGetDimension(Mesh m, PetscInt *dim) { return *dim = m->dim, PETSC_SUCCESS; }
PetscErrorCode MeshGetDM(Mesh m, MeshDMType t, DM *dm)
{
  MockCheck(t == MESH_DM_SCALAR || t == MESH_DM_VECTOR || t == MESH_DM_STAG_SCALAR, "MeshGetDM: the mock has no staggered-vector DM");
  *dm = m->dm[t];
  return PETSC_SUCCESS;
}
PetscErrorCode MeshCreateGlobalVector(Mesh m, MeshDMType t, Vec *v)
{
  DM dm;
  PetscCall(MeshGetDM(m, t, &dm));
  *v = vec_new(dm, 0, global_size(dm));
  return PETSC_SUCCESS;
}
PetscErrorCode MeshGetNumberBoundaries(Mesh m, PetscInt *nb) { return *nb = 2 * m->dim, PETSC_SUCCESS; }
PetscErrorCode MeshCartGetGlobalSizes(Mesh m, PetscInt *M, PetscInt *N, PetscInt *P) { return *M = m->N[0], *N = m->N[1], *P = m->dim == 3 ? m->N[2] : GARBAGE, PETSC_SUCCESS; }
PetscErrorCode MeshCartGetNumRanks(Mesh m, PetscInt *x, PetscInt *y, PetscInt *z) { return *x = 1, *y = 1, *z = m->dim == 3 ? 1 : GARBAGE, PETSC_SUCCESS; }
PetscErrorCode MeshCartGetBoundaryTypes(Mesh m, MeshCartBoundaryType *x, MeshCartBoundaryType *y, MeshCartBoundaryType *z)
{
  *x = m->per[0] ? MESHCART_BOUNDARY_PERIODIC : MESHCART_BOUNDARY_NONE;
  *y = m->per[1] ? MESHCART_BOUNDARY_PERIODIC : MESHCART_BOUNDARY_NONE;
  if (m->dim == 3) *z = m->per[2] ? MESHCART_BOUNDARY_PERIODIC : MESHCART_BOUNDARY_NONE;
  return PETSC_SUCCESS;
}
PetscErrorCode MeshCartGetCoordinateArraysRead(Mesh m, const PetscScalar ***ax, const PetscScalar ***ay, const PetscScalar ***az) { return DMStagGetProductCoordinateArraysRead(m->dm[0], ax, ay, az); }
PetscErrorCode MeshCartRestoreCoordinateArraysRead(Mesh m, const PetscScalar ***ax, const PetscScalar ***ay, const PetscScalar ***az) { return DMStagRestoreProductCoordinateArraysRead(m->dm[0], ax, ay, az); }
PetscErrorCode MeshCartGetCoordinateLocationSlot(Mesh m, MeshCartCoordinateStencilLocation loc, PetscInt *slot) { return DMStagGetProductCoordinateLocationSlot(m->dm[0], loc == MESHCART_PREV ? DMSTAG_LEFT : DMSTAG_RIGHT, slot); }
PetscErrorCode MeshCartGetCorners(Mesh m, PetscInt *x, PetscInt *y, PetscInt *z, PetscInt *mm, PetscInt *n, PetscInt *p) { return DMStagGetCorners(m->dm[0], x, y, z, mm, n, p, NULL, NULL, NULL); }
PetscErrorCode MeshCartGetIsLastRank(Mesh m, PetscBool *x, PetscBool *y, PetscBool *z) { return *x = PETSC_TRUE, *y = PETSC_TRUE, *z = m->dim == 3 ? PETSC_TRUE : PETSC_FALSE, PETSC_SUCCESS; }
PetscErrorCode MeshCartGetIsFirstRank(Mesh m, PetscBool *x, PetscBool *y, PetscBool *z) { return *x = PETSC_TRUE, *y = PETSC_TRUE, *z = m->dim == 3 ? PETSC_TRUE : PETSC_FALSE, PETSC_SUCCESS; }

/* ------------------------------------------------------------------ viewers */
PetscErrorCode MockViewerASCIIOpen(FILE *f, PetscViewer *viewer)
{
  *viewer = (PetscViewer)xcalloc(sizeof(**viewer));
  header(*viewer, "PetscViewer", PETSCVIEWERASCII);
  (*viewer)->f = f;
  return PETSC_SUCCESS;
}
PetscErrorCode MockViewerStoreOpen(PetscViewer *viewer)
{
  *viewer = (PetscViewer)xcalloc(sizeof(**viewer));
  header(*viewer, "PetscViewer", "mockstore");
  return PETSC_SUCCESS;
}
PetscErrorCode PetscViewerDestroy(PetscViewer *pv)
{
  PetscViewer v = *pv;
  if (!v) return PETSC_SUCCESS;
  *pv = NULL;
  while (v->store) {
    struct stored *s = v->store;
    v->store         = s->next;
    xfree(s->a), xfree(s);
  }
  xfree(v);
  return PETSC_SUCCESS;
}
PetscErrorCode PetscViewerASCIIPrintf(PetscViewer v, const char fmt[], ...)
{
  va_list ap;
  MockCheck(v->f, "PetscViewerASCIIPrintf: not an ASCII viewer");
  va_start(ap, fmt);
  vfprintf(v->f, fmt, ap);
  va_end(ap);
  return PETSC_SUCCESS;
}
PetscErrorCode VecView(Vec x, PetscViewer v)
{
  if (v->f) {
    fprintf(v->f, "Vec Object: %s, %zu entries\n", x->hdr.name, x->n);
  } else {
    struct stored *s;
    MockCheck(x->hdr.name[0], "VecView into a store needs a named vector");
    for (s = v->store; s; s = s->next)
      if (!strcmp(s->name, x->hdr.name)) break;
    if (!s) {
      s = (struct stored *)xcalloc(sizeof(*s));
      snprintf(s->name, sizeof(s->name), "%s", x->hdr.name);
      s->n = x->n, s->a = (double *)xcalloc(sizeof(double) * x->n);
      s->next = v->store, v->store = s;
    }
    MockCheck(s->n == x->n, "VecView: size of %s changed", s->name);
    memcpy(s->a, x->a, sizeof(double) * x->n);
  }
  return PETSC_SUCCESS;
}
PetscErrorCode FlucaVecLoad(Vec x, PetscViewer v)
{
  struct stored *s;
  for (s = v->store; s; s = s->next)
    if (!strcmp(s->name, x->hdr.name)) break;
  MockCheck(s && s->n == x->n, "FlucaVecLoad: no stored vector named \"%s\" of this size", x->hdr.name);
  memcpy(x->a, s->a, sizeof(double) * x->n);
  ++x->hdr.state;
  return PETSC_SUCCESS;
}

/* ------------------------------------------------------------------ NS base class */
static struct {
  char name[32];
  PetscErrorCode (*create)(NS);
} ns_types[8];
static int ns_ntypes = 0;
static const char *const field_name[3] = {NS_FIELD_VELOCITY, NS_FIELD_FACE_NORMAL_VELOCITY, NS_FIELD_PRESSURE};
static const MeshDMType  field_dm[3]   = {MESH_DM_VECTOR, MESH_DM_STAG_SCALAR, MESH_DM_SCALAR};
static struct _p_IS      field_is[3]   = {{.field = 0}, {.field = 1}, {.field = 2}};
/* what struct _p_NS of the stub header does not carry */
static struct ns_extra {
  NS               ns;
  Mat              J;
  Vec              x, r;
} extras[8];
static struct ns_extra *extra_of(NS ns)
{
  int i;
  for (i = 0; i < 8; ++i)
    if (extras[i].ns == ns) return &extras[i];
  for (i = 0; i < 8; ++i)
    if (!extras[i].ns) return extras[i].ns = ns, &extras[i];
  abort();
}

PetscErrorCode NSRegister(const char name[], PetscErrorCode (*create)(NS))
{
  int i;
  for (i = 0; i < ns_ntypes; ++i)
    if (!strcmp(ns_types[i].name, name)) break;
  MockCheck(i < 8, "too many NS types");
  if (i == ns_ntypes) ++ns_ntypes;
  snprintf(ns_types[i].name, sizeof(ns_types[i].name), "%s", name);
  ns_types[i].create = create;
  return PETSC_SUCCESS;
}
PetscErrorCode NSCreate(MPI_Comm comm, NS *ns)
{
  *ns = (NS)xcalloc(sizeof(**ns));
  header(*ns, "NS", "");
  (*ns)->hdr.comm = comm;
  (*ns)->rho = 1., (*ns)->mu = 1., (*ns)->dt = 1.;
  return PETSC_SUCCESS;
}
PetscErrorCode NSSetType(NS ns, const char type[])
{
  int i;
  for (i = 0; i < ns_ntypes; ++i)
    if (!strcmp(ns_types[i].name, type)) break;
  MockCheck(i < ns_ntypes, "Unknown NS type: %s", type);
  if (ns->ops->destroy) PetscCall(ns->ops->destroy(ns)); /* nsbasic.c:55-79 */
  memset(ns->ops, 0, sizeof(ns->ops));
  ns->setupcalled = PETSC_FALSE;
  snprintf(ns->hdr.type_name, sizeof(ns->hdr.type_name), "%s", type);
  return ns_types[i].create(ns);
}
PetscErrorCode NSSetMesh(NS ns, Mesh mesh)
{
  PetscInt nb;
  PetscCall(MeshDestroy(&ns->mesh));
  PetscCall(PetscFree(ns->bcs));
  ns->mesh = mesh, ++mesh->hdr.refct;
  PetscCall(MeshGetNumberBoundaries(mesh, &nb));
  PetscCall(PetscCallocStub(sizeof(NSBoundaryCondition) * (size_t)nb, &ns->bcs));
  return PETSC_SUCCESS;
}
PetscErrorCode NSSetDensity(NS ns, PetscReal rho) { return ns->rho = rho, PETSC_SUCCESS; }
PetscErrorCode NSSetViscosity(NS ns, PetscReal mu) { return ns->mu = mu, PETSC_SUCCESS; }
PetscErrorCode NSSetTimeStepSize(NS ns, PetscReal dt) { return ns->dt = dt, PETSC_SUCCESS; }
PetscErrorCode NSSetBoundaryCondition(NS ns, PetscInt index, NSBoundaryCondition bc)
{
  MockCheck(ns->mesh && index >= 0 && index < 2 * ns->mesh->dim, "NSSetBoundaryCondition: bad boundary index");
  ns->bcs[index] = bc;
  return PETSC_SUCCESS;
}
PetscErrorCode NSSetFromOptions(NS ns)
{
  const char *type = option_find("-ns_type"); /* nsopts.c:179-181 */
  if (type) PetscCall(NSSetType(ns, type));
  if (ns->ops->setfromoptions) PetscCall(ns->ops->setfromoptions(ns, NULL));
  return PETSC_SUCCESS;
}
PetscErrorCode NSGetField(NS ns, const char name[], PetscInt *idx, MeshDMType *dmtype, IS *is)
{
  int f;
  for (f = 0; f < 3; ++f)
    if (!strcmp(field_name[f], name)) break;
  MockCheck(f < 3, "Field \"%s\" not found", name);
  (void)ns;
  if (idx) *idx = f;
  if (dmtype) *dmtype = field_dm[f];
  if (is) *is = &field_is[f];
  return PETSC_SUCCESS;
}
PetscErrorCode NSGetSolutionSubVector(NS ns, const char name[], Vec *sub)
{
  IS is;
  PetscCall(NSGetField(ns, name, NULL, NULL, &is));
  return VecGetSubVector(ns->sol, is, sub);
}
PetscErrorCode NSRestoreSolutionSubVector(NS ns, const char name[], Vec *sub)
{
  IS is;
  PetscCall(NSGetField(ns, name, NULL, NULL, &is));
  return VecRestoreSubVector(ns->sol, is, sub);
}
PetscErrorCode NSGetSolution(NS ns, Vec *sol) { return *sol = ns->sol, PETSC_SUCCESS; }
static PetscErrorCode solution_nest(NS ns, Vec *nest)
{
  Vec sub[3];
  int f;
  for (f = 0; f < 3; ++f) {
    PetscCall(MeshCreateGlobalVector(ns->mesh, field_dm[f], &sub[f]));
    PetscCall(PetscObjectSetName((PetscObject)sub[f], field_name[f]));
  }
  *nest = nest_new(3, sub);
  for (f = 0; f < 3; ++f) PetscCall(VecDestroy(&sub[f]));
  return PETSC_SUCCESS;
}
PetscErrorCode MockNSCreateVecs(NS ns, Vec *x, Vec *f)
{
  if (x) PetscCall(solution_nest(ns, x));
  if (f) PetscCall(solution_nest(ns, f));
  return PETSC_SUCCESS;
}
PetscErrorCode NSSetUp(NS ns)
{
  struct ns_extra *e = extra_of(ns);
  int              matrixfree = 0;
  if (ns->setupcalled) return PETSC_SUCCESS;
  MockCheck(ns->hdr.type_name[0], "NSSetUp: the mock has no default type");
  MockCheck(ns->mesh, "Mesh not set");
  PetscCall(solution_nest(ns, &ns->sol)); /* nsbasic.c:178-199 */
#ifdef FLUCA_NS_HAS_MATRIXFREE
  matrixfree = ns->matrixfree; /* glue/patches/0001: no J, no null space, no SNES for such a type */
#endif
  if (!matrixfree) {
    int f;
    e->J = (Mat)xcalloc(sizeof(*e->J));
    header(e->J, "Mat", "nest");
    e->J->nest = 1;
    MockCheck(ns->ops->formjacobian, "NSFormJacobian: no method for this type"); /* PetscUseTypeMethod, nsbasic.c:308 */
    PetscCall(ns->ops->formjacobian(ns, e->x, e->J, NS_INIT_JACOBIAN));           /* nsbasic.c:205: before ops->setup */
    /* MatCreateVecs(J) needs every block row and column to have a layout: the diagonal blocks must exist with the field sizes */
    MockCheck(e->J->assembled, "the Jacobian was not assembled by formjacobian(INIT)");
    for (f = 0; f < 3; ++f) {
      DM       dm;
      PetscInt n;
      PetscCall(MeshGetDM(ns->mesh, field_dm[f], &dm));
      PetscCall(DMStagGetEntries(dm, &n));
      MockCheck(e->J->block[f][f] && e->J->block[f][f]->rows == n && e->J->block[f][f]->cols == n, "MatCreateVecs: block (%d,%d) of J missing or of the wrong size", f, f);
    }
    PetscCall(MockNSCreateVecs(ns, &e->x, &e->r));
  }
  if (ns->ops->setup) PetscCall(ns->ops->setup(ns));
  ns->setupcalled = PETSC_TRUE;
  return PETSC_SUCCESS;
}
PetscErrorCode NSStep(NS ns)
{
  int matrixfree = 0;
#ifdef FLUCA_NS_HAS_MATRIXFREE
  matrixfree = ns->matrixfree;
#endif
  if (!matrixfree) { /* nsbasic.c:281-282: 7N doubles copied on the host every step */
    if (!ns->sol0) PetscCall(VecDuplicate(ns->sol, &ns->sol0));
    PetscCall(VecCopy(ns->sol, ns->sol0));
  }
  MockCheck(ns->ops->step, "NSStep: no method for this type");
  PetscCall(ns->ops->step(ns));
  if (ns->reason >= 0) ++ns->step, ns->t += ns->dt;
  return PETSC_SUCCESS;
}
PetscErrorCode NSFormFunction(NS ns, Vec x, Vec f)
{
  MockCheck(ns->ops->formfunction, "NSFormFunction: no method for this type");
  return ns->ops->formfunction(ns, x, f);
}
PetscErrorCode NSView(NS ns, PetscViewer viewer)
{
  PetscCall(PetscViewerASCIIPrintf(viewer, "NS Object: type %s, step %d, time %g\n", ns->hdr.type_name, (int)ns->step, ns->t));
  if (ns->ops->view) PetscCall(ns->ops->view(ns, viewer));
  return PETSC_SUCCESS;
}
PetscErrorCode NSViewSolution(NS ns, PetscViewer viewer)
{
  int f;
  for (f = 0; f < 3; ++f) { /* nssol.c:142-147: the fields of ns->sol are viewed BEFORE the type's hook runs */
    Vec sub;
    PetscCall(VecGetSubVector(ns->sol, &field_is[f], &sub));
    PetscCall(VecView(sub, viewer));
    PetscCall(VecRestoreSubVector(ns->sol, &field_is[f], &sub));
  }
  if (ns->ops->viewsolution) PetscCall(ns->ops->viewsolution(ns, viewer));
  viewer->step = ns->step, viewer->time = ns->t; /* the output sequence number of the mesh (cartcgns.c) */
  return PETSC_SUCCESS;
}
PetscErrorCode NSLoadSolution(NS ns, PetscViewer viewer)
{
  int f;
  MockCheck(ns->setupcalled, "This function must be called after NSSetUp()");
  for (f = 0; f < 3; ++f) {
    Vec sub;
    PetscCall(VecGetSubVector(ns->sol, &field_is[f], &sub));
    PetscCall(FlucaVecLoad(sub, viewer));
    PetscCall(VecRestoreSubVector(ns->sol, &field_is[f], &sub));
  }
  MockCheck(ns->ops->loadsolution, "NSLoadSolution: no method for this type");
  PetscCall(ns->ops->loadsolution(ns, viewer));
  ns->step = viewer->step, ns->t = viewer->time; /* nssol.c:200-202 */
  return PETSC_SUCCESS;
}
PetscErrorCode NSDestroy(NS *pns)
{
  NS               ns = *pns;
  struct ns_extra *e;
  if (!ns) return PETSC_SUCCESS;
  *pns = NULL;
  e    = extra_of(ns);
  if (ns->ops->destroy) PetscCall(ns->ops->destroy(ns));
  PetscCall(VecDestroy(&ns->sol));
  PetscCall(VecDestroy(&ns->sol0));
  PetscCall(VecDestroy(&e->x));
  PetscCall(VecDestroy(&e->r));
  PetscCall(MatDestroy(&e->J));
  e->ns = NULL;
  PetscCall(PetscFree(ns->bcs));
  PetscCall(MeshDestroy(&ns->mesh));
  xfree(ns);
  return PETSC_SUCCESS;
}
