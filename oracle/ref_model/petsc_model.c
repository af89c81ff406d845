/* oracle/ref_model/petsc_model.c -- TEST INFRASTRUCTURE ONLY: bodies of the single-rank PETSc model declared in
 * include/petsc_model.h (see there for what is modelled and what is not).  Written from the PETSc manual pages; no PETSc source. */
#include "petsc_model_impl.h"
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>
#include <time.h>

PetscClassId PC_CLASSID = 11, VEC_CLASSID = 12, MAT_CLASSID = 13;

static int    mt_on = -1;
static double mt_sum[MT_NSLOTS];
static long   mt_calls[MT_NSLOTS];
double ModelWallTime(void)
{
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}
void ModelTimingAdd(int slot, double seconds) { mt_sum[slot] += seconds, ++mt_calls[slot]; }
void ModelTimingReport(FILE *f)
{
  static const char *const names[MT_NSLOTS] = {"KSPSolve", "MatMatMult", "MatSetValues", "MatAXPY", "MatMult", "MatDuplicate"};
  int i;
  for (i = 0; i < MT_NSLOTS; ++i) fprintf(f, "[PETSc model timing] %-14s %10ld calls %10.3f s\n", names[i], mt_calls[i], mt_sum[i]);
}

/* ------------------------------------------------------------------ errors, memory, strings */
static char last_error[1024];
const char *ModelLastError(void) { return last_error; }
PetscErrorCode ModelError(MPI_Comm comm, int err, const char *file, int line, const char *fmt, ...)
{
  va_list ap;
  int     n;
  (void)comm;
  n = snprintf(last_error, sizeof(last_error), "[PETSc model] error %d at %s:%d: ", err, file, line);
  va_start(ap, fmt);
  vsnprintf(last_error + n, sizeof(last_error) - (size_t)n, fmt, ap);
  va_end(ap);
  fprintf(stderr, "%s\n", last_error);
  return err ? err : PETSC_ERR_LIB;
}
PetscErrorCode ModelErrorTrace(PetscErrorCode e, const char *file, int line, const char *fn)
{
  fprintf(stderr, "[PETSc model]   from %s() %s:%d\n", fn, file, line);
  return e;
}
PetscErrorCode ModelMalloc(size_t n, int zero, void *pp)
{
  void *p = malloc(n ? n : 1);
  if (!p) return ModelError(0, 55, __FILE__, __LINE__, "out of memory");
  memset(p, zero ? 0 : 0xA5, n ? n : 1); /* PetscMalloc1 does not clear */
  *(void **)pp = p;
  return PETSC_SUCCESS;
}
PetscErrorCode ModelFree(void *p)
{
  free(p);
  return PETSC_SUCCESS;
}
static void *zalloc(size_t n)
{
  void *p = calloc(1, n ? n : 1);
  if (!p) abort();
  return p;
}
PetscErrorCode PetscMemzero(void *p, size_t n) { return memset(p, 0, n), PETSC_SUCCESS; }
PetscErrorCode PetscSNPrintf(char *s, size_t n, const char fmt[], ...)
{
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(s, n, fmt, ap);
  va_end(ap);
  return PETSC_SUCCESS;
}
PetscErrorCode PetscStrcmp(const char a[], const char b[], PetscBool *e) { return *e = (a && b && !strcmp(a, b)) ? PETSC_TRUE : PETSC_FALSE, PETSC_SUCCESS; }
PetscErrorCode PetscStrallocpy(const char s[], char **t)
{
  *t = s ? strdup(s) : NULL;
  return PETSC_SUCCESS;
}
PetscErrorCode PetscInfoModel(void *o, const char *fmt, ...) { return (void)o, (void)fmt, PETSC_SUCCESS; }
PetscErrorCode PetscPrintf(MPI_Comm c, const char fmt[], ...)
{
  va_list ap;
  (void)c;
  va_start(ap, fmt);
  vfprintf(stdout, fmt, ap);
  va_end(ap);
  return PETSC_SUCCESS;
}
int MPI_Comm_rank(MPI_Comm c, int *r) { return (void)c, *r = 0, 0; }
int MPI_Comm_size(MPI_Comm c, int *s) { return (void)c, *s = 1, 0; }
int MPI_Bcast(void *b, int n, MPI_Datatype t, int root, MPI_Comm c) { return (void)b, (void)n, (void)t, (void)root, (void)c, 0; }
int MPI_Allreduce(const void *s, void *r, int n, MPI_Datatype t, MPI_Op op, MPI_Comm c)
{
  (void)op, (void)c;
  if (s != MPI_IN_PLACE) memcpy(r, s, (size_t)n * (t == MPI_DOUBLE ? sizeof(double) : t == MPI_INT ? sizeof(int) : 1));
  return 0;
}

/* ------------------------------------------------------------------ PetscObject */
struct composed_object {
  char                   *name;
  PetscObject             obj;
  struct composed_object *next;
};
struct composed_function {
  char *name;
  void (*f)(void);
  struct composed_function *next;
};
void ModelHeaderInit(void *o, PetscClassId classid, const char *cls, const char *type, PetscErrorCode (*destroy)(PetscObject))
{
  PetscObject h = (PetscObject)o;
  h->classid = classid, h->class_name = cls, h->type_name = type ? strdup(type) : NULL, h->refct = 1, h->destroy_model = destroy;
}
void ModelHeaderFree(void *o)
{
  PetscObject h = (PetscObject)o;
  while (h->olist) {
    struct composed_object *c = h->olist;
    h->olist                  = c->next;
    if (c->obj) PetscObjectDereference(c->obj);
    free(c->name), free(c);
  }
  while (h->flist) {
    struct composed_function *c = h->flist;
    h->flist                    = c->next;
    free(c->name), free(c);
  }
  free(h->type_name), free(h->name), free(h->prefix);
  h->type_name = h->name = h->prefix = NULL;
}
MPI_Comm       PetscObjectComm(PetscObject o) { return o->comm; }
PetscErrorCode PetscObjectGetComm(PetscObject o, MPI_Comm *c) { return *c = o->comm, PETSC_SUCCESS; }
PetscErrorCode PetscObjectTypeCompare(PetscObject o, const char t[], PetscBool *m) { return *m = (o && o->type_name && !strcmp(o->type_name, t)) ? PETSC_TRUE : PETSC_FALSE, PETSC_SUCCESS; }
PetscErrorCode PetscObjectStateGet(PetscObject o, PetscObjectState *s) { return *s = o->state, PETSC_SUCCESS; }
PetscErrorCode PetscObjectSetName(PetscObject o, const char n[])
{
  free(o->name);
  o->name = strdup(n);
  return PETSC_SUCCESS;
}
PetscErrorCode PetscObjectReference(PetscObject o)
{
  if (o) ++o->refct;
  return PETSC_SUCCESS;
}
PetscErrorCode PetscObjectCompose(PetscObject o, const char name[], PetscObject x)
{
  struct composed_object **pc = &o->olist, *c;
  for (; *pc; pc = &(*pc)->next)
    if (!strcmp((*pc)->name, name)) break;
  if (*pc) {
    c = *pc;
    if (c->obj) PetscObjectDereference(c->obj);
    if (!x) {
      *pc = c->next;
      free(c->name), free(c);
      return PETSC_SUCCESS;
    }
  } else {
    if (!x) return PETSC_SUCCESS;
    c       = (struct composed_object *)zalloc(sizeof(*c));
    c->name = strdup(name), c->next = o->olist, o->olist = c;
  }
  c->obj = x, ++x->refct;
  return PETSC_SUCCESS;
}
PetscErrorCode PetscObjectQuery(PetscObject o, const char name[], PetscObject *x)
{
  struct composed_object *c;
  *x = NULL;
  for (c = o->olist; c; c = c->next)
    if (!strcmp(c->name, name)) *x = c->obj;
  return PETSC_SUCCESS;
}
PetscErrorCode PetscObjectComposeFunctionModel(PetscObject o, const char name[], void (*f)(void))
{
  struct composed_function **pc = &o->flist, *c;
  for (; *pc; pc = &(*pc)->next)
    if (!strcmp((*pc)->name, name)) break;
  if (*pc) {
    c = *pc;
    if (!f) {
      *pc = c->next;
      free(c->name), free(c);
      return PETSC_SUCCESS;
    }
  } else {
    if (!f) return PETSC_SUCCESS;
    c       = (struct composed_function *)zalloc(sizeof(*c));
    c->name = strdup(name), c->next = o->flist, o->flist = c;
  }
  c->f = f;
  return PETSC_SUCCESS;
}
PetscErrorCode PetscObjectQueryFunctionModel(PetscObject o, const char name[], void (**f)(void))
{
  struct composed_function *c;
  *f = NULL;
  for (c = o->flist; c; c = c->next)
    if (!strcmp(c->name, name)) *f = c->f;
  return PETSC_SUCCESS;
}
PetscErrorCode PetscObjectIncrementTabLevel(PetscObject o, PetscObject p, PetscInt n) { return o->tablevel = (p ? p->tablevel : 0) + n, PETSC_SUCCESS; }
PetscErrorCode PetscObjectSetOptions(PetscObject o, void *opt) { return o->options = opt, PETSC_SUCCESS; }
PetscErrorCode PetscObjectGetOptionsPrefix(PetscObject o, const char *p[]) { return *p = o->prefix, PETSC_SUCCESS; }

/* ------------------------------------------------------------------ viewer */
PetscErrorCode PetscViewerASCIIPrintf(PetscViewer v, const char fmt[], ...)
{
  va_list ap;
  if (!v || !v->f) return PETSC_SUCCESS;
  va_start(ap, fmt);
  vfprintf(v->f, fmt, ap);
  va_end(ap);
  return PETSC_SUCCESS;
}
PetscErrorCode PetscViewerASCIIPushTab(PetscViewer v) { return (void)v, PETSC_SUCCESS; }
PetscErrorCode PetscViewerASCIIPopTab(PetscViewer v) { return (void)v, PETSC_SUCCESS; }

/* ------------------------------------------------------------------ Vec */
static PetscErrorCode vec_destroy_obj(PetscObject o)
{
  Vec v = (Vec)o;
  return VecDestroy(&v);
}
Vec ModelVecCreate(DM dm, int local, PetscInt n)
{
  Vec v = (Vec)zalloc(sizeof(*v));
  ModelHeaderInit(v, VEC_CLASSID, "Vec", "seq", vec_destroy_obj);
  v->dm = dm, v->local = local, v->n = n;
  v->a = (double *)zalloc(sizeof(double) * (size_t)n);
  return v;
}
Vec ModelVecCreateNest(int nsub, Vec sub[])
{
  Vec v = (Vec)zalloc(sizeof(*v));
  int i;
  ModelHeaderInit(v, VEC_CLASSID, "Vec", VECNEST, vec_destroy_obj);
  v->nsub = nsub;
  for (i = 0; i < nsub; ++i) v->sub[i] = sub[i], ++sub[i]->hdr.refct, v->n += sub[i]->n;
  return v;
}
PetscErrorCode VecDestroy(Vec *pv)
{
  Vec v = *pv;
  int i;
  if (!v) return PETSC_SUCCESS;
  *pv = NULL;
  if (--v->hdr.refct > 0) return PETSC_SUCCESS;
  for (i = 0; i < v->nsub; ++i) PetscCall(VecDestroy(&v->sub[i]));
  ModelHeaderFree(v);
  free(v->a), free(v);
  return PETSC_SUCCESS;
}
PetscErrorCode VecDuplicate(Vec x, Vec *y)
{
  if (x->nsub) {
    Vec sub[3];
    int s;
    for (s = 0; s < x->nsub; ++s) PetscCall(VecDuplicate(x->sub[s], &sub[s]));
    *y = ModelVecCreateNest(x->nsub, sub);
    for (s = 0; s < x->nsub; ++s) PetscCall(VecDestroy(&sub[s]));
  } else *y = ModelVecCreate(x->dm, x->local, x->n);
  return PETSC_SUCCESS;
}
#define SAME_LAYOUT(x, y) PetscCheck((x)->nsub == (y)->nsub && (x)->n == (y)->n, 0, PETSC_ERR_ARG_WRONG, "vector layouts differ (%d/%d entries, %d/%d blocks)", (int)(x)->n, (int)(y)->n, (x)->nsub, (y)->nsub)
PetscErrorCode VecSet(Vec v, PetscScalar a)
{
  int s, i;
  for (s = 0; s < v->nsub; ++s) PetscCall(VecSet(v->sub[s], a));
  if (!v->nsub)
    for (i = 0; i < v->n; ++i) v->a[i] = a;
  ++v->hdr.state;
  return PETSC_SUCCESS;
}
PetscErrorCode VecZeroEntries(Vec v) { return VecSet(v, 0.); }
PetscErrorCode VecCopy(Vec x, Vec y)
{
  int s;
  SAME_LAYOUT(x, y);
  if (x->nsub && getenv("PETSC_MODEL_TRACE")) fprintf(stderr, "[PETSc model trace] VecCopy of a nest (%d entries)\n", (int)x->n);
  for (s = 0; s < x->nsub; ++s) PetscCall(VecCopy(x->sub[s], y->sub[s]));
  if (!x->nsub && x != y) memcpy(y->a, x->a, sizeof(double) * (size_t)x->n);
  ++y->hdr.state;
  return PETSC_SUCCESS;
}
PetscErrorCode VecScale(Vec v, PetscScalar a)
{
  int s, i;
  for (s = 0; s < v->nsub; ++s) PetscCall(VecScale(v->sub[s], a));
  if (!v->nsub)
    for (i = 0; i < v->n; ++i) v->a[i] *= a;
  ++v->hdr.state;
  return PETSC_SUCCESS;
}
PetscErrorCode VecAXPY(Vec y, PetscScalar a, Vec x)
{
  int s, i;
  SAME_LAYOUT(x, y);
  for (s = 0; s < y->nsub; ++s) PetscCall(VecAXPY(y->sub[s], a, x->sub[s]));
  if (!y->nsub)
    for (i = 0; i < y->n; ++i) y->a[i] += a * x->a[i];
  ++y->hdr.state;
  return PETSC_SUCCESS;
}
PetscErrorCode VecAYPX(Vec y, PetscScalar a, Vec x)
{
  int s, i;
  SAME_LAYOUT(x, y);
  for (s = 0; s < y->nsub; ++s) PetscCall(VecAYPX(y->sub[s], a, x->sub[s]));
  if (!y->nsub)
    for (i = 0; i < y->n; ++i) y->a[i] = x->a[i] + a * y->a[i];
  ++y->hdr.state;
  return PETSC_SUCCESS;
}
PetscErrorCode VecWAXPY(Vec w, PetscScalar a, Vec x, Vec y)
{
  int s, i;
  SAME_LAYOUT(x, y);
  SAME_LAYOUT(x, w);
  for (s = 0; s < w->nsub; ++s) PetscCall(VecWAXPY(w->sub[s], a, x->sub[s], y->sub[s]));
  if (!w->nsub)
    for (i = 0; i < w->n; ++i) w->a[i] = a * x->a[i] + y->a[i];
  ++w->hdr.state;
  return PETSC_SUCCESS;
}
PetscErrorCode VecAXPBYPCZ(Vec z, PetscScalar a, PetscScalar b, PetscScalar c, Vec x, Vec y)
{
  int i;
  SAME_LAYOUT(x, y);
  SAME_LAYOUT(x, z);
  PetscCheck(!z->nsub, 0, PETSC_ERR_SUP, "VecAXPBYPCZ on a nest");
  for (i = 0; i < z->n; ++i) z->a[i] = a * x->a[i] + b * y->a[i] + (c == 0. ? 0. : c * z->a[i]); /* c = 0 must not propagate what z held */
  ++z->hdr.state;
  return PETSC_SUCCESS;
}
PetscErrorCode VecReciprocal(Vec v)
{
  int i;
  for (i = 0; i < v->n; ++i)
    if (v->a[i] != 0.) v->a[i] = 1. / v->a[i];
  ++v->hdr.state;
  return PETSC_SUCCESS;
}
PetscErrorCode VecPointwiseMult(Vec w, Vec x, Vec y)
{
  int i;
  SAME_LAYOUT(x, y);
  SAME_LAYOUT(x, w);
  for (i = 0; i < w->n; ++i) w->a[i] = x->a[i] * y->a[i];
  ++w->hdr.state;
  return PETSC_SUCCESS;
}
PetscErrorCode VecGetSize(Vec v, PetscInt *n) { return *n = v->n, PETSC_SUCCESS; }
PetscErrorCode VecGetSubVector(Vec v, IS is, Vec *sub)
{
  PetscCheck(v->nsub > 0 && is && is->field >= 0 && is->field < v->nsub, 0, PETSC_ERR_ARG_WRONG, "VecGetSubVector: the model only splits a nest by the index set of one of its fields");
  *sub = v->sub[is->field];
  return PETSC_SUCCESS;
}
PetscErrorCode VecRestoreSubVector(Vec v, IS is, Vec *sub)
{
  PetscCheck(v->nsub > 0 && is && *sub == v->sub[is->field], 0, PETSC_ERR_ARG_WRONG, "VecRestoreSubVector: not the vector handed out");
  ++v->hdr.state;
  *sub = NULL;
  return PETSC_SUCCESS;
}
PetscErrorCode VecAssemblyBegin(Vec v) { return (void)v, PETSC_SUCCESS; }
PetscErrorCode VecAssemblyEnd(Vec v) { return (void)v, PETSC_SUCCESS; }
PetscErrorCode VecView(Vec v, PetscViewer w) { return v->view_op ? v->view_op(v, w) : PETSC_SUCCESS; }
/* flat copies of a (possibly nested) vector, block after block */
void ModelVecGather(Vec v, double *out)
{
  int s, o = 0;
  if (!v->nsub) memcpy(out, v->a, sizeof(double) * (size_t)v->n);
  for (s = 0; s < v->nsub; ++s) memcpy(out + o, v->sub[s]->a, sizeof(double) * (size_t)v->sub[s]->n), o += v->sub[s]->n;
}
void ModelVecScatter(Vec v, const double *in)
{
  int s, o = 0;
  if (!v->nsub) memcpy(v->a, in, sizeof(double) * (size_t)v->n);
  for (s = 0; s < v->nsub; ++s) memcpy(v->sub[s]->a, in + o, sizeof(double) * (size_t)v->sub[s]->n), o += v->sub[s]->n;
  ++v->hdr.state;
}

/* ------------------------------------------------------------------ Mat */
static PetscErrorCode mat_destroy_obj(PetscObject o)
{
  Mat A = (Mat)o;
  return MatDestroy(&A);
}
static void mat_alloc_rows(Mat A)
{
  A->rn = (int *)zalloc(sizeof(int) * (size_t)A->m), A->rcap = (int *)zalloc(sizeof(int) * (size_t)A->m);
  A->rc = (int **)zalloc(sizeof(int *) * (size_t)A->m), A->rv = (double **)zalloc(sizeof(double *) * (size_t)A->m);
}
Mat ModelMatCreateAIJ(PetscInt m, PetscInt n)
{
  Mat A = (Mat)zalloc(sizeof(*A));
  ModelHeaderInit(A, MAT_CLASSID, "Mat", MATAIJ, mat_destroy_obj);
  A->m = m, A->n = n;
  mat_alloc_rows(A);
  return A;
}
double *ModelMatEntry(Mat A, int i, int j, int create)
{
  int k;
  for (k = 0; k < A->rn[i]; ++k)
    if (A->rc[i][k] == j) return &A->rv[i][k];
  if (!create) return NULL;
  if (A->rn[i] == A->rcap[i]) {
    A->rcap[i] = A->rcap[i] ? 2 * A->rcap[i] : 8;
    A->rc[i]   = (int *)realloc(A->rc[i], sizeof(int) * (size_t)A->rcap[i]);
    A->rv[i]   = (double *)realloc(A->rv[i], sizeof(double) * (size_t)A->rcap[i]);
  }
  A->rc[i][A->rn[i]] = j, A->rv[i][A->rn[i]] = 0.;
  return &A->rv[i][A->rn[i]++];
}
PetscErrorCode MatCreate(MPI_Comm c, Mat *A)
{
  *A = (Mat)zalloc(sizeof(**A));
  ModelHeaderInit(*A, MAT_CLASSID, "Mat", NULL, mat_destroy_obj);
  (*A)->hdr.comm = c;
  return PETSC_SUCCESS;
}
PetscErrorCode MatSetSizes(Mat A, PetscInt m, PetscInt n, PetscInt M, PetscInt N)
{
  (void)M, (void)N;
  PetscCheck(!A->rn, 0, PETSC_ERR_ARG_WRONGSTATE, "MatSetSizes after the rows exist");
  A->m = m, A->n = n;
  mat_alloc_rows(A);
  return PETSC_SUCCESS;
}
PetscErrorCode MatSetType(Mat A, MatType t)
{
  free(A->hdr.type_name);
  A->hdr.type_name = strdup(t);
  return PETSC_SUCCESS;
}
PetscErrorCode MatSetUp(Mat A) { return (void)A, PETSC_SUCCESS; }
PetscErrorCode MatSetLocalToGlobalMapping(Mat A, ISLocalToGlobalMapping r, ISLocalToGlobalMapping c) { return A->rl2g = r, A->cl2g = c, PETSC_SUCCESS; }
PetscErrorCode MatSetOption(Mat A, MatOption o, PetscBool b) { return (void)A, (void)o, (void)b, PETSC_SUCCESS; }
static PetscErrorCode MatSetValuesLocal_impl(Mat A, PetscInt nr, const PetscInt ir[], PetscInt nc, const PetscInt ic[], const PetscScalar v[], InsertMode mode)
{
  int r, c;
  PetscCheck(A->rl2g && A->cl2g, 0, PETSC_ERR_ARG_WRONGSTATE, "MatSetValuesLocal: no local-to-global mapping");
  PetscCheck(mode == INSERT_VALUES || mode == ADD_VALUES, 0, PETSC_ERR_ARG_WRONG, "MatSetValuesLocal: mode");
  for (r = 0; r < nr; ++r) {
    PetscCheck(ir[r] >= 0 && ir[r] < A->rl2g->n, 0, PETSC_ERR_ARG_OUTOFRANGE, "MatSetValuesLocal: local row %d out of range", (int)ir[r]);
    {
      const int gi = A->rl2g->idx[ir[r]];
      if (gi < 0) continue; /* negative indices are ignored */
      PetscCheck(gi < A->m, 0, PETSC_ERR_ARG_OUTOFRANGE, "MatSetValuesLocal: row %d of %d", gi, (int)A->m);
      for (c = 0; c < nc; ++c) {
        PetscCheck(ic[c] >= 0 && ic[c] < A->cl2g->n, 0, PETSC_ERR_ARG_OUTOFRANGE, "MatSetValuesLocal: local column %d out of range", (int)ic[c]);
        {
          const int gj = A->cl2g->idx[ic[c]];
          double   *e;
          if (gj < 0) continue;
          PetscCheck(gj < A->n, 0, PETSC_ERR_ARG_OUTOFRANGE, "MatSetValuesLocal: column %d of %d", gj, (int)A->n);
          e = ModelMatEntry(A, gi, gj, 1);
          if (mode == ADD_VALUES) *e += v[r * nc + c];
          else *e = v[r * nc + c];
        }
      }
    }
  }
  ++A->hdr.state;
  return PETSC_SUCCESS;
}
PetscErrorCode MatSetValuesLocal(Mat A, PetscInt nr, const PetscInt ir[], PetscInt nc, const PetscInt ic[], const PetscScalar v[], InsertMode mode)
{
  if (mt_on < 0) mt_on = getenv("PETSC_MODEL_TIMING") ? 1 : 0;
  if (!mt_on) return MatSetValuesLocal_impl(A, nr, ir, nc, ic, v, mode);
  {
    const double         t0 = ModelWallTime();
    const PetscErrorCode e  = MatSetValuesLocal_impl(A, nr, ir, nc, ic, v, mode);
    ModelTimingAdd(MT_MATSETVALUES, ModelWallTime() - t0);
    return e;
  }
}
PetscErrorCode MatAssemblyBegin(Mat A, MatAssemblyType t) { return (void)A, (void)t, PETSC_SUCCESS; }
PetscErrorCode MatAssemblyEnd(Mat A, MatAssemblyType t) { return (void)t, A->assembled = 1, PETSC_SUCCESS; }
PetscErrorCode MatDestroy(Mat *pA)
{
  Mat A = *pA;
  int i, j;
  if (!A) return PETSC_SUCCESS;
  *pA = NULL;
  if (--A->hdr.refct > 0) return PETSC_SUCCESS;
  for (i = 0; i < 3; ++i)
    for (j = 0; j < 3; ++j) PetscCall(MatDestroy(&A->blk[i][j]));
  for (i = 0; i < A->m && A->rc; ++i) free(A->rc[i]), free(A->rv[i]);
  PetscCall(MatNullSpaceDestroy(&A->nullspace));
  ModelHeaderFree(A);
  free(A->rn), free(A->rcap), free(A->rc), free(A->rv), free(A);
  return PETSC_SUCCESS;
}
#define PLAIN(A) PetscCheck(!(A)->nest, 0, PETSC_ERR_SUP, "%s on a MatNest", __func__)
PetscErrorCode MatZeroEntries(Mat A)
{
  int i, k;
  PLAIN(A);
  for (i = 0; i < A->m; ++i)
    for (k = 0; k < A->rn[i]; ++k) A->rv[i][k] = 0.; /* the nonzero pattern stays */
  return PETSC_SUCCESS;
}
PetscErrorCode MatScale(Mat A, PetscScalar a)
{
  int i, k;
  PLAIN(A);
  for (i = 0; i < A->m; ++i)
    for (k = 0; k < A->rn[i]; ++k) A->rv[i][k] *= a;
  return PETSC_SUCCESS;
}
PetscErrorCode MatShift(Mat A, PetscScalar a)
{
  int i;
  PLAIN(A);
  for (i = 0; i < A->m && i < A->n; ++i) *ModelMatEntry(A, i, i, 1) += a;
  return PETSC_SUCCESS;
}
static PetscErrorCode MatAXPY_impl(Mat Y, PetscScalar a, Mat X, MatStructure s)
{
  int i, k;
  (void)s;
  PLAIN(Y);
  PLAIN(X);
  PetscCheck(Y->m == X->m && Y->n == X->n, 0, PETSC_ERR_ARG_WRONG, "MatAXPY: sizes differ");
  for (i = 0; i < X->m; ++i)
    for (k = 0; k < X->rn[i]; ++k) *ModelMatEntry(Y, i, X->rc[i][k], 1) += a * X->rv[i][k];
  return PETSC_SUCCESS;
}
PetscErrorCode MatAXPY(Mat Y, PetscScalar a, Mat X, MatStructure s)
{
  if (mt_on < 0) mt_on = getenv("PETSC_MODEL_TIMING") ? 1 : 0;
  if (!mt_on) return MatAXPY_impl(Y, a, X, s);
  {
    const double         t0 = ModelWallTime();
    const PetscErrorCode e  = MatAXPY_impl(Y, a, X, s);
    ModelTimingAdd(MT_MATAXPY, ModelWallTime() - t0);
    return e;
  }
}
static PetscErrorCode mult_plain(Mat A, const double *x, double *y, int add)
{
  int i, k;
  for (i = 0; i < A->m; ++i) {
    double s = add ? y[i] : 0.;
    for (k = 0; k < A->rn[i]; ++k) s += A->rv[i][k] * x[A->rc[i][k]];
    y[i] = s;
  }
  return PETSC_SUCCESS;
}
static PetscErrorCode MatMult_impl(Mat A, Vec x, Vec y)
{
  if (A->nest) {
    int i, j;
    PetscCheck(x->nsub == 3 && y->nsub == 3 && x != y, 0, PETSC_ERR_ARG_WRONG, "MatMult(nest): needs two distinct nest vectors");
    for (i = 0; i < 3; ++i) {
      PetscCall(VecSet(y->sub[i], 0.));
      for (j = 0; j < 3; ++j)
        if (A->blk[i][j]) {
          PetscCheck(A->blk[i][j]->m == y->sub[i]->n && A->blk[i][j]->n == x->sub[j]->n, 0, PETSC_ERR_ARG_WRONG, "MatMult(nest): block (%d,%d) does not fit", i, j);
          PetscCall(mult_plain(A->blk[i][j], x->sub[j]->a, y->sub[i]->a, 1));
        }
    }
    return PETSC_SUCCESS;
  }
  PetscCheck(!x->nsub && !y->nsub && x->n == A->n && y->n == A->m && x != y, 0, PETSC_ERR_ARG_WRONG, "MatMult: %d x %d matrix, x has %d, y has %d entries", (int)A->m, (int)A->n, (int)x->n, (int)y->n);
  ++y->hdr.state;
  return mult_plain(A, x->a, y->a, 0);
}
PetscErrorCode MatMult(Mat A, Vec x, Vec y)
{
  if (mt_on < 0) mt_on = getenv("PETSC_MODEL_TIMING") ? 1 : 0;
  if (!mt_on) return MatMult_impl(A, x, y);
  {
    const double         t0 = ModelWallTime();
    const PetscErrorCode e  = MatMult_impl(A, x, y);
    ModelTimingAdd(MT_MATMULT, ModelWallTime() - t0);
    return e;
  }
}
PetscErrorCode MatMultAdd(Mat A, Vec x, Vec y, Vec z)
{
  PLAIN(A);
  PetscCheck(x->n == A->n && y->n == A->m && z->n == A->m && x != z, 0, PETSC_ERR_ARG_WRONG, "MatMultAdd: sizes");
  if (z != y) memcpy(z->a, y->a, sizeof(double) * (size_t)A->m);
  ++z->hdr.state;
  return mult_plain(A, x->a, z->a, 1);
}
static PetscErrorCode MatMatMult_impl(Mat A, Mat B, MatReuse r, PetscReal fill, Mat *C)
{
  int i, k, l;
  (void)fill;
  PLAIN(A);
  PLAIN(B);
  PetscCheck(r == MAT_INITIAL_MATRIX && A->n == B->m, 0, PETSC_ERR_ARG_WRONG, "MatMatMult: reuse / sizes");
  *C = ModelMatCreateAIJ(A->m, B->n);
  for (i = 0; i < A->m; ++i)
    for (k = 0; k < A->rn[i]; ++k) {
      const int    kk = A->rc[i][k];
      const double a  = A->rv[i][k];
      for (l = 0; l < B->rn[kk]; ++l) *ModelMatEntry(*C, i, B->rc[kk][l], 1) += a * B->rv[kk][l]; /* symbolic product: cancelled entries stay as explicit zeros */
    }
  (*C)->assembled = 1;
  return PETSC_SUCCESS;
}
PetscErrorCode MatMatMult(Mat A, Mat B, MatReuse r, PetscReal fill, Mat *C)
{
  if (mt_on < 0) mt_on = getenv("PETSC_MODEL_TIMING") ? 1 : 0;
  if (!mt_on) return MatMatMult_impl(A, B, r, fill, C);
  {
    const double         t0 = ModelWallTime();
    const PetscErrorCode e  = MatMatMult_impl(A, B, r, fill, C);
    ModelTimingAdd(MT_MATMATMULT, ModelWallTime() - t0);
    return e;
  }
}
static PetscErrorCode MatDuplicate_impl(Mat A, MatDuplicateOption o, Mat *B)
{
  int i, k;
  PLAIN(A);
  *B = ModelMatCreateAIJ(A->m, A->n);
  for (i = 0; i < A->m; ++i)
    for (k = 0; k < A->rn[i]; ++k) *ModelMatEntry(*B, i, A->rc[i][k], 1) = o == MAT_COPY_VALUES ? A->rv[i][k] : 0.;
  (*B)->rl2g = A->rl2g, (*B)->cl2g = A->cl2g, (*B)->assembled = 1;
  return PETSC_SUCCESS;
}
PetscErrorCode MatDuplicate(Mat A, MatDuplicateOption o, Mat *B)
{
  if (mt_on < 0) mt_on = getenv("PETSC_MODEL_TIMING") ? 1 : 0;
  if (!mt_on) return MatDuplicate_impl(A, o, B);
  {
    const double         t0 = ModelWallTime();
    const PetscErrorCode e  = MatDuplicate_impl(A, o, B);
    ModelTimingAdd(MT_MATDUP, ModelWallTime() - t0);
    return e;
  }
}
PetscErrorCode MatDiagonalScale(Mat A, Vec l, Vec r)
{
  int i, k;
  PLAIN(A);
  for (i = 0; i < A->m; ++i)
    for (k = 0; k < A->rn[i]; ++k) A->rv[i][k] *= (l ? l->a[i] : 1.) * (r ? r->a[A->rc[i][k]] : 1.);
  return PETSC_SUCCESS;
}
PetscErrorCode MatGetDiagonal(Mat A, Vec d)
{
  int i;
  PLAIN(A);
  PetscCheck(d->n == A->m, 0, PETSC_ERR_ARG_WRONG, "MatGetDiagonal: size");
  for (i = 0; i < A->m; ++i) {
    const double *e = ModelMatEntry(A, i, i, 0);
    d->a[i]         = e ? *e : 0.;
  }
  return PETSC_SUCCESS;
}
PetscErrorCode MatGetRowSum(Mat A, Vec d)
{
  int i, k;
  PLAIN(A);
  PetscCheck(d->n == A->m, 0, PETSC_ERR_ARG_WRONG, "MatGetRowSum: size");
  for (i = 0; i < A->m; ++i) {
    double s = 0.;
    for (k = 0; k < A->rn[i]; ++k) s += A->rv[i][k];
    d->a[i] = s;
  }
  return PETSC_SUCCESS;
}
PetscErrorCode MatCreateVecs(Mat A, Vec *right, Vec *left)
{
  if (A->nest) {
    int side, f;
    for (side = 0; side < 2; ++side) {
      Vec *out = side ? left : right, sub[3];
      if (!out) continue;
      for (f = 0; f < 3; ++f) sub[f] = ModelVecCreate(NULL, 0, side ? A->isr[f]->n : A->isc[f]->n); /* the layouts of a nest come from its index sets */
      *out = ModelVecCreateNest(3, sub);
      for (f = 0; f < 3; ++f) PetscCall(VecDestroy(&sub[f]));
    }
    return PETSC_SUCCESS;
  }
  if (right) *right = ModelVecCreate(NULL, 0, A->n);
  if (left) *left = ModelVecCreate(NULL, 0, A->m);
  return PETSC_SUCCESS;
}
PetscErrorCode MatCreateConstantDiagonal(MPI_Comm c, PetscInt m, PetscInt n, PetscInt M, PetscInt N, PetscScalar d, Mat *A)
{
  int i;
  (void)c, (void)M, (void)N;
  *A = ModelMatCreateAIJ(m, n);
  for (i = 0; i < m && i < n; ++i) *ModelMatEntry(*A, i, i, 1) = d;
  (*A)->assembled = 1;
  return PETSC_SUCCESS;
}
PetscErrorCode MatCreateNest(MPI_Comm c, PetscInt nr, const IS isr[], PetscInt nc, const IS isc[], const Mat a[], Mat *J)
{
  int i;
  PetscCheck(nr == 3 && nc == 3 && !a, c, PETSC_ERR_SUP, "the model has 3 x 3 nests created empty");
  if (getenv("PETSC_MODEL_TRACE")) fprintf(stderr, "[PETSc model trace] MatCreateNest\n");
  *J = (Mat)zalloc(sizeof(**J));
  ModelHeaderInit(*J, MAT_CLASSID, "Mat", MATNEST, mat_destroy_obj);
  (*J)->nest = 1;
  for (i = 0; i < 3; ++i) (*J)->isr[i] = isr[i], (*J)->isc[i] = isc[i];
  return PETSC_SUCCESS;
}
PetscErrorCode MatNestSetSubMat(Mat J, PetscInt i, PetscInt j, Mat B)
{
  PetscCheck(J->nest && i >= 0 && i < 3 && j >= 0 && j < 3, 0, PETSC_ERR_ARG_WRONG, "MatNestSetSubMat");
  PetscCall(MatDestroy(&J->blk[i][j]));
  J->blk[i][j] = B, ++B->hdr.refct;
  return PETSC_SUCCESS;
}
PetscErrorCode MatNestGetSubMat(Mat J, PetscInt i, PetscInt j, Mat *B)
{
  PetscCheck(J->nest && i >= 0 && i < 3 && j >= 0 && j < 3, 0, PETSC_ERR_ARG_WRONG, "MatNestGetSubMat");
  *B = J->blk[i][j]; /* borrowed */
  return PETSC_SUCCESS;
}
PetscErrorCode MatNestGetSize(Mat J, PetscInt *m, PetscInt *n) { return (void)J, *m = 3, *n = 3, PETSC_SUCCESS; }
PetscErrorCode MatNestGetISs(Mat J, IS r[], IS c[])
{
  int i;
  for (i = 0; i < 3; ++i) {
    if (r) r[i] = J->isr[i];
    if (c) c[i] = J->isc[i];
  }
  return PETSC_SUCCESS;
}
PetscErrorCode MatNestSetVecType(Mat J, VecType t) { return (void)J, (void)t, PETSC_SUCCESS; }
PetscErrorCode MatCreateSubMatrix(Mat J, IS r, IS c, MatReuse reuse, Mat *B)
{
  PetscCheck(J->nest && r && c && reuse == MAT_INITIAL_MATRIX, 0, PETSC_ERR_SUP, "MatCreateSubMatrix: the model extracts whole blocks of a nest by the index sets of its fields");
  *B = J->blk[r->field][c->field];
  if (*B) ++(*B)->hdr.refct; /* MatCreateSubMatrix_Nest hands out the block itself with a new reference */
  return PETSC_SUCCESS;
}
PetscErrorCode MatNullSpaceCreate(MPI_Comm c, PetscBool has_cnst, PetscInt n, const Vec v[], MatNullSpace *ns)
{
  (void)c;
  *ns = (MatNullSpace)zalloc(sizeof(**ns));
  (*ns)->refct = 1, (*ns)->has_cnst = has_cnst;
  PetscCheck(n <= 1, 0, PETSC_ERR_SUP, "the model holds at most one null vector");
  if (n == 1) (*ns)->vec = v[0], ++v[0]->hdr.refct;
  return PETSC_SUCCESS;
}
PetscErrorCode MatNullSpaceDestroy(MatNullSpace *pns)
{
  MatNullSpace ns = *pns;
  if (!ns) return PETSC_SUCCESS;
  *pns = NULL;
  if (--ns->refct > 0) return PETSC_SUCCESS;
  PetscCall(VecDestroy(&ns->vec));
  free(ns);
  return PETSC_SUCCESS;
}
/* x <- x - <x, n> n for the (normalised) null vector; a constant null space removes the mean */
PetscErrorCode MatNullSpaceRemove(MatNullSpace ns, Vec x)
{
  const int n = x->n;
  double   *a = (double *)malloc(sizeof(double) * (size_t)n), s = 0.;
  int       i;
  ModelVecGather(x, a);
  if (ns->has_cnst) {
    for (i = 0; i < n; ++i) s += a[i];
    for (i = 0; i < n; ++i) a[i] -= s / n;
  }
  if (ns->vec) {
    double *v = (double *)malloc(sizeof(double) * (size_t)n);
    PetscCheck(ns->vec->n == n, 0, PETSC_ERR_ARG_WRONG, "MatNullSpaceRemove: size");
    ModelVecGather(ns->vec, v);
    for (s = 0., i = 0; i < n; ++i) s += a[i] * v[i];
    for (i = 0; i < n; ++i) a[i] -= s * v[i];
    free(v);
  }
  ModelVecScatter(x, a);
  free(a);
  return PETSC_SUCCESS;
}
PetscErrorCode MatSetNullSpace(Mat A, MatNullSpace ns)
{
  if (ns) ++ns->refct;
  PetscCall(MatNullSpaceDestroy(&A->nullspace));
  A->nullspace = ns;
  return PETSC_SUCCESS;
}
PetscErrorCode MatGetNullSpace(Mat A, MatNullSpace *ns) { return *ns = A->nullspace, PETSC_SUCCESS; }

/* ------------------------------------------------------------------ dense LU: the "exact KSP" */
/* solves the n x n system a x = b in place (row-major a is destroyed); returns the smallest pivot relative to the largest */
double ModelDenseSolve(int n, double *a, double *b)
{
  int    i, j, k;
  double pmin = 1e300, pmax = 0.;
  for (k = 0; k < n; ++k) {
    int    p = k;
    double m = fabs(a[(size_t)k * n + k]);
    for (i = k + 1; i < n; ++i)
      if (fabs(a[(size_t)i * n + k]) > m) m = fabs(a[(size_t)i * n + k]), p = i;
    if (m < pmin) pmin = m;
    if (m > pmax) pmax = m;
    if (m == 0.) return 0.;
    if (p != k) {
      for (j = k; j < n; ++j) {
        const double t       = a[(size_t)k * n + j];
        a[(size_t)k * n + j] = a[(size_t)p * n + j], a[(size_t)p * n + j] = t;
      }
      {
        const double t = b[k];
        b[k] = b[p], b[p] = t;
      }
    }
    for (i = k + 1; i < n; ++i) {
      const double f = a[(size_t)i * n + k] / a[(size_t)k * n + k];
      if (f == 0.) continue;
      for (j = k + 1; j < n; ++j) a[(size_t)i * n + j] -= f * a[(size_t)k * n + j];
      b[i] -= f * b[k];
    }
  }
  for (k = n - 1; k >= 0; --k) {
    double s = b[k];
    for (j = k + 1; j < n; ++j) s -= a[(size_t)k * n + j] * b[j];
    b[k] = s / a[(size_t)k * n + k];
  }
  return pmin / pmax;
}
static int    ksp_default_iterative = 0;
static double ksp_default_rtol      = 1e-5;
void          ModelKSPSetDefaults(int iterative) { ksp_default_iterative = iterative; }
void          ModelKSPSetDefaultRtol(double rtol) { ksp_default_rtol = rtol > 0. ? rtol : 1e-5; }
int        ModelKSPGetDefaultIterative(void) { return ksp_default_iterative; }
PetscErrorCode KSPCreate(MPI_Comm c, KSP *k)
{
  *k = (KSP)zalloc(sizeof(**k));
  ModelHeaderInit(*k, 14, "KSP", "exact", NULL);
  (*k)->hdr.comm = c, (*k)->iterative = -1; /* -1: the default at the time of the solve */
  return PETSC_SUCCESS;
}
PetscErrorCode KSPSetOperators(KSP k, Mat A, Mat P)
{
  (void)P;
  if (A) ++A->hdr.refct;
  PetscCall(MatDestroy(&k->A));
  k->A = A;
  return PETSC_SUCCESS;
}
static PetscErrorCode KSPSolve_impl(KSP k, Vec b, Vec x)
{
  Mat       A = k->A;
  const int n = A ? A->m : 0, bordered = A && A->nullspace ? 1 : 0, N = n + bordered;
  double   *a, *r, piv;
  int       i, q;
  PetscCheck(A && !A->nest && A->m == A->n && b->n == n && x->n == n, 0, PETSC_ERR_ARG_WRONG, "KSPSolve: needs a square AIJ operator and matching vectors");
  if (k->iterative > 0 || (k->iterative < 0 && ksp_default_iterative)) { /* an unconverged solve is not an error (KSP_DIVERGED_ITS) */
    const int rc = ModelKSPSolveIterative(A, b->a, x->a, k->rtol > 0. ? k->rtol : ksp_default_rtol, k->maxit > 0 ? k->maxit : 10000, &k->last_its, &k->last_rel);
    PetscCheck(rc != 2, 0, PETSC_ERR_ARG_WRONGSTATE, "KSPSolve (ILU): the operator has a row without a diagonal entry");
    k->last_reason = rc ? -3 : 2, k->total_its += k->last_its;
    ++x->hdr.state, ++k->nsolves;
    if (getenv("PETSC_MODEL_KSP_MONITOR")) fprintf(stderr, "[PETSc model] KSP %s n=%d: %d iterations, relative preconditioned residual %.3e%s\n", k->hdr.prefix ? k->hdr.prefix : "", n, k->last_its, k->last_rel, rc ? " (iteration limit)" : "");
    return PETSC_SUCCESS;
  }
  PetscCheck(!bordered || (A->nullspace->has_cnst && !A->nullspace->vec), 0, PETSC_ERR_SUP, "KSPSolve: only the constant null space");
  a = (double *)zalloc(sizeof(double) * (size_t)N * N), r = (double *)zalloc(sizeof(double) * (size_t)N);
  for (i = 0; i < n; ++i) {
    for (q = 0; q < A->rn[i]; ++q) a[(size_t)i * N + A->rc[i][q]] += A->rv[i][q];
    r[i] = b->a[i];
    if (bordered) a[(size_t)i * N + n] = 1., a[(size_t)n * N + i] = 1.; /* zero-mean solution, the constant taken out of the residual */
  }
  piv = ModelDenseSolve(N, a, r);
  free(a);
  if (!(piv > 1e-14)) {
    free(r);
    SETERRQ(0, PETSC_ERR_LIB, "KSPSolve: operator singular to working precision (pivot ratio %g)", piv);
  }
  memcpy(x->a, r, sizeof(double) * (size_t)n);
  free(r);
  ++x->hdr.state, ++k->nsolves;
  return PETSC_SUCCESS;
}
PetscErrorCode KSPSolve(KSP k, Vec b, Vec x)
{
  if (mt_on < 0) mt_on = getenv("PETSC_MODEL_TIMING") ? 1 : 0;
  if (!mt_on) return KSPSolve_impl(k, b, x);
  {
    const double         t0 = ModelWallTime();
    const PetscErrorCode e  = KSPSolve_impl(k, b, x);
    ModelTimingAdd(MT_KSPSOLVE, ModelWallTime() - t0);
    return e;
  }
}
PetscErrorCode KSPDestroy(KSP *pk)
{
  KSP k = *pk;
  if (!k) return PETSC_SUCCESS;
  *pk = NULL;
  if (--k->hdr.refct > 0) return PETSC_SUCCESS;
  PetscCall(MatDestroy(&k->A));
  ModelHeaderFree(k);
  free(k);
  return PETSC_SUCCESS;
}
/* -<prefix>ksp_type exact | gmres (GMRES(30) + ILU(0)), -<prefix>ksp_rtol, -<prefix>ksp_max_it; the prefix is the KSP's own, whatever
   object's options are being processed around the call (PCSetFromOptions_ABF calls this for its sub-KSPs, abfpc.c:248-249) */
PetscErrorCode KSPSetFromOptions(KSP k)
{
  const char *outer = ModelOptionsPrefixGet();
  char        type[64] = "";
  PetscBool   set = PETSC_FALSE;
  PetscInt    maxit = k->maxit;
  ModelOptionsPrefixPush(k->hdr.prefix);
  PetscCall(ModelOptionsString("-ksp_type", type, sizeof(type), &set));
  if (set) {
    PetscCheck(!strcmp(type, "exact") || !strcmp(type, "gmres"), 0, PETSC_ERR_SUP, "the model's KSP types are exact and gmres, not %s", type);
    k->iterative = !strcmp(type, "gmres");
  }
  PetscCall(ModelOptionsReal("-ksp_rtol", &k->rtol, NULL));
  PetscCall(ModelOptionsInt("-ksp_max_it", &maxit, NULL));
  k->maxit = (int)maxit;
  ModelOptionsPrefixPush(outer);
  return PETSC_SUCCESS;
}
PetscErrorCode KSPSetOptionsPrefix(KSP k, const char p[])
{
  free(k->hdr.prefix);
  k->hdr.prefix = p ? strdup(p) : NULL;
  return PETSC_SUCCESS;
}
PetscErrorCode KSPView(KSP k, PetscViewer v) { return (void)k, (void)v, PETSC_SUCCESS; }
PetscErrorCode SNESSolve(SNES s, Vec b, Vec x)
{
  PetscCheck(s, 0, PETSC_ERR_ARG_WRONGSTATE, "SNESSolve: no SNES");
  return s->solve ? s->solve(s, b, x) : ModelSNESSolvePicard(s, b, x);
}

/* ------------------------------------------------------------------ DMStag */
static const struct {
  DMStagStencilLocation loc;
  int                   mask, stratum;
} loc2d[4] = {{DMSTAG_DOWN_LEFT, 3, 0}, {DMSTAG_DOWN, 2, 1}, {DMSTAG_LEFT, 1, 1}, {DMSTAG_ELEMENT, 0, 2}},
  loc3d[8] = {{DMSTAG_BACK_DOWN_LEFT, 7, 0}, {DMSTAG_BACK_DOWN, 6, 1}, {DMSTAG_BACK_LEFT, 5, 1}, {DMSTAG_BACK, 4, 2}, {DMSTAG_DOWN_LEFT, 3, 1}, {DMSTAG_DOWN, 2, 2}, {DMSTAG_LEFT, 1, 2}, {DMSTAG_ELEMENT, 0, 3}};

static int entry_exists(DM dm, int l, const int g[3])
{
  int d;
  for (d = 0; d < dm->dim; ++d)
    if (!dm->per[d] && g[d] == dm->N[d] && !(dm->locmask[l] >> d & 1)) return 0;
  return 1;
}
/* dof0..dof3: vertices, edges, faces, elements (2-D: vertices, faces, elements); coordinates belong to the mesh */
struct model_coords *ModelCoordsCreate(int dim, const int N[3], const int per[3])
{
  struct model_coords *c = (struct model_coords *)zalloc(sizeof(*c));
  int                  d, li;
  c->refct = 1, c->dim = dim;
  for (d = 0; d < dim; ++d) {
    c->gs[d]    = per[d] ? -1 : 0;
    c->gn[d]    = N[d] + (per[d] ? 2 : 1);
    c->coord[d] = (double *)zalloc(sizeof(double) * (2 * (size_t)c->gn[d] + 1));
    c->ctab[d]  = (double **)zalloc(sizeof(double *) * (size_t)c->gn[d]);
    for (li = 0; li < c->gn[d]; ++li) c->ctab[d][li] = c->coord[d] + 2 * li;
    for (li = 0; li <= 2 * c->gn[d]; ++li) c->coord[d][li] = NAN;
  }
  return c;
}
void ModelCoordsDestroy(struct model_coords *c)
{
  int d;
  if (!c || --c->refct > 0) return;
  for (d = 0; d < 3; ++d) free(c->coord[d]), free(c->ctab[d]);
  free(c);
}
DM ModelDMStagCreate(int dim, const int N[3], const int per[3], int d0, int d1, int d2, int d3, struct model_coords *coords)
{
  DM  dm = (DM)zalloc(sizeof(*dm));
  int d, l, li[3], g[3], w[3];
  ModelHeaderInit(dm, 15, "DM", "stag", NULL);
  dm->dim = dim, dm->coords = coords, dm->setup = 1;
  if (coords) ++coords->refct;
  dm->dof[0] = d0, dm->dof[1] = d1, dm->dof[2] = d2, dm->dof[3] = d3;
  for (d = 0; d < 3; ++d) {
    dm->N[d]   = d < dim ? N[d] : 1;
    dm->per[d] = d < dim ? per[d] : 1;
    dm->gs[d]  = d < dim && dm->per[d] ? -1 : 0;
    dm->gn[d]  = d < dim ? dm->N[d] + (dm->per[d] ? 2 : 1) : 1;
    dm->on[d]  = d < dim ? dm->N[d] + (dm->per[d] ? 0 : 1) : 1;
  }
  dm->nloc = dim == 2 ? 4 : 8;
  for (l = 0; l < dm->nloc; ++l) {
    dm->loc[l]     = dim == 2 ? loc2d[l].loc : loc3d[l].loc;
    dm->locmask[l] = dim == 2 ? loc2d[l].mask : loc3d[l].mask;
    dm->locdof[l]  = dm->dof[dim == 2 ? loc2d[l].stratum : loc3d[l].stratum];
    dm->locoff[l]  = dm->epe;
    dm->epe += dm->locdof[l];
  }
  /* global numbering: the existing entries of the owned elements in element order; local-to-global map of the ghosted box */
  {
    const size_t npad = (size_t)dm->on[0] * dm->on[1] * dm->on[2] * dm->epe;
    int         *gnum = (int *)zalloc(sizeof(int) * npad), c, cnt = 0;
    for (g[2] = 0; g[2] < dm->on[2]; ++g[2])
      for (g[1] = 0; g[1] < dm->on[1]; ++g[1])
        for (g[0] = 0; g[0] < dm->on[0]; ++g[0])
          for (l = 0; l < dm->nloc; ++l)
            for (c = 0; c < dm->locdof[l]; ++c) gnum[(((size_t)g[2] * dm->on[1] + g[1]) * dm->on[0] + g[0]) * dm->epe + dm->locoff[l] + c] = entry_exists(dm, l, g) ? cnt++ : -1;
    dm->nglobal  = cnt;
    dm->nlocal   = dm->gn[0] * dm->gn[1] * dm->gn[2] * dm->epe;
    dm->l2g.n    = dm->nlocal;
    dm->l2g.idx  = (int *)zalloc(sizeof(int) * (size_t)dm->nlocal);
    for (li[2] = 0; li[2] < dm->gn[2]; ++li[2])
      for (li[1] = 0; li[1] < dm->gn[1]; ++li[1])
        for (li[0] = 0; li[0] < dm->gn[0]; ++li[0]) {
          for (d = 0; d < 3; ++d) {
            g[d] = li[d] + dm->gs[d], w[d] = g[d];
            if (d < dim && dm->per[d]) w[d] = (g[d] + dm->N[d]) % dm->N[d];
          }
          for (l = 0; l < dm->nloc; ++l)
            for (c = 0; c < dm->locdof[l]; ++c)
              dm->l2g.idx[(((size_t)li[2] * dm->gn[1] + li[1]) * dm->gn[0] + li[0]) * dm->epe + dm->locoff[l] + c] = gnum[(((size_t)w[2] * dm->on[1] + w[1]) * dm->on[0] + w[0]) * dm->epe + dm->locoff[l] + c];
        }
    free(gnum);
  }
  return dm;
}
void ModelDMDestroy(DM dm)
{
  if (!dm) return;
  ModelHeaderFree(dm);
  ModelCoordsDestroy(dm->coords);
  free(dm->l2g.idx), free(dm);
}
PetscErrorCode DMGetDimension(DM dm, PetscInt *dim) { return *dim = dm->dim, PETSC_SUCCESS; }
PetscErrorCode DMGetLocalVector(DM dm, Vec *l) { return *l = ModelVecCreate(dm, 1, dm->nlocal), PETSC_SUCCESS; }
PetscErrorCode DMRestoreLocalVector(DM dm, Vec *l)
{
  PetscCheck(*l && (*l)->dm == dm && (*l)->local, 0, PETSC_ERR_ARG_WRONG, "DMRestoreLocalVector: not a local vector of this DM");
  return VecDestroy(l);
}
PetscErrorCode DMCreateGlobalVector(DM dm, Vec *g) { return *g = ModelVecCreate(dm, 0, dm->nglobal), PETSC_SUCCESS; }
PetscErrorCode DMGetGlobalVector(DM dm, Vec *g)
{
  int i;
  *g = ModelVecCreate(dm, 0, dm->nglobal);
  for (i = 0; i < dm->nglobal; ++i) (*g)->a[i] = NAN; /* contents of a checked-out work vector are undefined */
  return PETSC_SUCCESS;
}
PetscErrorCode DMRestoreGlobalVector(DM dm, Vec *g)
{
  PetscCheck(*g && (*g)->dm == dm && !(*g)->local, 0, PETSC_ERR_ARG_WRONG, "DMRestoreGlobalVector: not a global vector of this DM");
  return VecDestroy(g);
}
PetscErrorCode DMGlobalToLocal(DM dm, Vec g, InsertMode mode, Vec l)
{
  int i;
  PetscCheck(g && l && l->dm == dm && l->local && !g->local && g->n == dm->nglobal && mode == INSERT_VALUES, 0, PETSC_ERR_ARG_WRONG, "DMGlobalToLocal: vectors do not fit this DM (global has %d entries, DM %d)", g ? (int)g->n : -1, dm->nglobal);
  for (i = 0; i < dm->nlocal; ++i) l->a[i] = dm->l2g.idx[i] >= 0 ? g->a[dm->l2g.idx[i]] : NAN;
  ++l->hdr.state;
  return PETSC_SUCCESS;
}
PetscErrorCode DMLocalToGlobal(DM dm, Vec l, InsertMode mode, Vec g)
{
  int li[3], d, ghost, e;
  PetscCheck(g && l && l->dm == dm && l->local && !g->local && g->n == dm->nglobal, 0, PETSC_ERR_ARG_WRONG, "DMLocalToGlobal: vectors do not fit this DM");
  for (li[2] = 0; li[2] < dm->gn[2]; ++li[2])
    for (li[1] = 0; li[1] < dm->gn[1]; ++li[1])
      for (li[0] = 0; li[0] < dm->gn[0]; ++li[0]) {
        const size_t base = (((size_t)li[2] * dm->gn[1] + li[1]) * dm->gn[0] + li[0]) * dm->epe;
        for (ghost = 0, d = 0; d < dm->dim; ++d)
          if (dm->per[d] && (li[d] + dm->gs[d] < 0 || li[d] + dm->gs[d] >= dm->N[d])) ghost = 1;
        if (ghost && mode == INSERT_VALUES) continue;
        for (e = 0; e < dm->epe; ++e) {
          const int gi = dm->l2g.idx[base + e];
          if (gi < 0) continue;
          if (mode == ADD_VALUES) g->a[gi] += l->a[base + e];
          else g->a[gi] = l->a[base + e];
        }
      }
  ++g->hdr.state;
  return PETSC_SUCCESS;
}
PetscErrorCode DMGetLocalToGlobalMapping(DM dm, ISLocalToGlobalMapping *m) { return *m = &dm->l2g, PETSC_SUCCESS; }
PetscErrorCode DMGetMatType(DM dm, MatType *t) { return (void)dm, *t = MATAIJ, PETSC_SUCCESS; }
PetscErrorCode DMStagGetGlobalSizes(DM dm, PetscInt *M, PetscInt *N, PetscInt *P)
{
  if (M) *M = dm->N[0];
  if (N) *N = dm->N[1];
  if (P) *P = dm->dim == 3 ? dm->N[2] : MODEL_GARBAGE;
  return PETSC_SUCCESS;
}
PetscErrorCode DMStagGetCorners(DM dm, PetscInt *x, PetscInt *y, PetscInt *z, PetscInt *m, PetscInt *n, PetscInt *p, PetscInt *ex, PetscInt *ey, PetscInt *ez)
{
  if (x) *x = 0;
  if (y) *y = 0;
  if (z) *z = dm->dim == 3 ? 0 : MODEL_GARBAGE;
  if (m) *m = dm->N[0];
  if (n) *n = dm->N[1];
  if (p) *p = dm->dim == 3 ? dm->N[2] : MODEL_GARBAGE;
  if (ex) *ex = dm->per[0] ? 0 : 1;
  if (ey) *ey = dm->per[1] ? 0 : 1;
  if (ez) *ez = dm->dim == 3 ? (dm->per[2] ? 0 : 1) : MODEL_GARBAGE;
  return PETSC_SUCCESS;
}
PetscErrorCode DMStagGetIsFirstRank(DM dm, PetscBool *x, PetscBool *y, PetscBool *z)
{
  if (x) *x = PETSC_TRUE;
  if (y) *y = PETSC_TRUE;
  if (z) *z = dm->dim == 3 ? PETSC_TRUE : PETSC_FALSE;
  return PETSC_SUCCESS;
}
PetscErrorCode DMStagGetIsLastRank(DM dm, PetscBool *x, PetscBool *y, PetscBool *z) { return DMStagGetIsFirstRank(dm, x, y, z); }
PetscErrorCode DMStagGetEntries(DM dm, PetscInt *n) { return *n = dm->nglobal, PETSC_SUCCESS; }
static int find_loc(DM dm, DMStagStencilLocation loc)
{
  int l;
  for (l = 0; l < dm->nloc; ++l)
    if (dm->loc[l] == loc) return l;
  return -1;
}
PetscErrorCode DMStagGetLocationSlot(DM dm, DMStagStencilLocation loc, PetscInt c, PetscInt *slot)
{
  const int l = find_loc(dm, loc);
  PetscCheck(l >= 0, 0, PETSC_ERR_ARG_OUTOFRANGE, "DMStagGetLocationSlot: location %d is not stored by an element of a %d-D DMStag", (int)loc, dm->dim);
  PetscCheck(c >= 0 && c < dm->locdof[l], 0, PETSC_ERR_ARG_OUTOFRANGE, "DMStagGetLocationSlot: component %d of %d", (int)c, dm->locdof[l]);
  *slot = dm->locoff[l] + c;
  return PETSC_SUCCESS;
}
PetscErrorCode DMStagStencilToIndexLocal(DM dm, PetscInt dim, PetscInt n, const DMStagStencil *pos, PetscInt *ix)
{
  int q;
  PetscCheck(dim == dm->dim, 0, PETSC_ERR_ARG_WRONG, "DMStagStencilToIndexLocal: dimension");
  for (q = 0; q < n; ++q) {
    int                   e[3] = {pos[q].i, pos[q].j, dm->dim == 3 ? pos[q].k : 0}, l, d;
    DMStagStencilLocation loc  = pos[q].loc;
    /* a location on the upper side of an element is the lower-side location of the next element */
    if (loc == DMSTAG_RIGHT) loc = DMSTAG_LEFT, ++e[0];
    else if (loc == DMSTAG_UP) loc = DMSTAG_DOWN, ++e[1];
    else if (loc == DMSTAG_FRONT) loc = DMSTAG_BACK, ++e[2];
    l = find_loc(dm, loc);
    PetscCheck(l >= 0, 0, PETSC_ERR_SUP, "DMStagStencilToIndexLocal: location %d not modelled", (int)pos[q].loc);
    PetscCheck(pos[q].c >= 0 && pos[q].c < dm->locdof[l], 0, PETSC_ERR_ARG_OUTOFRANGE, "DMStagStencilToIndexLocal: component %d but the stratum has %d dof", (int)pos[q].c, dm->locdof[l]);
    for (d = 0; d < dm->dim; ++d) PetscCheck(e[d] - dm->gs[d] >= 0 && e[d] - dm->gs[d] < dm->gn[d], 0, PETSC_ERR_ARG_OUTOFRANGE, "DMStagStencilToIndexLocal: element (%d,%d,%d) is outside the ghosted region of this rank", e[0], e[1], e[2]);
    ix[q] = (int)(((((size_t)(e[2] - dm->gs[2]) * dm->gn[1] + (e[1] - dm->gs[1])) * dm->gn[0] + (e[0] - dm->gs[0])) * dm->epe) + dm->locoff[l] + pos[q].c);
  }
  return PETSC_SUCCESS;
}
PetscErrorCode DMStagVecSetValuesStencil(DM dm, Vec v, PetscInt n, const DMStagStencil *pos, const PetscScalar *val, InsertMode mode)
{
  int q;
  PetscCheck(v && !v->local && v->n == dm->nglobal, 0, PETSC_ERR_ARG_WRONG, "DMStagVecSetValuesStencil: needs a GLOBAL vector of this DM (%d entries, got %d)", dm->nglobal, v ? (int)v->n : -1);
  for (q = 0; q < n; ++q) {
    PetscInt ix;
    int      g;
    PetscCall(DMStagStencilToIndexLocal(dm, dm->dim, 1, &pos[q], &ix));
    g = dm->l2g.idx[ix];
    if (g < 0) continue;
    if (mode == ADD_VALUES) v->a[g] += val[q];
    else v->a[g] = val[q];
  }
  ++v->hdr.state;
  return PETSC_SUCCESS;
}
PetscErrorCode DMStagMatSetValuesStencil(DM dm, Mat A, PetscInt nr, const DMStagStencil *pr, PetscInt nc, const DMStagStencil *pc, const PetscScalar *val, InsertMode mode)
{
  PetscInt ir[64], ic[64];
  PetscCheck(nr <= 64 && nc <= 64, 0, PETSC_ERR_SUP, "DMStagMatSetValuesStencil: more than 64 points");
  PetscCall(DMStagStencilToIndexLocal(dm, dm->dim, nr, pr, ir));
  PetscCall(DMStagStencilToIndexLocal(dm, dm->dim, nc, pc, ic));
  if (!A->rl2g) A->rl2g = A->cl2g = &dm->l2g; /* a matrix made by DMCreateMatrix carries the DM's map */
  return MatSetValuesLocal(A, nr, ir, nc, ic, val, mode);
}
static PetscErrorCode stag_array(DM dm, Vec v, void *out)
{
  const size_t n0 = (size_t)dm->gn[0], n1 = (size_t)dm->gn[1], n2 = (size_t)dm->gn[2];
  size_t       i, j, k;
  PetscCheck(v && v->dm == dm && v->local, 0, PETSC_ERR_ARG_WRONG, "DMStagVecGetArray: needs a LOCAL vector of this DM");
  PetscCheck(!v->array_out, 0, PETSC_ERR_ARG_WRONGSTATE, "DMStagVecGetArray: array already checked out");
  {
    double **cells = (double **)zalloc(sizeof(double *) * n0 * n1 * n2);
    for (i = 0; i < n0 * n1 * n2; ++i) cells[i] = v->a + i * dm->epe;
    v->table[0] = cells;
    if (dm->dim == 2) {
      double ***rows = (double ***)zalloc(sizeof(double **) * n1);
      for (j = 0; j < n1; ++j) rows[j] = cells + j * n0 - dm->gs[0];
      v->table[1]       = rows;
      *(double ****)out = rows - dm->gs[1];
    } else {
      double  ***rows   = (double ***)zalloc(sizeof(double **) * n1 * n2);
      double ****planes = (double ****)zalloc(sizeof(double ***) * n2);
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
  PetscCheck(v && v->dm == dm && v->array_out, 0, PETSC_ERR_ARG_WRONGSTATE, "DMStagVecRestoreArray: no array checked out");
  for (t = 0; t < 3; ++t) free(v->table[t]), v->table[t] = NULL;
  v->array_out  = 0;
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
  PetscCheck(dm->coords, 0, PETSC_ERR_ARG_WRONGSTATE, "the DM has no coordinates");
  if (ax) *(double ***)ax = dm->coords->ctab[0] - dm->gs[0];
  if (ay) *(double ***)ay = dm->coords->ctab[1] - dm->gs[1];
  if (az && dm->dim == 3) *(double ***)az = dm->coords->ctab[2] - dm->gs[2];
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
  PetscCheck(loc == DMSTAG_LEFT || loc == DMSTAG_ELEMENT || loc == DMSTAG_RIGHT, 0, PETSC_ERR_ARG_OUTOFRANGE, "1-D product coordinates have LEFT, ELEMENT and RIGHT");
  *slot = loc == DMSTAG_LEFT ? 0 : (loc == DMSTAG_ELEMENT ? 1 : 2);
  return PETSC_SUCCESS;
}
