/* oracle/ref_model/petsc_model_impl.h -- TEST INFRASTRUCTURE ONLY: the object layouts of the PETSc model (petsc_model.c) and the
 * helpers ref_driver.c uses to reach into them.  The reference's sources never see this header (they include petsc_model.h only). */
#ifndef PETSC_MODEL_IMPL_H
#define PETSC_MODEL_IMPL_H
#include <stdio.h>
#include <petsc_model.h>

#define MODEL_GARBAGE (-777) /* what arguments documented as "ignored in lower dimensions" come back as */

struct _p_ISLocalToGlobalMapping {
  int  n;
  int *idx; /* local entry -> global entry; -1 for entries that do not exist in a partial element */
};
struct _p_IS {
  struct _p_PetscObject hdr;
  int                   field; /* the field of the solution nest this index set selects */
  PetscInt              n;     /* its number of entries */
};
/* 1-D product coordinates shared by the DMs of a mesh: per direction [ghosted element][LEFT, ELEMENT]; RIGHT = next LEFT */
struct model_coords {
  int      refct, dim, gs[3], gn[3];
  double  *coord[3];
  double **ctab[3];
};
struct model_coords *ModelCoordsCreate(int dim, const int N[3], const int per[3]);
void                 ModelCoordsDestroy(struct model_coords *);

struct _p_Vec {
  struct _p_PetscObject hdr;
  PetscErrorCode (*view_op)(Vec, PetscViewer);
  PetscErrorCode (*load_op)(Vec, PetscViewer);
  DM                    dm;
  int                   local;
  PetscInt              n;
  double               *a;
  int                   nsub;
  Vec                   sub[3];
  void                 *table[3];
  int                   array_out;
};
struct _p_MatNullSpace {
  int       refct;
  PetscBool has_cnst;
  Vec       vec;
};
struct _p_Mat {
  struct _p_PetscObject  hdr;
  PetscInt               m, n;
  int                   *rn, *rcap, **rc;
  double               **rv;
  ISLocalToGlobalMapping rl2g, cl2g;
  int                    nest, assembled;
  Mat                    blk[3][3];
  IS                     isr[3], isc[3];
  MatNullSpace           nullspace;
};
struct _p_KSP {
  struct _p_PetscObject hdr;
  Mat                   A;
  long                  nsolves;
  PC                    pc; /* the KSP of a SNES owns a PC (SNESGetKSP / KSPGetPC) */
  double                rtol;
  int                   iterative, maxit; /* 0: dense LU ("exact"); 1: GMRES(30) + ILU(0) of petsc_model_ksp.c */
  int                   last_its, last_reason;
  long                  total_its;
  double                last_rel;
};
/* PETSC_MODEL_TIMING=1: cumulative wall time of the model's expensive entry points, printed by PetscFinalize / ModelTimingReport
   (a stand-in for -log_view: where a run of the reference on the model spends its time) */
enum { MT_KSPSOLVE, MT_MATMATMULT, MT_MATSETVALUES, MT_MATAXPY, MT_MATMULT, MT_MATDUP, MT_NSLOTS };
double ModelWallTime(void);
void   ModelTimingAdd(int slot, double seconds);
void   ModelTimingReport(FILE *f);
/* defaults of KSPs nobody configured: ModelKSPSetDefaults(1, ...) / -model_solvers iterative = what serial PETSc would run */
void ModelKSPSetDefaults(int iterative);
void ModelKSPSetDefaultRtol(double rtol); /* of iterative KSPs whose rtol nobody set (PETSc: 1e-5) */
int  ModelKSPGetDefaultIterative(void);
int  ModelKSPSolveIterative(Mat A, const double *b, double *x, double rtol, int maxit, int *its, double *rel);
struct _p_SNES {
  struct _p_PetscObject hdr;
  void                 *ctx;
  PetscErrorCode (*solve)(SNES, Vec, Vec); /* ref_driver.c attaches its own; NULL: the generic Picard solve of petsc_model_app.c */
  KSP ksp;
  Vec r;
  Mat J, Jpre;
  PetscErrorCode (*bfunc)(SNES, Vec, Vec, void *);
  PetscErrorCode (*jfunc)(SNES, Vec, Mat, Mat, void *);
  PetscErrorCode (*func)(SNES, Vec, Vec, void *);
  PetscErrorCode (*guess)(SNES, Vec, void *);
  void *pctx, *fctx, *gctx;
  int   mode, its, nhist; /* 0 exact, 1 one PC application (-ns_ksp_type preonly), 2 right-preconditioned GMRES */
  double hist[256];
};
struct stored_vec {
  char              *name;
  PetscInt           n;
  double            *a;
  struct stored_vec *next;
};
struct _p_PetscViewer {
  struct _p_PetscObject hdr;
  FILE                 *f;
  char                 *path;  /* "flucacgns" model viewers dump named vectors (natural ordering of their DM) to this file */
  struct stored_vec    *store;
  PetscInt              step;
  PetscReal             time;
};
#define MODEL_MAXLOC 8
struct _p_DM {
  struct _p_PetscObject            hdr;
  int                              dim, N[3], per[3], dof[4], epe, gs[3], gn[3], on[3], nloc;
  DMStagStencilLocation            loc[MODEL_MAXLOC];
  int                              locmask[MODEL_MAXLOC], locoff[MODEL_MAXLOC], locdof[MODEL_MAXLOC];
  int                              nglobal, nlocal;
  struct _p_ISLocalToGlobalMapping l2g;
  struct model_coords             *coords; /* shared by the DMs of a mesh */
  int                              setup, ncomposite;
  DM                               composite[4];
  PetscInt                         ownership[3];
};

PetscErrorCode ModelSNESSolvePicard(SNES, Vec, Vec);
const char *ModelLastError(void);
void        ModelHeaderInit(void *obj, PetscClassId, const char *cls, const char *type, PetscErrorCode (*destroy)(PetscObject));
void        ModelHeaderFree(void *obj);
Vec         ModelVecCreate(DM dm, int local, PetscInt n);
Vec         ModelVecCreateNest(int nsub, Vec sub[]);
void        ModelVecGather(Vec, double *);
void        ModelVecScatter(Vec, const double *);
Mat         ModelMatCreateAIJ(PetscInt m, PetscInt n);
double     *ModelMatEntry(Mat, int i, int j, int create);
double      ModelDenseSolve(int n, double *a, double *b);
DM          ModelDMStagCreate(int dim, const int N[3], const int per[3], int d0, int d1, int d2, int d3, struct model_coords *coords);
void        ModelDMDestroy(DM);
#endif
