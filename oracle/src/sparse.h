/*
 * oracle/src/sparse.h -- TEST INFRASTRUCTURE (CPU oracle), not product code.
 *
 * Minimal CSR sparse-matrix toolbox standing in for the PETSc calls the reference makes on
 * the NS hot path (PETSc is an un-vendored dependency of thecasterian/fluca, version >= 3.23,
 * see /root/reference/fluca/CMakeLists.txt:5-14):
 *   MatSetValues(ADD/INSERT)+MatAssembly -> coo_add / coo_to_csr
 *   MatMult / MatMultAdd                 -> csr_mult / csr_mult_add
 *   MatMatMult                           -> csr_matmat          (abfpc.c:153,170; cnlinearcart2d.c:2035)
 *   MatAXPY(DIFFERENT_NONZERO_PATTERN)   -> csr_axpy
 *   PCILU / PCBJACOBI defaults           -> bjilu0_*            (PETSc default inner PC)
 *   KSPGMRES(30)                         -> gmres               (PETSc default KSP, nssol.c:21-29)
 */
#pragma once
#include <stddef.h>

typedef struct {
  int     nrows, ncols;
  long    nnz;
  int    *ptr; /* nrows+1 */
  int    *idx; /* nnz, sorted within a row */
  double *val; /* nnz */
} Csr;

typedef struct {
  int     nrows, ncols;
  long    n, cap;
  int    *r, *c;
  double *v;
} Coo;

Coo *coo_new(int nrows, int ncols);
void coo_add(Coo *m, int r, int c, double v);
void coo_free(Coo *m);
/* duplicates are summed (ADD_VALUES); explicit zeros are kept, as PETSc keeps them */
Csr *coo_to_csr(const Coo *m);

Csr *csr_copy(const Csr *a);
void csr_free(Csr *a);
void csr_scale(Csr *a, double s);
void csr_mult(const Csr *a, const double *x, double *y);                  /* y = A x     */
void csr_mult_add(const Csr *a, const double *x, const double *w, double *y); /* y = w + A x */
Csr *csr_matmat(const Csr *a, const Csr *b);                               /* A * B       */
Csr *csr_axpy(const Csr *y, double alpha, const Csr *x);                   /* Y + alpha X, union pattern */
Csr *csr_shift_identity(const Csr *a, double s);                           /* A + s I     */

/* block-Jacobi ILU(0): nblocks contiguous row blocks, off-block couplings dropped
 * (nblocks == 1 is plain ILU(0), PETSc's serial default; nblocks == ranks is PETSc's
 * parallel default bjacobi+ilu). Zero pivots are shifted like PETSc's MAT_SHIFT_NONZERO. */
typedef struct {
  int     n, nblocks;
  int    *bstart; /* nblocks+1 */
  Csr    *lu;     /* factors stored in A's (block-restricted) pattern */
  int    *diag;   /* position of the diagonal in each row */
} BJIlu0;

BJIlu0 *bjilu0_factor(const Csr *a, int nblocks);
void    bjilu0_solve(const BJIlu0 *f, const double *b, double *x);
void    bjilu0_free(BJIlu0 *f);

/* generic operator: y = Op(ctx) x */
typedef void (*OpFn)(void *ctx, const double *x, double *y);

typedef struct {
  int    its;
  int    converged; /* 1 = rtol/atol reached, 0 = maxit */
  double rnorm0, rnorm;
  int    nhist;
  double hist[512]; /* residual norm history (norm type of the solve) */
} KspInfo;

/* Restarted GMRES, classical Gram-Schmidt with one refinement pass (PETSc default is CGS with
 * refinement if needed). side = 0: left preconditioning, preconditioned residual norm (PETSc
 * default for inner KSPs). side = 1: right preconditioning, true residual norm (what
 * KSP_NORM_UNPRECONDITIONED selects for the outer KSP, nssol.c:25). x holds the initial guess.
 * project (may be NULL) is applied to every preconditioned vector / solution update
 * (MatNullSpaceRemove semantics of PETSc KSP). */
void gmres(int n, OpFn A, void *actx, OpFn M, void *mctx, OpFn project, void *pctx, const double *b, double *x, int restart, double rtol, double atol, int maxit, int side, KspInfo *info);

double vec_dot(int n, const double *a, const double *b);
double vec_norm(int n, const double *a);
