/* oracle/ref_model/petsc_model_ksp.c -- TEST INFRASTRUCTURE ONLY (part of the single-rank PETSc model, see README.md).
 *
 * The iterative KSP of the model: what a serial PETSc run does for a KSP nobody configured -- GMRES(30), left-preconditioned
 * with ILU(0) of the operator (explicitly stored zeros belong to the pattern), zero initial guess, convergence on the
 * preconditioned residual norm (rtol 1e-5, abstol 1e-50, 10000 iterations), a constant null space attached to the operator taken out
 * of every preconditioned vector.  Written from the KSPGMRES / PCILU / KSPSolve manual pages; no PETSc source.  It exists so that
 * the reference's own sources can run at sizes where the dense "exact" KSP of petsc_model.c cannot (the timed CPU baseline of
 * bench.py, `-model_solvers iterative` for the reference's programs).  Its iterates are NOT PETSc's (orthogonalisation order,
 * pivot handling), so iteration counts are indicative, not pinned. */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "petsc_model_impl.h"

typedef struct {
  int     n;
  int    *ia, *ja, *dg; /* CSR with sorted columns; dg[i] = position of the diagonal entry of row i */
  double *a, *lu;       /* operator values; ILU(0) factors in the same pattern (unit lower, diagonal stored inverted) */
} Csr;

static void csr_free(Csr *c) { free(c->ia), free(c->ja), free(c->dg), free(c->a), free(c->lu); }

static int csr_build(Mat A, Csr *c)
{
  const int n = (int)A->m;
  long      nnz = 0;
  int       i, k;
  memset(c, 0, sizeof(*c));
  c->n  = n;
  c->ia = (int *)malloc(sizeof(int) * ((size_t)n + 1)), c->dg = (int *)malloc(sizeof(int) * (size_t)n);
  for (i = 0; i < n; ++i) nnz += A->rn[i];
  c->ja = (int *)malloc(sizeof(int) * (size_t)nnz), c->a = (double *)malloc(sizeof(double) * (size_t)nnz), c->lu = (double *)malloc(sizeof(double) * (size_t)nnz);
  c->ia[0] = 0;
  for (i = 0; i < n; ++i) {
    const int o = c->ia[i], m = A->rn[i];
    for (k = 0; k < m; ++k) { /* insertion sort by column: rows hold a stencil's worth of entries */
      int          q = o + k;
      const int    col = A->rc[i][k];
      const double val = A->rv[i][k];
      while (q > o && c->ja[q - 1] > col) c->ja[q] = c->ja[q - 1], c->a[q] = c->a[q - 1], --q;
      c->ja[q] = col, c->a[q] = val;
    }
    c->ia[i + 1] = o + m;
    c->dg[i]     = -1;
    for (k = o; k < o + m; ++k)
      if (c->ja[k] == i) c->dg[i] = k;
    if (c->dg[i] < 0) return 1; /* PCILU needs the diagonal in the pattern */
  }
  return 0;
}

/* ILU(0), row by row (IKJ); a pivot below 2.2e-14 of its row's largest entry is replaced (PCILU's default shift is "nonzero") */
static void ilu0(Csr *c)
{
  const int n   = c->n;
  int      *pos = (int *)malloc(sizeof(int) * (size_t)n), i, k, q;
  memcpy(c->lu, c->a, sizeof(double) * (size_t)c->ia[n]);
  for (i = 0; i < n; ++i) pos[i] = -1;
  for (i = 0; i < n; ++i) {
    double rowmax = 0., piv;
    for (k = c->ia[i]; k < c->ia[i + 1]; ++k) pos[c->ja[k]] = k, rowmax = fmax(rowmax, fabs(c->a[k]));
    for (k = c->ia[i]; k < c->dg[i]; ++k) {
      const int    r = c->ja[k];
      const double l = c->lu[k] * c->lu[c->dg[r]]; /* the diagonal of a finished row is stored inverted */
      c->lu[k] = l;
      if (l != 0.)
        for (q = c->dg[r] + 1; q < c->ia[r + 1]; ++q)
          if (pos[c->ja[q]] >= 0) c->lu[pos[c->ja[q]]] -= l * c->lu[q];
    }
    piv = c->lu[c->dg[i]];
    if (fabs(piv) < 2.2e-14 * rowmax || piv == 0.) piv = (piv < 0. ? -1. : 1.) * fmax(2.2e-14 * rowmax, 1e-300) * 100.;
    c->lu[c->dg[i]] = 1. / piv;
    for (k = c->ia[i]; k < c->ia[i + 1]; ++k) pos[c->ja[k]] = -1;
  }
  free(pos);
}
static void ilu_solve(const Csr *c, const double *b, double *x)
{
  const int n = c->n;
  int       i, k;
  for (i = 0; i < n; ++i) {
    double s = b[i];
    for (k = c->ia[i]; k < c->dg[i]; ++k) s -= c->lu[k] * x[c->ja[k]];
    x[i] = s;
  }
  for (i = n - 1; i >= 0; --i) {
    double s = x[i];
    for (k = c->dg[i] + 1; k < c->ia[i + 1]; ++k) s -= c->lu[k] * x[c->ja[k]];
    x[i] = s * c->lu[c->dg[i]];
  }
}
static void csr_mult(const Csr *c, const double *x, double *y)
{
  int i, k;
  for (i = 0; i < c->n; ++i) {
    double s = 0.;
    for (k = c->ia[i]; k < c->ia[i + 1]; ++k) s += c->a[k] * x[c->ja[k]];
    y[i] = s;
  }
}
static double dotn(int n, const double *x, const double *y)
{
  double s0 = 0., s1 = 0., s2 = 0., s3 = 0.; /* four partial sums: the compiler may not reassociate a single one */
  int    i;
  for (i = 0; i + 3 < n; i += 4) s0 += x[i] * y[i], s1 += x[i + 1] * y[i + 1], s2 += x[i + 2] * y[i + 2], s3 += x[i + 3] * y[i + 3];
  for (; i < n; ++i) s0 += x[i] * y[i];
  return (s0 + s1) + (s2 + s3);
}
static void remove_mean(int n, double *x)
{
  double s = 0.;
  int    i;
  for (i = 0; i < n; ++i) s += x[i];
  s /= n;
  for (i = 0; i < n; ++i) x[i] -= s;
}

/* x = approximate solution of A x = b; returns 0 (converged), 1 (iteration limit), 2 (no diagonal); *its = Krylov iterations */
int ModelKSPSolveIterative(Mat A, const double *b, double *x, double rtol, int maxit, int *its_out, double *rnorm_out)
{
  enum { M = 30 };
  const int n = (int)A->m, cnst = A->nullspace && A->nullspace->has_cnst;
  Csr       c;
  double   *V[M + 1], *w, H[M + 1][M], cs[M], sn[M], g[M + 1], y[M], rnorm, rnorm0 = -1.;
  int       i, j, k, its = 0, done = 0, rc = 0;
  if (csr_build(A, &c)) return csr_free(&c), 2;
  ilu0(&c);
  for (k = 0; k <= M; ++k) V[k] = (double *)malloc(sizeof(double) * (size_t)n);
  w = (double *)malloc(sizeof(double) * (size_t)n);
  memset(x, 0, sizeof(double) * (size_t)n);
  for (;;) {
    /* preconditioned residual of the current x (x = 0 on entry: M^-1 b) */
    if (its) {
      csr_mult(&c, x, w);
      for (i = 0; i < n; ++i) w[i] = b[i] - w[i];
      ilu_solve(&c, w, V[0]);
    } else ilu_solve(&c, b, V[0]);
    if (cnst) remove_mean(n, V[0]);
    rnorm = sqrt(dotn(n, V[0], V[0]));
    if (rnorm0 < 0.) rnorm0 = rnorm;
    if (done || rnorm <= fmax(rtol * rnorm0, 1e-50) || its >= maxit) break;
    for (i = 0; i < n; ++i) V[0][i] /= rnorm;
    memset(g, 0, sizeof(g));
    g[0] = rnorm;
    for (k = 0; k < M && its < maxit; ++k) {
      double *vn = V[k + 1];
      csr_mult(&c, V[k], w);
      ilu_solve(&c, w, vn);
      if (cnst) remove_mean(n, vn);
      for (j = 0; j <= k; ++j) {
        const double hjk = dotn(n, vn, V[j]);
        H[j][k] = hjk;
        for (i = 0; i < n; ++i) vn[i] -= hjk * V[j][i];
      }
      H[k + 1][k] = sqrt(dotn(n, vn, vn));
      if (H[k + 1][k] > 0.)
        for (i = 0; i < n; ++i) vn[i] /= H[k + 1][k];
      for (j = 0; j < k; ++j) {
        const double a = H[j][k], d = H[j + 1][k];
        H[j][k] = cs[j] * a + sn[j] * d, H[j + 1][k] = -sn[j] * a + cs[j] * d;
      }
      {
        const double a = H[k][k], d = H[k + 1][k], r = hypot(a, d);
        cs[k] = r > 0. ? a / r : 1., sn[k] = r > 0. ? d / r : 0.;
        H[k][k] = r, H[k + 1][k] = 0.;
        g[k + 1] = -sn[k] * g[k], g[k] = cs[k] * g[k];
      }
      ++its;
      rnorm = fabs(g[k + 1]);
      if (rnorm <= fmax(rtol * rnorm0, 1e-50)) {
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
    for (j = 0; j < k; ++j)
      for (i = 0; i < n; ++i) x[i] += y[j] * V[j][i];
    if (done) break; /* KSPGMRES trusts the recurrence for the converged norm */
  }
  if (!done && rnorm > fmax(rtol * rnorm0, 1e-50)) rc = 1;
  if (its_out) *its_out = its;
  if (rnorm_out) *rnorm_out = rnorm0 > 0. ? rnorm / rnorm0 : 0.;
  for (k = 0; k <= M; ++k) free(V[k]);
  free(w);
  csr_free(&c);
  return rc;
}
