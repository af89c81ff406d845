/*
 * oracle/src/sparse.c -- TEST INFRASTRUCTURE (CPU oracle), not product code.
 * See sparse.h for which PETSc call each routine stands in for.
 */
#include "sparse.h"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static void *xmalloc(size_t n)
{
  void *p = malloc(n ? n : 1);
  if (!p) {
    fprintf(stderr, "oracle: out of memory (%zu bytes)\n", n);
    abort();
  }
  return p;
}

Coo *coo_new(int nrows, int ncols)
{
  Coo *m   = (Coo *)xmalloc(sizeof(Coo));
  m->nrows = nrows;
  m->ncols = ncols;
  m->n     = 0;
  m->cap   = 1024;
  m->r     = (int *)xmalloc(sizeof(int) * m->cap);
  m->c     = (int *)xmalloc(sizeof(int) * m->cap);
  m->v     = (double *)xmalloc(sizeof(double) * m->cap);
  return m;
}

void coo_add(Coo *m, int r, int c, double v)
{
  if (r < 0 || r >= m->nrows || c < 0 || c >= m->ncols) {
    fprintf(stderr, "oracle: coo_add index out of range (%d,%d) in %dx%d\n", r, c, m->nrows, m->ncols);
    abort();
  }
  if (m->n == m->cap) {
    m->cap *= 2;
    m->r = (int *)realloc(m->r, sizeof(int) * m->cap);
    m->c = (int *)realloc(m->c, sizeof(int) * m->cap);
    m->v = (double *)realloc(m->v, sizeof(double) * m->cap);
    if (!m->r || !m->c || !m->v) abort();
  }
  m->r[m->n] = r;
  m->c[m->n] = c;
  m->v[m->n] = v;
  m->n++;
}

void coo_free(Coo *m)
{
  if (!m) return;
  free(m->r);
  free(m->c);
  free(m->v);
  free(m);
}

typedef struct {
  int    c;
  double v;
} ColVal;

static int cmp_colval(const void *a, const void *b)
{
  int ca = ((const ColVal *)a)->c, cb = ((const ColVal *)b)->c;
  return (ca > cb) - (ca < cb);
}

Csr *coo_to_csr(const Coo *m)
{
  Csr  *a   = (Csr *)xmalloc(sizeof(Csr));
  int  *cnt = (int *)calloc((size_t)m->nrows + 1, sizeof(int));
  long  k;
  int   i;
  a->nrows = m->nrows;
  a->ncols = m->ncols;
  for (k = 0; k < m->n; ++k) cnt[m->r[k] + 1]++;
  for (i = 0; i < m->nrows; ++i) cnt[i + 1] += cnt[i];
  ColVal *tmp  = (ColVal *)xmalloc(sizeof(ColVal) * (size_t)(m->n ? m->n : 1));
  int    *fill = (int *)xmalloc(sizeof(int) * (size_t)(m->nrows + 1));
  memcpy(fill, cnt, sizeof(int) * (size_t)(m->nrows + 1));
  for (k = 0; k < m->n; ++k) {
    int p    = fill[m->r[k]]++;
    tmp[p].c = m->c[k];
    tmp[p].v = m->v[k];
  }
  a->ptr = (int *)xmalloc(sizeof(int) * (size_t)(m->nrows + 1));
  a->idx = (int *)xmalloc(sizeof(int) * (size_t)(m->n ? m->n : 1));
  a->val = (double *)xmalloc(sizeof(double) * (size_t)(m->n ? m->n : 1));
  long nnz = 0;
  for (i = 0; i < m->nrows; ++i) {
    int s = cnt[i], e = cnt[i + 1], p;
    a->ptr[i] = (int)nnz;
    /* stable insertion order inside equal columns does not matter: values are summed */
    qsort(tmp + s, (size_t)(e - s), sizeof(ColVal), cmp_colval);
    for (p = s; p < e; ++p) {
      if (p > s && tmp[p].c == tmp[p - 1].c) a->val[nnz - 1] += tmp[p].v;
      else {
        a->idx[nnz] = tmp[p].c;
        a->val[nnz] = tmp[p].v;
        nnz++;
      }
    }
  }
  a->ptr[m->nrows] = (int)nnz;
  a->nnz           = nnz;
  free(tmp);
  free(fill);
  free(cnt);
  return a;
}

Csr *csr_copy(const Csr *a)
{
  Csr *b   = (Csr *)xmalloc(sizeof(Csr));
  b->nrows = a->nrows;
  b->ncols = a->ncols;
  b->nnz   = a->nnz;
  b->ptr   = (int *)xmalloc(sizeof(int) * (size_t)(a->nrows + 1));
  b->idx   = (int *)xmalloc(sizeof(int) * (size_t)(a->nnz ? a->nnz : 1));
  b->val   = (double *)xmalloc(sizeof(double) * (size_t)(a->nnz ? a->nnz : 1));
  memcpy(b->ptr, a->ptr, sizeof(int) * (size_t)(a->nrows + 1));
  memcpy(b->idx, a->idx, sizeof(int) * (size_t)a->nnz);
  memcpy(b->val, a->val, sizeof(double) * (size_t)a->nnz);
  return b;
}

void csr_free(Csr *a)
{
  if (!a) return;
  free(a->ptr);
  free(a->idx);
  free(a->val);
  free(a);
}

void csr_scale(Csr *a, double s)
{
  long k;
  for (k = 0; k < a->nnz; ++k) a->val[k] *= s;
}

void csr_mult(const Csr *a, const double *x, double *y)
{
  int i;
#pragma omp parallel for schedule(static)
  for (i = 0; i < a->nrows; ++i) {
    double s = 0.;
    int    p;
    for (p = a->ptr[i]; p < a->ptr[i + 1]; ++p) s += a->val[p] * x[a->idx[p]];
    y[i] = s;
  }
}

void csr_mult_add(const Csr *a, const double *x, const double *w, double *y)
{
  int i;
#pragma omp parallel for schedule(static)
  for (i = 0; i < a->nrows; ++i) {
    double s = 0.;
    int    p;
    for (p = a->ptr[i]; p < a->ptr[i + 1]; ++p) s += a->val[p] * x[a->idx[p]];
    y[i] = w[i] + s;
  }
}

Csr *csr_matmat(const Csr *a, const Csr *b)
{
  /* Gustavson row-by-row product; the symbolic pattern keeps entries that cancel to zero,
   * as PETSc's MatMatMult does (this is why the reference's S carries explicit zeros). */
  Coo    *c    = coo_new(a->nrows, b->ncols);
  double *acc  = (double *)calloc((size_t)b->ncols, sizeof(double));
  int    *mark = (int *)xmalloc(sizeof(int) * (size_t)b->ncols);
  int    *list = (int *)xmalloc(sizeof(int) * (size_t)b->ncols);
  int     i, j;
  for (j = 0; j < b->ncols; ++j) mark[j] = -1;
  for (i = 0; i < a->nrows; ++i) {
    int nl = 0, p, q;
    for (p = a->ptr[i]; p < a->ptr[i + 1]; ++p) {
      int    k  = a->idx[p];
      double av = a->val[p];
      for (q = b->ptr[k]; q < b->ptr[k + 1]; ++q) {
        int col = b->idx[q];
        if (mark[col] != i) {
          mark[col]  = i;
          list[nl++] = col;
          acc[col]   = 0.;
        }
        acc[col] += av * b->val[q];
      }
    }
    for (p = 0; p < nl; ++p) coo_add(c, i, list[p], acc[list[p]]);
  }
  Csr *r = coo_to_csr(c);
  coo_free(c);
  free(acc);
  free(mark);
  free(list);
  return r;
}

Csr *csr_axpy(const Csr *y, double alpha, const Csr *x)
{
  Coo *c = coo_new(y->nrows, y->ncols);
  int  i, p;
  for (i = 0; i < y->nrows; ++i) {
    for (p = y->ptr[i]; p < y->ptr[i + 1]; ++p) coo_add(c, i, y->idx[p], y->val[p]);
    for (p = x->ptr[i]; p < x->ptr[i + 1]; ++p) coo_add(c, i, x->idx[p], alpha * x->val[p]);
  }
  Csr *r = coo_to_csr(c);
  coo_free(c);
  return r;
}

Csr *csr_shift_identity(const Csr *a, double s)
{
  Coo *c = coo_new(a->nrows, a->ncols);
  int  i, p;
  for (i = 0; i < a->nrows; ++i) {
    for (p = a->ptr[i]; p < a->ptr[i + 1]; ++p) coo_add(c, i, a->idx[p], a->val[p]);
    coo_add(c, i, i, s);
  }
  Csr *r = coo_to_csr(c);
  coo_free(c);
  return r;
}

/* ---------------------------------------------------------------- ILU(0) per block */

BJIlu0 *bjilu0_factor(const Csr *a, int nblocks)
{
  int n = a->nrows, b, i;
  if (nblocks < 1) nblocks = 1;
  if (nblocks > n) nblocks = n;
  BJIlu0 *f  = (BJIlu0 *)xmalloc(sizeof(BJIlu0));
  f->n       = n;
  f->nblocks = nblocks;
  f->bstart  = (int *)xmalloc(sizeof(int) * (size_t)(nblocks + 1));
  for (b = 0; b <= nblocks; ++b) f->bstart[b] = (int)((long)n * b / nblocks);

  /* block-restricted copy of A */
  Coo *c = coo_new(n, n);
  for (b = 0; b < nblocks; ++b) {
    int lo = f->bstart[b], hi = f->bstart[b + 1], p;
    for (i = lo; i < hi; ++i)
      for (p = a->ptr[i]; p < a->ptr[i + 1]; ++p)
        if (a->idx[p] >= lo && a->idx[p] < hi) coo_add(c, i, a->idx[p], a->val[p]);
  }
  f->lu = coo_to_csr(c);
  coo_free(c);
  f->diag = (int *)xmalloc(sizeof(int) * (size_t)n);
  Csr *lu = f->lu;
  for (i = 0; i < n; ++i) {
    int p;
    f->diag[i] = -1;
    for (p = lu->ptr[i]; p < lu->ptr[i + 1]; ++p)
      if (lu->idx[p] == i) f->diag[i] = p;
    if (f->diag[i] < 0) {
      fprintf(stderr, "oracle: ILU(0) needs a stored diagonal (row %d)\n", i);
      abort();
    }
  }
#pragma omp parallel for schedule(static, 1)
  for (b = 0; b < nblocks; ++b) {
    int  lo = f->bstart[b], hi = f->bstart[b + 1], ii;
    int *pos = (int *)xmalloc(sizeof(int) * (size_t)n);
    for (ii = 0; ii < n; ++ii) pos[ii] = -1;
    for (ii = lo; ii < hi; ++ii) {
      int p, q;
      for (p = lu->ptr[ii]; p < lu->ptr[ii + 1]; ++p) pos[lu->idx[p]] = p;
      for (p = lu->ptr[ii]; p < lu->ptr[ii + 1] && lu->idx[p] < ii; ++p) {
        int    k   = lu->idx[p];
        double lik = lu->val[p] / lu->val[f->diag[k]];
        lu->val[p] = lik;
        for (q = f->diag[k] + 1; q < lu->ptr[k + 1]; ++q) {
          int pp = pos[lu->idx[q]];
          if (pp >= 0) lu->val[pp] -= lik * lu->val[q];
        }
      }
      /* PETSc PCILU default: MAT_SHIFT_NONZERO with zeropivot/shift 100*eps */
      {
        double *d = &lu->val[f->diag[ii]];
        if (fabs(*d) < 100. * 2.220446049250313e-16) *d = (*d < 0. ? -1. : 1.) * 100. * 2.220446049250313e-16;
      }
      for (p = lu->ptr[ii]; p < lu->ptr[ii + 1]; ++p) pos[lu->idx[p]] = -1;
    }
    free(pos);
  }
  return f;
}

void bjilu0_solve(const BJIlu0 *f, const double *b, double *x)
{
  const Csr *lu = f->lu;
  int        bl;
#pragma omp parallel for schedule(static, 1)
  for (bl = 0; bl < f->nblocks; ++bl) {
    int lo = f->bstart[bl], hi = f->bstart[bl + 1], i, p;
    for (i = lo; i < hi; ++i) {
      double s = b[i];
      for (p = lu->ptr[i]; p < f->diag[i]; ++p) s -= lu->val[p] * x[lu->idx[p]];
      x[i] = s;
    }
    for (i = hi - 1; i >= lo; --i) {
      double s = x[i];
      for (p = f->diag[i] + 1; p < lu->ptr[i + 1]; ++p) s -= lu->val[p] * x[lu->idx[p]];
      x[i] = s / lu->val[f->diag[i]];
    }
  }
}

void bjilu0_free(BJIlu0 *f)
{
  if (!f) return;
  free(f->bstart);
  free(f->diag);
  csr_free(f->lu);
  free(f);
}

/* ---------------------------------------------------------------- vectors + GMRES */

double vec_dot(int n, const double *a, const double *b)
{
  double s = 0.;
  int    i;
#pragma omp parallel for reduction(+ : s) schedule(static)
  for (i = 0; i < n; ++i) s += a[i] * b[i];
  return s;
}

double vec_norm(int n, const double *a)
{
  return sqrt(vec_dot(n, a, a));
}

static void vec_axpy(int n, double alpha, const double *x, double *y)
{
  int i;
#pragma omp parallel for schedule(static)
  for (i = 0; i < n; ++i) y[i] += alpha * x[i];
}

static void apply_pc(int n, OpFn M, void *mctx, OpFn project, void *pctx, const double *in, double *out)
{
  if (M) M(mctx, in, out);
  else memcpy(out, in, sizeof(double) * (size_t)n);
  if (project) project(pctx, out, out);
}

void gmres(int n, OpFn A, void *actx, OpFn M, void *mctx, OpFn project, void *pctx, const double *b, double *x, int restart, double rtol, double atol, int maxit, int side, KspInfo *info)
{
  int     m  = restart;
  double *V  = (double *)xmalloc(sizeof(double) * (size_t)n * (size_t)(m + 1));
  double *H  = (double *)calloc((size_t)(m + 1) * (size_t)m, sizeof(double));
  double *cs = (double *)xmalloc(sizeof(double) * (size_t)m);
  double *sn = (double *)xmalloc(sizeof(double) * (size_t)m);
  double *g  = (double *)xmalloc(sizeof(double) * (size_t)(m + 1));
  double *y  = (double *)xmalloc(sizeof(double) * (size_t)m);
  double *w  = (double *)xmalloc(sizeof(double) * (size_t)n);
  double *t  = (double *)xmalloc(sizeof(double) * (size_t)n);
  int     its = 0, i, k, done = 0;
  double  rnorm0 = -1., rnorm = 0.;

  info->nhist     = 0;
  info->converged = 0;

  while (!done) {
    /* residual of the current iterate, in the norm this solve monitors */
    A(actx, x, w);
    for (i = 0; i < n; ++i) w[i] = b[i] - w[i];
    if (side == 0) apply_pc(n, M, mctx, project, pctx, w, V);
    else memcpy(V, w, sizeof(double) * (size_t)n);
    rnorm = vec_norm(n, V);
    if (rnorm0 < 0.) {
      rnorm0 = rnorm;
      if (info->nhist < 512) info->hist[info->nhist++] = rnorm;
    }
    if (rnorm <= fmax(rtol * rnorm0, atol) || rnorm == 0.) {
      info->converged = 1;
      break;
    }
    if (its >= maxit) break;
    for (i = 0; i < n; ++i) V[i] /= rnorm;
    memset(g, 0, sizeof(double) * (size_t)(m + 1));
    g[0] = rnorm;

    for (k = 0; k < m && its < maxit; ++k) {
      double *vk = V + (size_t)n * k, *vn = V + (size_t)n * (k + 1);
      int     j, pass;
      if (side == 0) {
        A(actx, vk, t);
        apply_pc(n, M, mctx, project, pctx, t, vn);
      } else {
        apply_pc(n, M, mctx, project, pctx, vk, t);
        A(actx, t, vn);
      }
      for (j = 0; j <= k; ++j) H[j * m + k] = 0.;
      for (pass = 0; pass < 2; ++pass) { /* CGS + one refinement */
        for (j = 0; j <= k; ++j) y[j] = vec_dot(n, V + (size_t)n * j, vn);
        for (j = 0; j <= k; ++j) {
          vec_axpy(n, -y[j], V + (size_t)n * j, vn);
          H[j * m + k] += y[j];
        }
      }
      double hn = vec_norm(n, vn);
      H[(k + 1) * m + k] = hn;
      if (hn > 0.)
        for (i = 0; i < n; ++i) vn[i] /= hn;
      for (j = 0; j < k; ++j) {
        double a = H[j * m + k], bb = H[(j + 1) * m + k];
        H[j * m + k]       = cs[j] * a + sn[j] * bb;
        H[(j + 1) * m + k] = -sn[j] * a + cs[j] * bb;
      }
      {
        double a = H[k * m + k], bb = H[(k + 1) * m + k], r = hypot(a, bb);
        cs[k]              = r > 0. ? a / r : 1.;
        sn[k]              = r > 0. ? bb / r : 0.;
        H[k * m + k]       = r;
        H[(k + 1) * m + k] = 0.;
        g[k + 1]           = -sn[k] * g[k];
        g[k]               = cs[k] * g[k];
      }
      its++;
      rnorm = fabs(g[k + 1]);
      if (info->nhist < 512) info->hist[info->nhist++] = rnorm;
      if (rnorm <= fmax(rtol * rnorm0, atol) || hn == 0.) {
        k++;
        done            = 1;
        info->converged = 1;
        break;
      }
    }
    /* back substitution for the k columns built in this cycle */
    {
      int kk = k, j, l;
      for (j = kk - 1; j >= 0; --j) {
        double s = g[j];
        for (l = j + 1; l < kk; ++l) s -= H[j * m + l] * y[l];
        y[j] = s / H[j * m + j];
      }
      memset(w, 0, sizeof(double) * (size_t)n);
      for (j = 0; j < kk; ++j) vec_axpy(n, y[j], V + (size_t)n * j, w);
      if (side == 1) {
        apply_pc(n, M, mctx, project, pctx, w, t);
        vec_axpy(n, 1., t, x);
      } else vec_axpy(n, 1., w, x);
    }
    if (its >= maxit) done = 1;
  }
  info->its    = its;
  info->rnorm0 = rnorm0;
  info->rnorm  = rnorm;
  free(V);
  free(H);
  free(cs);
  free(sn);
  free(g);
  free(y);
  free(w);
  free(t);
}
