// ibm.cu -- immersed-boundary coupling kernels (see ibm.h for the layout of the method).
//
// No reference code exists for this row of the scope table; the definition is the
// "immersed boundary" section of oracle/src/ns.c, which the parity tests compare these kernels with.
#include "solver.h"
#include <algorithm>
#ifndef FLUCA_HOSTEMU
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#endif

namespace fluca {

// ------------------------------------------------------------------ discrete delta functions
// Peskin's 4-point function (support 2h) and Roma's 3-point function (support 1.5h)
FL_HD double ib_phi(int npts, double r)
{
  r = fabs(r);
  if (npts == 4) {
    if (r < 1.) return 0.125 * (3. - 2. * r + sqrt(1. + 4. * r - 4. * r * r));
    if (r < 2.) return 0.125 * (5. - 2. * r - sqrt(-7. + 12. * r - 4. * r * r));
    return 0.;
  }
  if (r < 0.5) return (1. + sqrt(1. - 3. * r * r)) / 3.;
  if (r < 1.5) {
    const double q = 1. - r;
    return (5. - 3. * r - sqrt(1. - 3. * q * q)) / 6.;
  }
  return 0.;
}

// direction d of one marker: position folded into a periodic domain, first support cell, width of the holding cell
FL_HD void ib_locate(const IbmDev &I, int d, double X, int &base, double &h, double &Xw)
{
  const int     n  = I.nc[d];
  const double *xf = I.xf[d];
  if (I.per[d]) {
    double s = fmod(X - I.x0[d], I.len[d]);
    if (s < 0.) s += I.len[d];
    X = I.x0[d] + s;
  }
  int lo = 0, hi = n; // xf[lo] <= X < xf[hi], clamped to the domain
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (xf[mid] <= X) lo = mid;
    else hi = mid;
  }
  h    = xf[lo + 1] - xf[lo];
  base = (I.npts == 4) ? ((X < I.xc[d][lo]) ? lo - 2 : lo - 1) : lo - 1;
  Xw   = X;
}

// weight of (unwrapped) cell index i along direction d
FL_HD double ib_weight(const IbmDev &I, int d, int i, double Xw, double h)
{
  const int n = I.nc[d];
  double    xc;
  if (i < 0) {
    if (!I.per[d]) return 0.;
    xc = I.xc[d][i + n] - I.len[d];
  } else if (i >= n) {
    if (!I.per[d]) return 0.;
    xc = I.xc[d][i - n] + I.len[d];
  } else xc = I.xc[d][i];
  return ib_phi(I.npts, (xc - Xw) / h);
}

FL_HD int ib_wrap(int i, int n) { return i < 0 ? i + n : (i >= n ? i - n : i); }

// sort key: linear index of the (unwrapped, shifted) first support cell, z-major
FL_HD unsigned long long ib_key(const IbmDev &I, const int base[3])
{
  const unsigned long long ex = (unsigned long long)I.nc[0] + 8, ey = (unsigned long long)I.nc[1] + 8;
  return ((unsigned long long)(base[2] + 4) * ey + (unsigned long long)(base[1] + 4)) * ex + (unsigned long long)(base[0] + 4);
}

IbmDev ibm_dev(const Solver &s)
{
  const Geom &g = s.gh.g;
  const Ibm  &b = s.ibm;
  IbmDev      I;
  memset(&I, 0, sizeof(I));
  I.dim = s.dim, I.npts = b.npts, I.n = b.n;
  I.nc[0] = g.nx, I.nc[1] = g.ny, I.nc[2] = s.dim == 3 ? g.nzg : 1;
  for (int d = 0; d < 3; ++d) {
    I.per[d] = d < s.dim ? g.t[d].per : 0;
    I.xf[d] = b.coord[d][0], I.xc[d] = b.coord[d][1];
    I.x0[d]  = d < s.dim ? s.gh.xf[d].front() : 0.;
    I.len[d] = d < s.dim ? s.gh.xf[d].back() - s.gh.xf[d].front() : 1.;
    I.X[d] = b.X[d], I.Ud[d] = b.Ud[d];
  }
  I.dV = b.dV, I.perm = b.perm, I.lq = b.lq, I.lseg = b.lseg, I.nl = b.nl, I.nlseg = b.nlseg;
  I.k0 = s.dim == 3 ? g.k0 : 0, I.nzl = g.nzl, I.px = g.px, I.py = g.py;
  return I;
}

// ------------------------------------------------------------------ marker upload, sort, segments
namespace {
template <class T>
T *ibm_alloc(Ibm &b, size_t count)
{
  T *p = (T *)dev_alloc(sizeof(T) * (count ? count : 1));
  b.owned.push_back(p);
  return p;
}
void ibm_free_markers(Ibm &b)
{
  for (void *p : b.owned) dev_free(p);
  b.owned.clear();
  for (int d = 0; d < 3; ++d) b.X[d] = b.Ud[d] = b.Um[d] = b.Dl[d] = b.F[d] = nullptr;
  b.dV = b.Umbuf = b.xbuf = b.gbuf = nullptr, b.perm = b.lq = b.lseg = b.own = nullptr;
  b.sh[0] = b.sh[1] = nullptr;
  b.n = 0, b.nl = 0, b.nlseg = 0, b.cap = 0, b.nsh[0] = b.nsh[1] = 0, b.shcap = 0, b.sparse = false;
}
} // namespace

void ibm_destroy(Solver &s)
{
  ibm_free_markers(s.ibm);
  for (int d = 0; d < 3; ++d)
    for (int k = 0; k < 2; ++k) dev_free(s.ibm.coord[d][k]), s.ibm.coord[d][k] = nullptr;
}

void ibm_set_markers(Solver &s, long n, const double *X, const double *Ud, const double *dV, int npts)
{
  Ibm &b = s.ibm;
  if (n < 0 || (n > 0 && (!X || !Ud || !dV))) throw Error(FL_ERR_ARG, "bad marker arrays");
  if (npts != 0 && npts != 3 && npts != 4) throw Error(FL_ERR_ARG, "the discrete delta function has 3 or 4 points");
  if (n >= (1L << 31) / 4) throw Error(FL_ERR_ARG, "too many markers");
  s.ex.sync();
  { // the arrays are laid out with stride n (Um is one contiguous block for the allreduce); the ownership lists depend on
    // the positions, so everything is rebuilt on every call
    ibm_free_markers(b);
    if (n == 0) return;
    const int dim = s.dim;
    double   *blk = ibm_alloc<double>(b, (size_t)n * (5 * dim + 1));
    for (int d = 0; d < dim; ++d) b.X[d] = blk + (size_t)n * d, b.Ud[d] = blk + (size_t)n * (dim + d), b.Um[d] = blk + (size_t)n * (2 * dim + d), b.Dl[d] = blk + (size_t)n * (3 * dim + d), b.F[d] = blk + (size_t)n * (4 * dim + d);
    b.Umbuf = b.Um[0];
    b.dV    = blk + (size_t)n * 5 * dim;
    b.perm  = ibm_alloc<int>(b, (size_t)n);
    b.cap   = n;
  }
  b.n = n, b.npts = npts == 3 ? 3 : 4;
  for (int d = 0; d < s.dim; ++d) {
    if (!b.coord[d][0]) {
      const std::vector<double> &xf = s.gh.xf[d], &xc = s.gh.xc[d];
      b.coord[d][0] = (double *)dev_alloc(sizeof(double) * xf.size());
      b.coord[d][1] = (double *)dev_alloc(sizeof(double) * xc.size());
      copy_h2d(s.ex, b.coord[d][0], xf.data(), sizeof(double) * xf.size());
      copy_h2d(s.ex, b.coord[d][1], xc.data(), sizeof(double) * xc.size());
    }
    copy_h2d(s.ex, b.X[d], X + (size_t)n * d, sizeof(double) * n);
    copy_h2d(s.ex, b.Ud[d], Ud + (size_t)n * d, sizeof(double) * n);
  }
  copy_h2d(s.ex, b.dV, dV, sizeof(double) * n);
  dev_zero(s.ex, b.Umbuf, sizeof(double) * n * 3 * s.dim); // Um, Dl, F
  s.ex.sync();

  // keys of the first support cell
  IbmDev              I = ibm_dev(s);
  unsigned long long *keys = (unsigned long long *)dev_alloc(sizeof(unsigned long long) * 2 * n);
  int                *idx  = (int *)dev_alloc(sizeof(int) * 2 * n);
  int                *flag = idx + n;
  unsigned long long *keys_sorted = keys + n;
  const int           dim = s.dim;
  for_range(s.ex, n, FL_LAMBDA(long m) {
    int base[3] = {0, 0, 0};
    for (int d = 0; d < dim; ++d) {
      double h, Xw;
      ib_locate(I, d, I.X[d][m], base[d], h, Xw);
    }
    keys[m] = ib_key(I, base);
    idx[m]  = (int)m;
  });
  int *perm = b.perm;
  // sorted order on the host: keys in sorted order + the permutation (set_markers is off the hot path; 12 bytes per marker)
  std::vector<unsigned long long> hk((size_t)n);
  std::vector<int>                hperm((size_t)n);
#ifndef FLUCA_HOSTEMU
  {
    size_t tb = 0;
    FL_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, keys, keys_sorted, idx, perm, (int)n, 0, 64, s.ex.stream));
    void *tmp = dev_alloc(tb);
    FL_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tb, keys, keys_sorted, idx, perm, (int)n, 0, 64, s.ex.stream));
    s.ex.stats.launches++;
    copy_d2h(s.ex, hk.data(), keys_sorted, sizeof(unsigned long long) * n);
    copy_d2h(s.ex, hperm.data(), perm, sizeof(int) * n);
    s.ex.sync();
    dev_free(tmp);
    (void)flag;
  }
#else
  {
    std::vector<int> order(n);
    for (long m = 0; m < n; ++m) order[m] = (int)m;
    std::stable_sort(order.begin(), order.end(), [&](int a, int c) { return keys[a] < keys[c]; });
    for (long q = 0; q < n; ++q) perm[q] = order[q], hperm[q] = order[q], hk[q] = keys[order[q]];
    (void)keys_sorted, (void)flag;
  }
#endif
  dev_free(keys), dev_free(idx);

  // ---- slab ownership.  The key is z-major: base plane of the support = key / ((nx + 8)(ny + 8)) - 4.  A marker is LOCAL when
  // one of its npts support planes (wrapped in a periodic z) lies in [k0, k0 + nzl); its other planes belong to the lower or
  // the upper neighbour slab (a support is thinner than a slab, checked below), which makes it SHARED with that neighbour.
  const Geom &g = s.gh.g;
  const int   nranks = s.comm->nranks, np = b.npts;
  const bool  perz = s.dim == 3 && g.t[2].per;
  const unsigned long long exy = ((unsigned long long)g.nx + 8) * ((unsigned long long)g.ny + 8);
  std::vector<int> lq, lseg, shl[2], own((size_t)n, 0);
  bool             try_sparse = nranks > 1 && s.dim == 3;
  for (long q = 0; q < n; ++q) {
    bool local = true, dn = false, up = false;
    if (s.dim == 3 && nranks > 1) {
      const int base = (int)(hk[q] / exy) - 4;
      // the image of the support (shifted by a period if need be) that overlaps the slab
      int shift = 0;
      local = false;
      for (int sgn = 0; sgn < (perz ? 3 : 1) && !local; ++sgn) {
        shift = sgn == 0 ? 0 : (sgn == 1 ? g.nzg : -g.nzg);
        for (int t = 0; t < np; ++t) {
          const int k = base + t + shift;
          if (!perz && (base + t < 0 || base + t >= g.nzg)) continue;
          if (k >= g.k0 && k < g.k0 + g.nzl) local = true;
        }
      }
      if (local)
        for (int t = 0; t < np; ++t) {
          if (!perz && (base + t < 0 || base + t >= g.nzg)) continue; // outside the domain: zero weight
          const int k = base + t + shift;
          if (k < g.k0) dn = true;
          if (k >= g.k0 + g.nzl) up = true;
        }
    }
    if (!local) continue;
    if (lq.empty() || hk[q] != hk[lq.back()]) lseg.push_back((int)lq.size());
    lq.push_back((int)q);
    if (dn) shl[0].push_back(hperm[q]);
    if (up) shl[1].push_back(hperm[q]);
    if (!dn) own[hperm[q]] = 1; // of the two ranks that share a marker the lower one reports it
  }
  lseg.push_back((int)lq.size());
  if (try_sparse) {
    // every rank must be able to take part: slabs at least as thick as a support; the exchange count is the largest list
    double *tmp = (double *)dev_alloc(sizeof(double) * 3 * (nranks + 1));
    const double mine[3] = {g.nzl >= np ? 1. : 0., (double)shl[0].size(), (double)shl[1].size()};
    copy_h2d(s.ex, tmp, mine, sizeof(mine));
    s.comm->allgather(s.ex, tmp, tmp + 3, 3);
    std::vector<double> all((size_t)3 * nranks);
    copy_d2h(s.ex, all.data(), tmp + 3, sizeof(double) * 3 * nranks);
    s.ex.sync();
    dev_free(tmp);
    long cap = 0;
    for (int r = 0; r < nranks; ++r) {
      if (all[3 * r] == 0.) try_sparse = false;
      cap = std::max(cap, (long)std::max(all[3 * r + 1], all[3 * r + 2]));
    }
    b.shcap = cap;
  }
  b.sparse = try_sparse;
  if (!b.sparse) {
    // replicated sums: every rank walks every marker (partial sums over its own planes) and one allreduce completes them
    lq.resize(n);
    lseg.clear();
    for (long q = 0; q < n; ++q) {
      if (q == 0 || hk[q] != hk[q - 1]) lseg.push_back((int)q);
      lq[q] = (int)q;
    }
    lseg.push_back((int)n);
    shl[0].clear(), shl[1].clear();
    std::fill(own.begin(), own.end(), 1);
    b.shcap = 0;
  }
  b.nl = (long)lq.size(), b.nlseg = (int)lseg.size() - 1;
  b.lq   = ibm_alloc<int>(b, lq.size());
  b.lseg = ibm_alloc<int>(b, lseg.size());
  b.own  = ibm_alloc<int>(b, (size_t)n);
  b.gbuf = ibm_alloc<double>(b, (size_t)n * s.dim);
  copy_h2d(s.ex, b.lq, lq.data(), sizeof(int) * lq.size());
  copy_h2d(s.ex, b.lseg, lseg.data(), sizeof(int) * lseg.size());
  copy_h2d(s.ex, b.own, own.data(), sizeof(int) * n);
  for (int sd = 0; sd < 2; ++sd) {
    b.nsh[sd] = (long)shl[sd].size();
    b.sh[sd]  = ibm_alloc<int>(b, shl[sd].size());
    copy_h2d(s.ex, b.sh[sd], shl[sd].data(), sizeof(int) * shl[sd].size());
  }
  b.xbuf = ibm_alloc<double>(b, (size_t)4 * b.shcap * s.dim);
  s.ex.sync();
}

// ------------------------------------------------------------------ gather / scatter
#ifndef FLUCA_HOSTEMU
namespace {
const unsigned FULL = 0xffffffffu;

// the lanes of a warp share one marker: lanes 0..DIM-1 locate it, lanes 0..4*DIM-1 evaluate one 1-D weight each
template <int DIM>
__device__ __forceinline__ void warp_marker(const IbmDev &I, int m, int lane, int base[3], double &wl)
{
  int    b = 0;
  double h = 1., Xw = 0.;
  if (lane < DIM) ib_locate(I, lane, I.X[lane][m], b, h, Xw);
  const int    dl = lane >> 2, ql = lane & 3, src = dl < DIM ? dl : 0;
  const int    bd = __shfl_sync(FULL, b, src);
  const double hd = __shfl_sync(FULL, h, src), Xd = __shfl_sync(FULL, Xw, src);
  wl = (dl < DIM && ql < I.npts) ? ib_weight(I, dl, bd + ql, Xd, hd) : 0.;
  base[0] = __shfl_sync(FULL, b, 0), base[1] = __shfl_sync(FULL, b, 1), base[2] = DIM == 3 ? __shfl_sync(FULL, b, 2) : 0;
}

// support point p of a marker: weight product (all lanes call) and padded-array index of the cell, -1 if this rank
// does not hold it
template <int DIM>
__device__ __forceinline__ double warp_point(const IbmDev &I, int p, bool act, const int base[3], double wl, int &cell, int ijk[3])
{
  const int np = I.npts, pp = act ? p : 0;
  const int qx = pp % np, qy = (pp / np) % np, qz = DIM == 3 ? pp / (np * np) : 0;
  double    ww = __shfl_sync(FULL, wl, qx) * __shfl_sync(FULL, wl, 4 + qy);
  if (DIM == 3) ww *= __shfl_sync(FULL, wl, 8 + qz);
  cell = -1;
  if (!act || ww == 0.) return 0.;
  int i = base[0] + qx, j = base[1] + qy, k = DIM == 3 ? base[2] + qz : 0;
  if (I.per[0]) i = ib_wrap(i, I.nc[0]);
  if (I.per[1]) j = ib_wrap(j, I.nc[1]);
  if (DIM == 3 && I.per[2]) k = ib_wrap(k, I.nc[2]);
  const int kl = k - I.k0;
  if (kl < 0 || kl >= I.nzl) return 0.;
  ijk[0] = i, ijk[1] = j, ijk[2] = k;
  cell = i + I.px * (j + I.py * (kl + 1));
  return ww;
}

template <int DIM>
__global__ void __launch_bounds__(256) k_ibm_interp(const IbmDev I, const CV3 v, double *um0, double *um1, double *um2)
{
  const int  lane = threadIdx.x & 31;
  const long wid = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((long)gridDim.x * blockDim.x) >> 5;
  const int  npd = DIM == 3 ? I.npts * I.npts * I.npts : I.npts * I.npts;
  double    *um[3] = {um0, um1, um2};
  for (long t = wid; t < I.nl; t += nw) {
    const int m = I.perm[I.lq[t]];
    int       base[3];
    double    wl;
    warp_marker<DIM>(I, m, lane, base, wl);
    double acc[DIM];
#pragma unroll
    for (int c = 0; c < DIM; ++c) acc[c] = 0.;
    for (int p0 = 0; p0 < npd; p0 += 32) {
      int          cell, ijk[3];
      const double ww = warp_point<DIM>(I, p0 + lane, p0 + lane < npd, base, wl, cell, ijk);
      if (cell >= 0) {
#pragma unroll
        for (int c = 0; c < DIM; ++c) acc[c] += ww * v.c[c][cell];
      }
    }
#pragma unroll
    for (int c = 0; c < DIM; ++c) {
      double a = acc[c];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(FULL, a, o);
      if (lane == 0) um[c][m] = a;
    }
  }
}

template <int DIM>
__global__ void __launch_bounds__(256) k_ibm_spread(const IbmDev I, const double *f0, const double *f1, const double *f2, const V3 out, const V3 out2)
{
  const int     lane = threadIdx.x & 31;
  const long    wid = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((long)gridDim.x * blockDim.x) >> 5;
  const int     npd = DIM == 3 ? I.npts * I.npts * I.npts : I.npts * I.npts;
  const double *fm[3] = {f0, f1, f2};
  for (long sg = wid; sg < I.nlseg; sg += nw) {
    const int q0 = I.lseg[sg], q1 = I.lseg[sg + 1];
    double    acc[2][DIM];
    int       cells[2];
    double    ivol[2];
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      cells[t] = -1, ivol[t] = 0.;
#pragma unroll
      for (int c = 0; c < DIM; ++c) acc[t][c] = 0.;
    }
    for (int q = q0; q < q1; ++q) {
      const int m = I.perm[I.lq[q]];
      int       base[3];
      double    wl;
      warp_marker<DIM>(I, m, lane, base, wl);
      const double dv = I.dV[m];
      double       fv[DIM];
#pragma unroll
      for (int c = 0; c < DIM; ++c) fv[c] = fm[c][m] * dv;
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        if (t * 32 >= npd) break;
        int          cell, ijk[3];
        const double ww = warp_point<DIM>(I, t * 32 + lane, t * 32 + lane < npd, base, wl, cell, ijk);
        if (cell >= 0) {
          // every marker of the segment has the same first support cell, hence the same target cells
          if (cells[t] < 0) {
            double vol = (I.xf[0][ijk[0] + 1] - I.xf[0][ijk[0]]) * (I.xf[1][ijk[1] + 1] - I.xf[1][ijk[1]]);
            if (DIM == 3) vol *= I.xf[2][ijk[2] + 1] - I.xf[2][ijk[2]];
            cells[t] = cell, ivol[t] = 1. / vol;
          }
#pragma unroll
          for (int c = 0; c < DIM; ++c) acc[t][c] += ww * fv[c];
        }
      }
    }
#pragma unroll
    for (int t = 0; t < 2; ++t)
      if (cells[t] >= 0) {
#pragma unroll
        for (int c = 0; c < DIM; ++c) {
          const double a = acc[t][c] * ivol[t];
          atomicAdd(out.c[c] + cells[t], a);
          if (out2.c[0]) atomicAdd(out2.c[c] + cells[t], a);
        }
      }
  }
}
} // namespace
#else
namespace {
// test double: the same arithmetic, marker by marker
template <int DIM>
void host_transfer(const IbmDev &I, int mode, const CV3 &v, double *const um[3], double *const fm[3], const V3 &out, const V3 &out2)
{
  const int np = I.npts;
  for (long q = 0; q < I.nl; ++q) {
    const int m = I.perm[I.lq[q]];
    int       base[3] = {0, 0, 0};
    double    w[3][4] = {{1., 0., 0., 0.}, {1., 0., 0., 0.}, {1., 0., 0., 0.}};
    for (int d = 0; d < DIM; ++d) {
      double h, Xw;
      ib_locate(I, d, I.X[d][m], base[d], h, Xw);
      for (int t = 0; t < np; ++t) w[d][t] = ib_weight(I, d, base[d] + t, Xw, h);
    }
    double acc[3] = {0., 0., 0.};
    for (int qz = 0; qz < (DIM == 3 ? np : 1); ++qz)
      for (int qy = 0; qy < np; ++qy)
        for (int qx = 0; qx < np; ++qx) {
          const double ww = w[0][qx] * w[1][qy] * (DIM == 3 ? w[2][qz] : 1.);
          if (ww == 0.) continue;
          int i = base[0] + qx, j = base[1] + qy, k = DIM == 3 ? base[2] + qz : 0;
          if (I.per[0]) i = ib_wrap(i, I.nc[0]);
          if (I.per[1]) j = ib_wrap(j, I.nc[1]);
          if (DIM == 3 && I.per[2]) k = ib_wrap(k, I.nc[2]);
          const int kl = k - I.k0;
          if (kl < 0 || kl >= I.nzl) continue;
          const long cell = i + (long)I.px * (j + (long)I.py * (kl + 1));
          if (mode == 0) {
            for (int c = 0; c < DIM; ++c) acc[c] += ww * v.c[c][cell];
          } else {
            double vol = (I.xf[0][i + 1] - I.xf[0][i]) * (I.xf[1][j + 1] - I.xf[1][j]);
            if (DIM == 3) vol *= I.xf[2][k + 1] - I.xf[2][k];
            for (int c = 0; c < DIM; ++c) {
              const double a = ww * fm[c][m] * I.dV[m] / vol;
              out.c[c][cell] += a;
              if (out2.c[0]) out2.c[c][cell] += a;
            }
          }
        }
    if (mode == 0)
      for (int c = 0; c < DIM; ++c) um[c][m] = acc[c];
  }
}
} // namespace
#endif

void ibm_interpolate(Solver &s, const V3 &v)
{
  Ibm &b = s.ibm;
  if (b.n <= 0) return;
  KScope       ks(s.ex, KT_IBM);
  const IbmDev I = ibm_dev(s);
  s.ex.stats.launches++;
#ifndef FLUCA_HOSTEMU
  {
    KTimer kt(s.ex, s.ex.kt_current);
    long   blocks = (b.nl * 32 + 255) / 256, cap = (long)s.ex.sm_count * 16;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    if (s.dim == 2) k_ibm_interp<2><<<(unsigned)blocks, 256, 0, s.ex.stream>>>(I, CV3(v), b.Um[0], b.Um[1], b.Um[2]);
    else k_ibm_interp<3><<<(unsigned)blocks, 256, 0, s.ex.stream>>>(I, CV3(v), b.Um[0], b.Um[1], b.Um[2]);
    FL_CUDA(cudaGetLastError());
  }
#else
  {
    V3 none;
    none.c[0] = none.c[1] = none.c[2] = nullptr;
    if (s.dim == 2) host_transfer<2>(I, 0, CV3(v), b.Um, nullptr, none, none);
    else host_transfer<3>(I, 0, CV3(v), b.Um, nullptr, none, none);
  }
#endif
  // every rank summed over its own planes only
  if (!b.sparse) {
    s.comm->allsum(s.ex, b.Umbuf, (int)(b.n * s.dim)); // replicated sums: one allreduce over all markers
    return;
  }
  if (b.shcap <= 0) return;
  // ownership exchange: the partial sums of the markers whose support crosses a slab face go to that neighbour, its partial
  // sums come back, and both hold the complete velocity (they both need it: each spreads the force into its own planes)
  const int  dim = s.dim;
  const long cap = b.shcap, n0 = b.nsh[0], n1 = b.nsh[1];
  double    *sd = b.xbuf, *rd = b.xbuf + cap * dim, *su = b.xbuf + 2 * cap * dim, *ru = b.xbuf + 3 * cap * dim;
  const int *l0 = b.sh[0], *l1 = b.sh[1];
  double    *u0 = b.Um[0], *u1 = b.Um[1], *u2 = b.Um[2];
  for_range(s.ex, cap * dim, FL_LAMBDA(long e) {
    const long   k = e % cap;
    const int    c = (int)(e / cap);
    const double *u = c == 0 ? u0 : (c == 1 ? u1 : u2);
    sd[e] = k < n0 ? u[l0[k]] : 0.;
    su[e] = k < n1 ? u[l1[k]] : 0.;
  });
  s.comm->sendrecv(s.ex, sd, rd, su, ru, cap * dim, s.gh.g.t[2].per != 0);
  for_range(s.ex, cap * dim, FL_LAMBDA(long e) {
    const long k = e % cap;
    const int  c = (int)(e / cap);
    double    *u = c == 0 ? u0 : (c == 1 ? u1 : u2);
    if (k < n0) u[l0[k]] += rd[e];
    if (k < n1) u[l1[k]] += ru[e];
  });
}

void ibm_gather_global(Solver &s, double *const src[3], double *host)
{
  Ibm &b = s.ibm;
  if (b.n <= 0) return;
  const long n = b.n;
  if (b.sparse) {
    const int *own = b.own;
    for (int c = 0; c < s.dim; ++c) {
      double       *gb = b.gbuf + (size_t)n * c;
      const double *sc = src[c];
      for_range(s.ex, n, FL_LAMBDA(long m) { gb[m] = own[m] ? sc[m] : 0.; });
    }
    s.comm->allsum(s.ex, b.gbuf, (int)(n * s.dim));
    copy_d2h(s.ex, host, b.gbuf, sizeof(double) * n * s.dim);
  } else {
    for (int c = 0; c < s.dim; ++c) copy_d2h(s.ex, host + (size_t)n * c, src[c], sizeof(double) * n);
  }
  s.ex.sync();
}

void ibm_spread(Solver &s, double *const Fm[3], const V3 &f, const V3 *f2)
{
  V3 o2;
  o2.c[0] = o2.c[1] = o2.c[2] = nullptr;
  if (f2) o2 = *f2;
  Ibm &b = s.ibm;
  if (b.n <= 0) return;
  KScope       ks(s.ex, KT_IBM);
  const IbmDev I = ibm_dev(s);
  s.ex.stats.launches++;
#ifndef FLUCA_HOSTEMU
  {
    KTimer kt(s.ex, s.ex.kt_current);
    long   blocks = ((long)b.nlseg * 32 + 255) / 256, cap = (long)s.ex.sm_count * 16;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    if (s.dim == 2) k_ibm_spread<2><<<(unsigned)blocks, 256, 0, s.ex.stream>>>(I, Fm[0], Fm[1], nullptr, f, o2);
    else k_ibm_spread<3><<<(unsigned)blocks, 256, 0, s.ex.stream>>>(I, Fm[0], Fm[1], Fm[2], f, o2);
    FL_CUDA(cudaGetLastError());
  }
#else
  {
    if (s.dim == 2) host_transfer<2>(I, 1, CV3(), nullptr, Fm, f, o2);
    else host_transfer<3>(I, 1, CV3(), nullptr, Fm, f, o2);
  }
#endif
}

struct IbP3 {
  double *c[3];
};

// direct forcing with an implicit predictor (DESIGN.md "IBM coupling"; oracle/src/ns.c ib_force_rhs)
void ibm_force_rhs(Solver &s)
{
  Ibm &b = s.ibm;
  if (b.n <= 0) return;
  momentum_solve(s, s.rm, s.vstar, s.have_guess); // predictor v~ = A^-1 r_mom (guess: previous velocity, do_step)
  const int    dim = s.dim, passes = b.iters > 1 ? b.iters : 1;
  const long   nl = b.nl;
  const int   *perm = b.perm, *lq = b.lq;
  const double fscale = s.sp.rho / s.sp.dt;
  const double *dv = b.dV;
  const IbP3   UD = {{b.Ud[0], b.Ud[1], b.Ud[2]}}, UM = {{b.Um[0], b.Um[1], b.Um[2]}}, DL = {{b.Dl[0], b.Dl[1], b.Dl[2]}}, FF = {{b.F[0], b.F[1], b.F[2]}};
  // multi-direct forcing: every pass interpolates the corrected predictor and spreads the remaining slip
  for (int it = 0; it < passes; ++it) {
    ibm_interpolate(s, s.vstar);
    {
      KScope ks(s.ex, KT_IBM);
      for_range(s.ex, nl, FL_LAMBDA(long t) {
        const int m = perm[lq[t]];
        for (int c = 0; c < dim; ++c) {
          const double d = UD.c[c][m] - UM.c[c][m];
          DL.c[c][m]     = d;
          FF.c[c][m]     = (it == 0 ? 0. : FF.c[c][m]) + fscale * d * dv[m]; // force of marker m on the fluid
        }
      });
    }
    // the increment also goes into the predictor: multi-direct forcing needs it, and v~ + f is the guess of the first
    // momentum solve of the step (A (v~ + f) = r_mom + f + O(dt) f)
    ibm_spread(s, b.Dl, s.rm, &s.vstar);
  }
  s.have_guess = s.allow_guess;
}

} // namespace fluca
