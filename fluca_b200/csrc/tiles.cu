// tiles.cu -- the TMA-staged versions of the hot 3-D stencil operators (see tma.h for the pipeline).
//
//   momentum operator  y = A x (+ two fused dot products)   reference: MatMult with the assembled A of
//                      NSFormJacobian(UPDATE), cnlinearcart3d.c:2930-2941, inside KSPSolve(kspA) abfpc.c:72
//   Poisson operator   q = P p (+ <a, q>)                   reference: MatMult with S inside KSPSolve(kspS) abfpc.c:77
//
// The arithmetic is that of a_apply_core / poisson_apply_cell in stencil.h (same tables, same
// operation order per direction); only the operand source differs: shared-memory tiles filled by
// TMA instead of global loads.  Planes that touch a physical z wall need the one-sided 4-point
// rows (cells k +- 2) and are computed by the direct-load kernel on those two planes; their
// partial sums are chained into the TMA launch through the `carry` argument of the reduction.
#include "solver.h"
#ifndef FLUCA_HOSTEMU
#include "tma.h"
#include <map>
#include <mutex>
#include <tuple>

namespace fluca {

// ------------------------------------------------------------------ tensor maps
namespace {
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn()
{
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void                           *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    // resolved through the runtime: the library does not link libcuda
    FL_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    if (!p || q != cudaDriverEntryPointSuccess) throw Error(FL_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    fn = (EncodeTiledFn)p;
  }
  return fn;
}

struct MapKey {
  const double *p;
  int           px, py, nz;
  bool operator<(const MapKey &o) const { return std::tie(p, px, py, nz) < std::tie(o.p, o.px, o.py, o.nz); }
};
std::map<MapKey, CUtensorMap> g_maps;
std::mutex                    g_maps_mutex;
} // namespace

const CUtensorMap &tensor_map_for(const double *field, int px, int py, int nplanes)
{
  std::lock_guard<std::mutex> lock(g_maps_mutex);
  MapKey                      key = {field, px, py, nplanes};
  auto                        it  = g_maps.find(key);
  if (it != g_maps.end()) return it->second;
  alignas(64) CUtensorMap m;
  const cuuint64_t        dims[3]    = {(cuuint64_t)px, (cuuint64_t)py, (cuuint64_t)nplanes};
  const cuuint64_t        strides[2] = {(cuuint64_t)px * sizeof(double), (cuuint64_t)px * py * sizeof(double)};
  const cuuint32_t        box[3]     = {(cuuint32_t)TLX, (cuuint32_t)TLY, 1u};
  const cuuint32_t        estr[3]    = {1u, 1u, 1u};
  CUresult r = encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, (void *)field, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw Error(FL_ERR_CUDA, "cuTensorMapEncodeTiled failed with code " + std::to_string((int)r));
  return g_maps.emplace(key, m).first->second;
}

void tensor_map_forget(const double *field)
{
  std::lock_guard<std::mutex> lock(g_maps_mutex);
  for (auto it = g_maps.begin(); it != g_maps.end();) {
    if (it->first.p == field) it = g_maps.erase(it);
    else ++it;
  }
}

bool tma_usable(const Solver &s)
{
  static const bool off = getenv("FLUCA_B200_NO_TMA") != nullptr;
  const Geom       &g   = s.gh.g;
  // periodic x is served by the wrap fix-up of the first / last tile of a row (tma.h); periodic y is not
  return !off && s.dim == 3 && !g.t[1].per && g.nx >= TMX && g.ny >= TMY;
}

// ------------------------------------------------------------------ momentum operator from shared-memory tiles
// field order inside a ring slot: x0 x1 x2 | v0_0 v0_1 v0_2 | U0_0 U0_1 U0_2
template <bool BND>
__device__ __forceinline__ void a_apply_tile(const Geom &g, const StepParams &sp, const BcDev &bc, const TileView &tv, int i, int j, int kl, double y[3])
{
  constexpr int FS = TILE_STRIDE;
  const int     lc = tv.lc;
  double        xc[3], vc[3], conv[3], lap[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    xc[c]   = tv.p0[c * FS + lc];
    vc[c]   = tv.p0[(3 + c) * FS + lc];
    conv[c] = 0.;
    lap[c]  = 0.;
  }
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    const Tab    &T  = g.t[d];
    const int     ig = d == 0 ? i : (d == 1 ? j : g.k0 + kl);
    // planes that touch a z wall never reach this function (a_apply_dots sends them to the direct-load kernel)
    const bool    lo = BND && d < 2 && !T.per && ig == 0, hi = BND && d < 2 && !T.per && ig == T.n - 1;
    const double *nm = d == 2 ? tv.pm : tv.p0, *np = d == 2 ? tv.pp : tv.p0;
    const int     st = d == 0 ? 1 : (d == 1 ? TLX : 0);
    const int     om = lc - st, op = lc + st;
    const double  hh = 0.5 * FL_LDG(T.hinv + ig);
    const double  Ul = tv.p0[(6 + d) * FS + lc], Uu = np[(6 + d) * FS + op];
    const double  al = FL_LDG(T.itw + 2 * ig), bl = FL_LDG(T.itw + 2 * ig + 1), au = FL_LDG(T.itw + 2 * ig + 2), bu = FL_LDG(T.itw + 2 * ig + 3);
    const double  lw0 = FL_LDG(T.lapw + (size_t)ig * 3), lw1 = FL_LDG(T.lapw + (size_t)ig * 3 + 1), lw2 = FL_LDG(T.lapw + (size_t)ig * 3 + 2);
    double        xm[3], xp[3], vm[3], vp[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      xm[c] = nm[c * FS + om];
      xp[c] = np[c * FS + op];
      vm[c] = nm[(3 + c) * FS + om];
      vp[c] = np[(3 + c) * FS + op];
    }
    const long   pt  = (lo || hi) ? bc_pt(g, 2 * d, i, j, kl) : 0;
    const double Ild = lo ? T.cv2_lo[0] * xc[d] + T.cv2_lo[1] * xp[d] : al * xm[d] + bl * xc[d];
    const double Iud = hi ? T.cv2_hi[0] * xm[d] + T.cv2_hi[1] * xc[d] : au * xc[d] + bu * xp[d];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const int w = (c == d) ? 1 : 0;
      double    vbl, vbu;
      if (lo) {
        vbl = T.it_lo[w][0] * vc[c] + T.it_lo[w][1] * vp[c];
        if (T.it_lo_bc != 0.) vbl += bc.vel[2 * d][0][c * bc.npts[2 * d] + pt];
      } else vbl = al * vm[c] + bl * vc[c];
      if (hi) {
        vbu = T.it_hi[w][0] * vm[c] + T.it_hi[w][1] * vc[c];
        if (T.it_hi_bc != 0.) vbu += bc.vel[2 * d + 1][0][c * bc.npts[2 * d + 1] + pt];
      } else vbu = au * vc[c] + bu * vp[c];
      const double Ilc = lo ? T.cv1_lo[w][0] * xc[c] + T.cv1_lo[w][1] * xp[c] : al * xm[c] + bl * xc[c];
      const double Iuc = hi ? T.cv1_hi[w][0] * xm[c] + T.cv1_hi[w][1] * xc[c] : au * xc[c] + bu * xp[c];
      conv[c] += hh * (Uu * Iuc + vbu * Iud - Ul * Ilc - vbl * Ild);
      double l;
      if (lo || hi) {
        const double *lw = T.lapw + ((size_t)w * T.n + ig) * 3;
        l                = lw[0] * xm[c] + lw[1] * xc[c] + lw[2] * xp[c];
        if (lo) l += T.lap_lo2[w] * tv.p0[c * FS + lc + 2 * st];
        if (hi) l += T.lap_hi2[w] * tv.p0[c * FS + lc - 2 * st];
      } else l = lw0 * xm[c] + lw1 * xc[c] + lw2 * xp[c];
      lap[c] += l;
    }
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) y[c] = xc[c] + sp.dt * conv[c] - sp.nu2 * lap[c];
}

// Interior cells of a mesh that is uniform in all three directions: every interior table row is a constant
// (interpolation 1/2, 1/2; second derivative (1, -2, 1)/h^2; flux divergence 1/h), so the operator needs no table
// load at all and the sums x_c + x_neighbour are shared between the convection and the diffusion terms:
//   conv_c = sum_d 1/(4 h_d) [ 2 Uu (x_c + x_c^+) - 2 Ul (x_c^- + x_c) + (v_c + v_c^+)(x_d + x_d^+)/... ]   (same algebra
// as a_apply_tile<false> with al = bl = au = bu = 1/2; the results differ from the table path by rounding only)
struct UniCoef {
  double q[3]; // 1 / (4 h_d)
  double l[3]; // 1 / h_d^2
  double l4;   // 4 (l_0 + l_1 + l_2)
};

__device__ __forceinline__ void a_apply_tile_uniform(const UniCoef &u, const StepParams &sp, const TileView &tv, double y[3])
{
  constexpr int FS = TILE_STRIDE;
  const int     lc = tv.lc;
  double        xc[3], vc[3], conv[3], lap[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    xc[c]   = tv.p0[c * FS + lc];
    vc[c]   = tv.p0[(3 + c) * FS + lc];
    conv[c] = 0.;
    lap[c]  = 0.;
  }
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    const double *nm = d == 2 ? tv.pm : tv.p0, *np = d == 2 ? tv.pp : tv.p0;
    const int     st = d == 0 ? 1 : (d == 1 ? TLX : 0);
    const int     om = lc - st, op = lc + st;
    const double  Ul = tv.p0[(6 + d) * FS + lc], Uu = np[(6 + d) * FS + op];
    double        sxl[3], sxu[3], svl[3], svu[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      sxl[c] = nm[c * FS + om] + xc[c];
      sxu[c] = xc[c] + np[c * FS + op];
      svl[c] = nm[(3 + c) * FS + om] + vc[c];
      svu[c] = vc[c] + np[(3 + c) * FS + op];
    }
    const double hl = 0.5 * sxl[d], hu = 0.5 * sxu[d];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      double t = Uu * sxu[c] - Ul * sxl[c];
      t        = fma(svu[c], hu, t);
      t        = fma(-svl[c], hl, t);
      conv[c]  = fma(u.q[d], t, conv[c]);
      lap[c]   = fma(u.l[d], sxl[c] + sxu[c], lap[c]);
    }
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) y[c] = xc[c] + sp.dt * conv[c] - sp.nu2 * (lap[c] - u.l4 * xc[c]);
}

template <int NRED>
struct AApplyTile : TileOpDefaults {
  static const int NIN = 9, NR = NRED, MINB = 2, STAGES = 4;
  Geom             g;
  StepParams       sp;
  BcDev            bc;
  UniCoef          uc;
  int              uniform; // all three directions uniform: interior warps use a_apply_tile_uniform
  const double    *a[3]; // dot partner (NR == 2); nullptr: the partner is x itself
  double          *y[3];
  double          *wout[3]; // coupled velocity block: on entry g = (dt/rho) G p; y = A x + g, wout = x + g (nullptr otherwise)
  int              keep_w = 1; // 0: wout is only read
  struct Regs {
    double a[3];
  };
  // bit 0: no wall in reach in x and y; bit 1: x-wall column, computed by the transposed wall launch instead (a single
  // wall lane would send its whole warp down the table-driven path: 12-25 % of the warps, profiles/r01p)
  __device__ int flags(int i, int j) const
  {
    const bool xin = g.t[0].per || (i > 0 && i < g.nx - 1);
    return ((xin && j > 0 && j < g.ny - 1) ? 1 : 0) | ((!g.t[0].per && (i == 0 || i == g.nx - 1)) ? 2 : 0);
  }
  __device__ void prefetch(Regs &rg, int off, int kl) const
  {
    if (NRED > 0) {
#pragma unroll
      for (int q = 0; q < 3; ++q)
        if (a[q]) rg.a[q] = a[q][off];
    } else if (wout[0]) {
#pragma unroll
      for (int q = 0; q < 3; ++q) rg.a[q] = wout[q][off];
    }
  }
  __device__ void cell(const TileView &tv, const Regs &rg, int fl, int i, int j, int kl, int c, double *acc) const
  {
    double     r[3];
    if (fl & 2) return;
    if (__all_sync(__activemask(), (fl & 1) != 0)) {
      if (uniform) a_apply_tile_uniform(uc, sp, tv, r);
      else a_apply_tile<false>(g, sp, bc, tv, i, j, kl, r);
    } else a_apply_tile<true>(g, sp, bc, tv, i, j, kl, r);
    double    d0 = 0., d1 = 0., d2 = 0., d3 = 0.;
    if (NRED == 0 && wout[0]) {
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        y[q][c] = r[q] + rg.a[q];
        if (keep_w) wout[q][c] = tv.p0[q * TILE_STRIDE + tv.lc] + rg.a[q];
      }
      return;
    }
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      y[q][c] = r[q];
      if (NRED > 0) {
        const double xv = tv.p0[q * TILE_STRIDE + tv.lc], av = a[q] ? rg.a[q] : xv;
        d0 += av * r[q];
        d1 += r[q] * r[q];
        d2 += xv * r[q];
        d3 += av * xv;
      }
    }
    if (NRED > 0) acc[0] += d0, acc[1] += d1, acc[2] += d2, acc[3] += d3;
  }
};

// shifts the plane index of a box functor (the z-wall planes are launched as 1-plane boxes)
template <class F>
struct PlaneAt {
  F   f;
  int kl;
  FL_HD void operator()(int i, int j, int, double *acc) const { f(i, j, kl, acc); }
};

// the two x-wall columns of planes [kbeg, kend): a block covers 32 rows j x 8 planes, so that a warp holds 32 wall cells
// and all eight warps of a block work (box = {ny, planes, 1})
template <class F>
struct XWallAt {
  F   f;
  int i, kbeg;
  FL_HD void operator()(int a, int b, int, double *acc) const { f(i, a, kbeg + b, acc); }
};

// y = A x ; out = {<a, y>, <y, y>, <x, y>, <a, x>} are left in ex.d_result (reduce_finish reads them)
void a_apply_dots_tma(Solver &s, const V3 &x, const V3 &y, const V3 &a, bool with_dots)
{
  const Geom &g  = s.gh.g;
  const bool  wl = g.t[2].wall_lo && !g.t[2].per, wh = g.t[2].wall_hi && !g.t[2].per;
  KGroup      grp(s.ex, KT_MOMENTUM_APPLY);
  // 1. the (up to two) planes at physical z walls: direct-load kernel with the one-sided rows
  const double *carry = nullptr;
  int           ncar  = 0;
  AApplyDots<3> f;
  f.g = g, f.sp = s.sp, f.bc = s.bc, f.x = CV3(x), f.v0 = CV3(s.v0), f.U0 = CV3(s.U0), f.a = CV3(with_dots ? a : x), f.y = y;
  const Box plane_box = {g.nx, g.ny, 1};
  const int kbeg = wl ? 1 : 0;
  int       kend = wh ? g.nzl - 1 : g.nzl;
  if (kend < kbeg) kend = kbeg;
  if (wl) {
    PlaneAt<AApplyDots<3>> pf = {f, 0};
    double                *res = s.ex.d_carry + Exec::MAXR * ncar;
    for_box_reduce<4>(s.ex, plane_box, pf, carry, res);
    carry = res, ++ncar;
  }
  if (wh && g.nzl - 1 >= kbeg) {
    PlaneAt<AApplyDots<3>> pf = {f, g.nzl - 1};
    double                *res = s.ex.d_carry + Exec::MAXR * ncar;
    for_box_reduce<4>(s.ex, plane_box, pf, carry, res);
    carry = res, ++ncar;
  }
  // 2. the x-wall columns of the remaining planes
  if (kend > kbeg && !g.t[0].per) {
    const Box wall_box = {g.ny, kend - kbeg, 1};
    for (int side = 0; side < 2; ++side) {
      XWallAt<AApplyDots<3>> xf = {f, side ? g.nx - 1 : 0, kbeg};
      double                *res = s.ex.d_carry + Exec::MAXR * ncar;
      for_box_reduce<4>(s.ex, wall_box, xf, carry, res);
      carry = res, ++ncar;
    }
  }
  // 3. everything else through the TMA pipeline
  const double *fields[9] = {x.c[0], x.c[1], x.c[2], s.v0.c[0], s.v0.c[1], s.v0.c[2], s.U0.c[0], s.U0.c[1], s.U0.c[2]};
  UniCoef       uc;
  static const bool no_uni = getenv("FLUCA_B200_NO_UNIFORM") != nullptr;
  const int     uniform = (!no_uni && g.t[0].uni && g.t[1].uni && g.t[2].uni) ? 1 : 0;
  uc.l4 = 0.;
  for (int d = 0; d < 3; ++d) uc.q[d] = 0.25 / g.t[d].uh, uc.l[d] = 1. / (g.t[d].uh * g.t[d].uh), uc.l4 += 4. * uc.l[d];
  if (with_dots) {
    AApplyTile<4> op;
    op.g = g, op.sp = s.sp, op.bc = s.bc, op.uc = uc, op.uniform = uniform;
    for (int c = 0; c < 3; ++c) op.a[c] = (a.c[c] == x.c[c]) ? nullptr : a.c[c], op.y[c] = y.c[c], op.wout[c] = nullptr;
    tma_launch(s.ex, op, fields, g.px, g.py, g.nzl + 2, g.nx, g.ny, kbeg, kend, carry, g.t[0].per != 0);
  } else {
    AApplyTile<0> op;
    op.g = g, op.sp = s.sp, op.bc = s.bc, op.uc = uc, op.uniform = uniform;
    for (int c = 0; c < 3; ++c) op.a[c] = nullptr, op.y[c] = y.c[c], op.wout[c] = nullptr;
    tma_launch(s.ex, op, fields, g.px, g.py, g.nzl + 2, g.nx, g.ny, kbeg, kend, nullptr, g.t[0].per != 0);
  }
  (void)ncar;
}

// plain (no reduction) versions of the wall adapters
template <class F>
struct PlaneAtPlain {
  F   f;
  int kl;
  FL_HD void operator()(int i, int j, int) const { f(i, j, kl); }
};
template <class F>
struct XWallAtPlain {
  F   f;
  int i, kbeg;
  FL_HD void operator()(int a, int b, int) const { f(i, a, kbeg + b); }
};

// velocity block of the coupled operator: y = A x + (dt/rho) G p, w = x + (dt/rho) G p
void coupled_cells_tma(Solver &s, const V3 &x, const double *p, const V3 &y, const V3 &w, bool keep_w)
{
  const Geom &g = s.gh.g;
  // 1. w = (dt/rho) G p everywhere (the tile kernel reads it as a per-cell operand and, with keep_w, overwrites it)
  GradCells<3> gc;
  gc.g = g, gc.dtrho = s.sp.dtrho, gc.p = p, gc.w = w;
  const Box all = {g.nx, g.ny, g.nzl};
  for_box<2, 3>(s.ex, all, gc); // 80 registers: three CTAs per SM (the streaming kernels follow occupancy, profiles/r04)
  // 2. wall planes and wall columns with the direct-load functor (it forms the gradient itself)
  const bool wl = g.t[2].wall_lo && !g.t[2].per, wh = g.t[2].wall_hi && !g.t[2].per;
  CoupledCells<3> f;
  f.g = g, f.sp = s.sp, f.bc = s.bc, f.x = CV3(x), f.v0 = CV3(s.v0), f.U0 = CV3(s.U0), f.p = p, f.y = y, f.w = keep_w ? w : V3{{nullptr, nullptr, nullptr}};
  const Box plane_box = {g.nx, g.ny, 1};
  const int kbeg = wl ? 1 : 0;
  int       kend = wh ? g.nzl - 1 : g.nzl;
  if (kend < kbeg) kend = kbeg;
  if (wl) {
    PlaneAtPlain<CoupledCells<3>> pf = {f, 0};
    for_box(s.ex, plane_box, pf);
  }
  if (wh && g.nzl - 1 >= kbeg) {
    PlaneAtPlain<CoupledCells<3>> pf = {f, g.nzl - 1};
    for_box(s.ex, plane_box, pf);
  }
  if (kend > kbeg && !g.t[0].per) {
    const Box wall_box = {g.ny, kend - kbeg, 1};
    for (int side = 0; side < 2; ++side) {
      XWallAtPlain<CoupledCells<3>> xf = {f, side ? g.nx - 1 : 0, kbeg};
      for_box(s.ex, wall_box, xf);
    }
  }
  // 3. everything else from TMA tiles
  const double *fields[9] = {x.c[0], x.c[1], x.c[2], s.v0.c[0], s.v0.c[1], s.v0.c[2], s.U0.c[0], s.U0.c[1], s.U0.c[2]};
  UniCoef       uc;
  static const bool no_uni = getenv("FLUCA_B200_NO_UNIFORM") != nullptr;
  uc.l4 = 0.;
  for (int d = 0; d < 3; ++d) uc.q[d] = 0.25 / g.t[d].uh, uc.l[d] = 1. / (g.t[d].uh * g.t[d].uh), uc.l4 += 4. * uc.l[d];
  AApplyTile<0> op;
  op.g = g, op.sp = s.sp, op.bc = s.bc, op.uc = uc, op.uniform = (!no_uni && g.t[0].uni && g.t[1].uni && g.t[2].uni) ? 1 : 0;
  for (int c = 0; c < 3; ++c) op.a[c] = nullptr, op.y[c] = y.c[c], op.wout[c] = w.c[c];
  op.keep_w = keep_w ? 1 : 0;
  tma_launch(s.ex, op, fields, g.px, g.py, g.nzl + 2, g.nx, g.ny, kbeg, kend, nullptr, g.t[0].per != 0);
}

// ------------------------------------------------------------------ Poisson operator from shared-memory tiles
// No plane needs special treatment: the wall rows of P use the cells (c, c+1) / (c-1, c) only.
template <int NRED>
struct PoissonTile : TileOpDefaults {
  static const int NIN = 1, NR = NRED, MINB = 4, STAGES = 8, PLANES = FL_TILE_PLANES;
  Geom             g;
  const double    *a; // dot partner; nullptr: p itself
  double          *out;
  int              uniform; // all directions uniform: interior warps use the constant row cd[d] (2 p - p- - p+)
  double           cd[3];
  struct Regs {
    double a;
  };
  __device__ int flags(int i, int j) const { return (uniform && (g.t[0].per || (i > 0 && i < g.nx - 1)) && j > 0 && j < g.ny - 1) ? 1 : 0; }
  __device__ void prefetch(Regs &rg, int off, int kl) const
  {
    if (NRED > 0 && a) rg.a = a[off];
  }
  __device__ void cell(const TileView &tv, const Regs &rg, int fl, int i, int j, int kl, int off, double *acc) const
  {
    const int    lc = tv.lc, ig[3] = {i, j, g.k0 + kl};
    const double pc = tv.p0[lc];
    const double pm[3] = {tv.p0[lc - 1], tv.p0[lc - TLX], tv.pm[lc]}, pp[3] = {tv.p0[lc + 1], tv.p0[lc + TLX], tv.pp[lc]};
    const bool   inter = fl && (g.t[2].per || (ig[2] > 0 && ig[2] < g.t[2].n - 1));
    double       v;
    if (__all_sync(__activemask(), inter)) v = cd[0] * (2. * pc - pm[0] - pp[0]) + cd[1] * (2. * pc - pm[1] - pp[1]) + cd[2] * (2. * pc - pm[2] - pp[2]);
    else v = poisson_row<3>(g, ig, pc, pm, pp);
    out[off] = v;
    if (NRED > 0) acc[0] += (a ? rg.a : pc) * v;
  }
};

// out = P p ; <a, out> is left in ex.d_result
void poisson_apply_dot_tma(Solver &s, const double *pin, double *out, const double *a)
{
  const Geom    &g = s.gh.g;
  PoissonTile<1> op;
  op.g = g, op.a = (a == pin) ? nullptr : a, op.out = out;
  static const bool no_uni = getenv("FLUCA_B200_NO_UNIFORM") != nullptr;
  op.uniform = (!no_uni && g.t[0].uni && g.t[1].uni && g.t[2].uni) ? 1 : 0;
  {
    const double hu[3] = {g.t[0].uh, g.t[1].uh, g.t[2].uh};
    for (int d = 0; d < 3; ++d) op.cd[d] = (hu[0] * hu[1] * hu[2] / hu[d]) / hu[d];
  }
  const double *fields[1] = {pin};
  tma_launch(s.ex, op, fields, g.px, g.py, g.nzl + 2, g.nx, g.ny, 0, g.nzl, nullptr, g.t[0].per != 0);
}

} // namespace fluca
#endif // !FLUCA_HOSTEMU
