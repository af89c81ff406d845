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
  return !off && s.dim == 3 && !g.t[0].per && !g.t[1].per && g.nx >= TMX && g.ny >= TMY;
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
    const bool    lo = BND && d < 2 && ig == 0, hi = BND && d < 2 && ig == T.n - 1;
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
  static const int  NIN = 9, NR = NRED, MINB = 2, STAGES = 4;
  static const bool SIDE = true;
  // side work: the cells whose rows reach two cells inward (one-sided wall rows) and would drag a whole warp down the
  // table-driven path -- the two x-wall columns of the tiled planes and the planes at a physical z wall.  They used to be
  // four extra launches (17 % of the operator's time at 256^3 for < 1 % of the cells, profiles/r02n); now they are extra CTAs
  // of the SAME launch, scheduled first, whose scattered loads run under the tile stream.
  CV3              xg, v0g, U0g; // global pointers of the tile fields, for the side CTAs
  const double    *pg = nullptr; // pressure of the coupled velocity block (the side cells form their own gradient)
  int              zwall[2] = {-1, -1}; // local planes at a physical z wall (-1: none)
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
  __device__ int flags(int i, int j) const { return ((i > 0 && i < g.nx - 1 && j > 0 && j < g.ny - 1) ? 1 : 0) | ((i == 0 || i == g.nx - 1) ? 2 : 0); }
  __host__ int side_blocks(const TmaGrid &tg) const
  {
    const int gy = (g.ny + 31) / 32, gk = tg.kend > tg.kbeg ? (tg.kend - tg.kbeg + 7) / 8 : 0;
    const int gp = ((g.nx + 31) / 32) * ((g.ny + 7) / 8);
    return 2 * gy * gk + ((zwall[0] >= 0) + (zwall[1] >= 0)) * gp;
  }
  __device__ void side_cell(int i, int j, int kl, double *acc) const
  {
    double r[3];
    a_apply_cell<3>(g, sp, bc, xg, v0g, U0g, i, j, kl, r);
    const int c = g.idx(i, j, kl);
    if (NRED == 0 && wout[0]) {
      Nbr<3> nb;
      nbr<3>(g, i, j, kl, nb);
      double gp[3];
      grad_cell<3>(g, pg, nb, gp);
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        y[q][c] = r[q] + sp.dtrho * gp[q];
        if (keep_w) wout[q][c] = xg.c[q][c] + sp.dtrho * gp[q];
      }
      return;
    }
    double d0 = 0., d1 = 0.;
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      y[q][c] = r[q];
      if (NRED > 0) {
        const double av = a[q] ? a[q][c] : xg.c[q][c];
        d0 += av * r[q];
        d1 += r[q] * r[q];
      }
    }
    if (NRED > 0) acc[0] += d0, acc[1] += d1;
  }
  __device__ void side(const TmaGrid &tg, int b, int tx, int ty, double *acc) const
  {
    const int gy = (g.ny + 31) / 32, gk = tg.kend > tg.kbeg ? (tg.kend - tg.kbeg + 7) / 8 : 0;
    const int nxw = 2 * gy * gk;
    if (b < nxw) {
      // an x-wall column: the CTA covers 32 rows j x 8 planes, so that a warp holds 32 wall cells
      const int sd = b / (gy * gk), r = b % (gy * gk);
      const int j = (r % gy) * 32 + tx, kl = tg.kbeg + (r / gy) * 8 + ty;
      if (j < g.ny && kl < tg.kend) side_cell(sd ? g.nx - 1 : 0, j, kl, acc);
      return;
    }
    b -= nxw;
    const int gx = (g.nx + 31) / 32, gp = gx * ((g.ny + 7) / 8);
    const int which = b / gp, r = b % gp;
    const int kl = (which == 0 && zwall[0] >= 0) ? zwall[0] : zwall[1];
    const int i = (r % gx) * 32 + tx, j = (r / gx) * 8 + ty;
    if (i < g.nx && j < g.ny) side_cell(i, j, kl, acc);
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
    double    d0 = 0., d1 = 0.;
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
        const double av = a[q] ? rg.a[q] : tv.p0[q * TILE_STRIDE + tv.lc];
        d0 += av * r[q];
        d1 += r[q] * r[q];
      }
    }
    if (NRED > 0) acc[0] += d0, acc[1] += d1;
  }
};

static UniCoef uni_coef(const Geom &g)
{
  UniCoef uc;
  uc.l4 = 0.;
  for (int d = 0; d < 3; ++d) uc.q[d] = 0.25 / g.t[d].uh, uc.l[d] = 1. / (g.t[d].uh * g.t[d].uh), uc.l4 += 4. * uc.l[d];
  return uc;
}

template <class Op>
static void a_tile_common(Solver &s, Op &op, const V3 &x, int &kbeg, int &kend)
{
  const Geom &g  = s.gh.g;
  const bool  wl = g.t[2].wall_lo && !g.t[2].per, wh = g.t[2].wall_hi && !g.t[2].per;
  static const bool no_uni = getenv("FLUCA_B200_NO_UNIFORM") != nullptr;
  op.g = g, op.sp = s.sp, op.bc = s.bc, op.uc = uni_coef(g);
  op.uniform = (!no_uni && g.t[0].uni && g.t[1].uni && g.t[2].uni) ? 1 : 0;
  op.xg = CV3(x), op.v0g = CV3(s.v0), op.U0g = CV3(s.U0);
  // planes that touch a physical z wall need the one-sided 4-point rows (cells k +- 2): side CTAs
  kbeg = wl ? 1 : 0;
  kend = wh ? g.nzl - 1 : g.nzl;
  if (kend < kbeg) kend = kbeg;
  op.zwall[0] = wl ? 0 : -1;
  op.zwall[1] = (wh && g.nzl - 1 >= kbeg) ? g.nzl - 1 : -1;
}

// y = A x ; out[0] = <a, y>, out[1] = <y, y> are left in ex.d_result (reduce_finish reads them).  ONE launch: tile CTAs for
// the planes and columns away from the walls, side CTAs for the rest.
void a_apply_dots_tma(Solver &s, const V3 &x, const V3 &y, const V3 &a, bool with_dots)
{
  const Geom   &g = s.gh.g;
  KGroup        grp(s.ex, KT_MOMENTUM_APPLY);
  const double *fields[9] = {x.c[0], x.c[1], x.c[2], s.v0.c[0], s.v0.c[1], s.v0.c[2], s.U0.c[0], s.U0.c[1], s.U0.c[2]};
  int           kbeg, kend;
  if (with_dots) {
    AApplyTile<2> op;
    a_tile_common(s, op, x, kbeg, kend);
    for (int c = 0; c < 3; ++c) op.a[c] = (a.c[c] == x.c[c]) ? nullptr : a.c[c], op.y[c] = y.c[c], op.wout[c] = nullptr;
    tma_launch(s.ex, op, fields, g.px, g.py, g.nzl + 2, g.nx, g.ny, kbeg, kend, nullptr);
  } else {
    AApplyTile<0> op;
    a_tile_common(s, op, x, kbeg, kend);
    for (int c = 0; c < 3; ++c) op.a[c] = nullptr, op.y[c] = y.c[c], op.wout[c] = nullptr;
    tma_launch(s.ex, op, fields, g.px, g.py, g.nzl + 2, g.nx, g.ny, kbeg, kend, nullptr);
  }
}

// velocity block of the coupled operator: y = A x + (dt/rho) G p, w = x + (dt/rho) G p
void coupled_cells_tma(Solver &s, const V3 &x, const double *p, const V3 &y, const V3 &w, bool keep_w)
{
  const Geom &g = s.gh.g;
  // 1. w = (dt/rho) G p everywhere (the tile kernel reads it as a per-cell operand and, with keep_w, overwrites it)
  GradCells<3> gc;
  gc.g = g, gc.dtrho = s.sp.dtrho, gc.p = p, gc.w = w;
  const Box all = {g.nx, g.ny, g.nzl};
  for_box<2>(s.ex, all, gc);
  // 2. one launch: tiles + side CTAs (the side cells form the gradient themselves)
  const double *fields[9] = {x.c[0], x.c[1], x.c[2], s.v0.c[0], s.v0.c[1], s.v0.c[2], s.U0.c[0], s.U0.c[1], s.U0.c[2]};
  AApplyTile<0> op;
  int           kbeg, kend;
  a_tile_common(s, op, x, kbeg, kend);
  for (int c = 0; c < 3; ++c) op.a[c] = nullptr, op.y[c] = y.c[c], op.wout[c] = w.c[c];
  op.keep_w = keep_w ? 1 : 0, op.pg = p;
  tma_launch(s.ex, op, fields, g.px, g.py, g.nzl + 2, g.nx, g.ny, kbeg, kend, nullptr);
}

// ------------------------------------------------------------------ Poisson operator from shared-memory tiles
// No plane needs special treatment: the wall rows of P use the cells (c, c+1) / (c-1, c) only.
template <int NRED>
struct PoissonTile : TileOpDefaults {
  static const int NIN = 1, NR = NRED, MINB = 4, STAGES = 8;
  Geom             g;
  const double    *a; // dot partner; nullptr: p itself
  double          *out;
  int              uniform; // all directions uniform: interior warps use the constant row cd[d] (2 p - p- - p+)
  double           cd[3];
  struct Regs {
    double a;
  };
  __device__ int flags(int i, int j) const { return (uniform && i > 0 && i < g.nx - 1 && j > 0 && j < g.ny - 1) ? 1 : 0; }
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
  tma_launch(s.ex, op, fields, g.px, g.py, g.nzl + 2, g.nx, g.ny, 0, g.nzl, nullptr);
}

} // namespace fluca
#endif // !FLUCA_HOSTEMU
