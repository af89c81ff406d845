// fd.cu -- the FlucaFD operator family behind the C ABI: the stencil layer (host C++) and, at the end
// of the file, the matrix-free device apply generated from it.
//
// Stencil layer: what the reference computes in FlucaFDSetUp / FlucaFDGetStencil for its five operator types -- the part
// every matrix-free apply of a composed operator is generated from.  Restated from
//   fluca/src/fd/interface/fdapply.c:22-45         FlucaFDGetStencilRaw, FlucaFDGetStencil
//   fluca/src/fd/utils/fdutils.c:56-100            coordinates beyond the local grid, ghost corners, Gaussian elimination
//   fluca/src/fd/utils/fdutils.c:102-489           stencil accumulation, off-grid removal per boundary condition, zero removal
//   fluca/src/fd/impls/{derivative,sum,scale,composition,secondordertvd}/*.c
// for one rank.  tests/test_fd_stencils.py drives it through the C ABI on the command lines of the reference's own tests
// (fluca/tests/fd/ex*.c) and compares with their stored outputs, character for character.
#include "../../include/fluca_b200.h"
#include "exec.h"
#include <climits>
#include <memory>

namespace fluca {
namespace fd {

static const int    MAXS           = FLUCA_B200_FD_MAX_STENCIL; // FLUCAFD_MAX_STENCIL_SIZE, flucafdimpl.h:8
static const double ZERO_PIVOT_TOL = 1e-14;                     // flucafdimpl.h:9
static const double COEFF_ATOL     = 1e-10;                     // flucafdimpl.h:10
static const double COEFF_RTOL     = 1e-8;                      // flucafdimpl.h:11
static const int    CONSTANT       = -7;                        // FLUCAFD_CONSTANT, flucafd.h:52

// DMStagStencilLocation values (petscdmstag.h): only ELEMENT, LEFT, DOWN, BACK and their combinations are inputs / outputs
enum { LOC_BACK_DOWN_LEFT = 1, LOC_BACK_DOWN = 2, LOC_BACK_LEFT = 4, LOC_BACK = 5, LOC_DOWN_LEFT = 10, LOC_DOWN = 11, LOC_LEFT = 13, LOC_ELEMENT = 14 };

static bool valid_location(int loc)
{
  switch (loc) { // FlucaFDValidateStencilLocation_Internal, fdutils.c:16-34
  case LOC_ELEMENT: case LOC_LEFT: case LOC_DOWN: case LOC_BACK: case LOC_DOWN_LEFT: case LOC_BACK_LEFT: case LOC_BACK_DOWN: case LOC_BACK_DOWN_LEFT: return true;
  default: return false;
  }
}
// bit d set: the location lies on the LEFT (0) / DOWN (1) / BACK (2) face
static int face_bits(int loc)
{
  switch (loc) {
  case LOC_LEFT: return 1;
  case LOC_DOWN: return 2;
  case LOC_BACK: return 4;
  case LOC_DOWN_LEFT: return 3;
  case LOC_BACK_LEFT: return 5;
  case LOC_BACK_DOWN: return 6;
  case LOC_BACK_DOWN_LEFT: return 7;
  default: return 0;
  }
}
static int location_of_bits(int b)
{
  static const int t[8] = {LOC_ELEMENT, LOC_LEFT, LOC_DOWN, LOC_DOWN_LEFT, LOC_BACK, LOC_BACK_LEFT, LOC_BACK_DOWN, LOC_BACK_DOWN_LEFT};
  return t[b & 7];
}
static bool use_face(int loc, int d) { return (face_bits(loc) >> d) & 1; }                          // fdutils.c:36-54
static int  boundary_location(int loc, int d) { return location_of_bits(face_bits(loc) | (1 << d)); } // fdutils.c:198-252

// FlucaFDSolveLinearSystem_Internal, fdutils.c:79-100: Gaussian elimination without pivoting (A is n x n, row-major)
static void solve(int n, double *A, double *b, double *x)
{
  for (int k = 0; k < n - 1; ++k) {
    if (!(std::fabs(A[k * n + k]) > ZERO_PIVOT_TOL)) throw Error(FL_ERR_INTERNAL, "zero pivot during forward elimination");
    for (int i = k + 1; i < n; ++i) {
      const double f = A[i * n + k] / A[k * n + k];
      for (int j = k; j < n; ++j) A[i * n + j] -= f * A[k * n + j];
      b[i] -= f * b[k];
    }
  }
  for (int i = n - 1; i >= 0; --i) {
    if (!(std::fabs(A[i * n + i]) > ZERO_PIVOT_TOL)) throw Error(FL_ERR_INTERNAL, "zero pivot during back substitution");
    double s = 0.;
    for (int j = i + 1; j < n; ++j) s += A[i * n + j] * x[j];
    x[i] = (b[i] - s) / A[i * n + i];
  }
}
static double ipow(double h, int r)
{
  double p = 1.;
  for (int q = 0; q < r; ++q) p *= h;
  return p;
}

struct Col {
  int i, j, k, loc, c;
  int  idx(int d) const { return d == 0 ? i : (d == 1 ? j : k); }
  void set(int d, int v) { (d == 0 ? i : (d == 1 ? j : k)) = v; }
  bool operator==(const Col &o) const { return i == o.i && j == o.j && k == o.k && loc == o.loc && c == o.c; }
};
struct Stencil {
  int    n = 0;
  Col    col[MAXS];
  double v[MAXS];
  void add(const Col &c, double w) // FlucaFDAddStencilPoint_Internal, fdutils.c:102-124
  {
    for (int q = 0; q < n; ++q)
      if (col[q] == c) {
        v[q] += w;
        return;
      }
    if (n >= MAXS) throw Error(FL_ERR_ARG, "Resulting stencil is too large");
    col[n] = c, v[n] = w, ++n;
  }
  void erase(int q)
  {
    for (int r = q; r < n - 1; ++r) col[r] = col[r + 1], v[r] = v[r + 1];
    --n;
  }
  void remove_zero() // FlucaFDRemoveZeroStencilPoints_Internal, fdutils.c:465-489
  {
    double s = 0.;
    for (int q = 0; q < n; ++q) s += std::fabs(v[q]);
    int m = 0;
    for (int q = 0; q < n; ++q)
      if (!(std::fabs(v[q]) < COEFF_ATOL || std::fabs(v[q] / s) < COEFF_RTOL)) col[m] = col[q], v[m] = v[q], ++m;
    n = m;
  }
};

struct Term { // FlucaFDTermLink
  int deriv[3] = {-1, -1, -1}, accu[3] = {INT_MAX, INT_MAX, INT_MAX};
  int input_loc = LOC_ELEMENT, input_c = 0;
  bool operator==(const Term &o) const
  {
    for (int d = 0; d < 3; ++d)
      if (deriv[d] != o.deriv[d] || accu[d] != o.accu[d]) return false;
    return input_loc == o.input_loc && input_c == o.input_c;
  }
};
static void merge_terms(std::vector<Term> &dst, const std::vector<Term> &src)
{
  for (const Term &t : src) {
    bool found = false;
    for (const Term &u : dst) found = found || (u == t);
    if (!found) dst.push_back(t);
  }
}

// the DMStag facts FlucaFDSetUp reads (fdbasic.c:165-186), one rank: sizes, periodicity, stencil width, product coordinates
struct Grid {
  int                 dim = 1, N[3] = {1, 1, 1}, per[3] = {0, 0, 0}, sw = 1;
  std::vector<double> xf[3], xc[3]; // N + 1 faces, N centres
  int                 refs = 1;
  // product coordinate array entry, ghost elements of a periodic direction included (they continue the grid by its period)
  double array_coord(int d, int idx, bool face) const
  {
    const int    n = N[d];
    const double L = xf[d][n] - xf[d][0];
    int          w = idx, shift = 0;
    const int    top = face ? n : n - 1; // last stored index
    while (w < 0) w += n, --shift;
    while (w > top) w -= n, ++shift;
    return (face ? xf[d][w] : xc[d][w]) + shift * L;
  }
  void ghost_corners(int d, bool face, int &gxs, int &gxm, int &gxe) const // FlucaFDGetGhostCorners_Internal, fdutils.c:67-77
  {
    gxs = per[d] ? -sw : 0;
    gxm = N[d] + (per[d] ? 2 * sw : 0);
    gxe = (face && !per[d]) ? 1 : 0;
  }
  // FlucaFDGetCoordinate_Internal, fdutils.c:56-65
  double coordinate(int d, int idx, bool face, int x, int n, double hp, double hn) const
  {
    if (x <= idx && idx < x + n) return array_coord(d, idx, face);
    if (idx < x) return array_coord(d, x, face) - (x - idx) * hp;
    return array_coord(d, x + n - 1, face) + (idx - (x + n - 1)) * hn;
  }
  void end_spacings(int d, int gxs, int gxm, double &hp, double &hn) const // derivative.c:79-82, fdutils.c:300-303
  {
    const int first = gxs, last = gxs + gxm - (per[d] ? 1 : 0);
    hp = array_coord(d, first + 1, true) - array_coord(d, first, true);
    hn = array_coord(d, last, true) - array_coord(d, last - 1, true);
  }
};

struct Op {
  Grid             *g = nullptr;
  int               input_loc = LOC_ELEMENT, input_c = 0, output_loc = LOC_ELEMENT, output_c = 0;
  int               bc_type[6] = {0, 0, 0, 0, 0, 0};
  double            bc_value[6] = {0, 0, 0, 0, 0, 0};
  std::vector<Term> terms;
  bool              setupcalled = false;
  int               refs = 1;
  virtual ~Op() { }
  virtual void setup() = 0;
  virtual void raw(int i, int j, int k, Stencil &st) = 0;
  void stencil_raw(int i, int j, int k, Stencil &st) // FlucaFDGetStencilRaw, fdapply.c:22-32
  {
    if (!setupcalled) throw Error(FL_ERR_ARG, "FlucaFD not setup");
    st.n = 0;
    raw(i, j, k, st);
    st.remove_zero();
  }
  void stencil(int i, int j, int k, Stencil &st); // FlucaFDGetStencil, fdapply.c:34-45
};

// IsOffGrid_Private, fdutils.c:126-168
static bool off_grid(const Op &fd, const Col &c, int &dir, bool &low)
{
  if (c.c < 0) return false;
  for (int d = 0; d < fd.g->dim; ++d) {
    int gxs, gxm, gxe;
    fd.g->ghost_corners(d, use_face(c.loc, d), gxs, gxm, gxe);
    if (c.idx(d) < gxs) {
      dir = d, low = true;
      return true;
    }
    if (c.idx(d) >= gxs + gxm + gxe) {
      dir = d, low = false;
      return true;
    }
  }
  return false;
}

// FlucaFDRemoveOffGridPoints_Internal, fdutils.c:254-463
static void remove_off_grid(const Op &fd, Stencil &st)
{
  const Grid &g = *fd.g;
  int         iter = 0;
  for (; iter < 100; ++iter) {
    int  off = -1, d = 0;
    bool low = false;
    for (int q = 0; q < st.n; ++q)
      if (off_grid(fd, st.col[q], d, low)) {
        off = q;
        break;
      }
    if (off < 0) break;
    const Col    oc = st.col[off];
    const double ov = st.v[off];
    const bool   per = g.per[d] != 0;
    const int    bc  = (low && !per) ? fd.bc_type[2 * d] : ((!low && !per) ? fd.bc_type[2 * d + 1] : FLUCA_B200_FD_BC_NONE);
    const bool   face = use_face(oc.loc, d);
    // GetStencilSizeForOffGridPoint_Private, fdutils.c:170-196
    int size = INT_MAX;
    for (const Term &t : fd.terms)
      if (t.deriv[d] != -1 && t.accu[d] != INT_MAX && t.input_loc == oc.loc && t.input_c == oc.c) size = std::min(size, t.deriv[d] + t.accu[d]);
    if (size == INT_MAX) throw Error(FL_ERR_ARG, "Cannot find a term from the given stencil point");
    if (size < 1) size = 1;
    if (size > MAXS) throw Error(FL_ERR_ARG, "Stencil size exceeds maximum");
    int gxs, gxm, gxe;
    g.ghost_corners(d, face, gxs, gxm, gxe);
    double hp, hn;
    g.end_spacings(d, gxs, gxm, hp, hn);
    const double off_x = g.coordinate(d, oc.idx(d), face, gxs, gxm + gxe, hp, hn);
    st.erase(off);
    double xs[MAXS], w[MAXS], A[MAXS * MAXS], b[MAXS];
    if (bc == FLUCA_B200_FD_BC_NONE) { // :326-356
      const int start = low ? gxs : gxs + gxm + gxe - size;
      for (int m = 0; m < size; ++m) xs[m] = g.coordinate(d, start + m, face, gxs, gxm + gxe, hp, hn);
      for (int r = 0; r < size; ++r) {
        for (int m = 0; m < size; ++m) A[r * size + m] = ipow(xs[m] - off_x, r);
        b[r] = r == 0 ? 1. : 0.;
      }
      solve(size, A, b, w);
      for (int m = 0; m < size; ++m) {
        Col c = oc;
        c.set(d, start + m);
        st.add(c, ov * w[m]);
      }
    } else if (bc == FLUCA_B200_FD_BC_DIRICHLET || bc == FLUCA_B200_FD_BC_NEUMANN) {
      int start = low ? gxs : gxs + gxm + gxe - (size - 1);
      if (bc == FLUCA_B200_FD_BC_DIRICHLET && face) start += low ? 1 : -1; // remove duplicate (:361-365)
      const int    bnd = low ? 0 : g.N[d];
      const double bx  = g.coordinate(d, bnd, true, gxs, gxm + gxe, hp, hn);
      Col          marker;
      marker.i = d == 0 ? bnd : oc.i, marker.j = d == 1 ? bnd : oc.j, marker.k = d == 2 ? bnd : oc.k;
      marker.loc = boundary_location(oc.loc, d);
      marker.c   = -(2 * d + (low ? 1 : 2));
      if (bc == FLUCA_B200_FD_BC_DIRICHLET) { // value from the boundary value and size - 1 on-grid points (:358-404)
        xs[0] = bx;
        for (int m = 0; m < size - 1; ++m) xs[m + 1] = g.coordinate(d, start + m, face, gxs, gxm + gxe, hp, hn);
        for (int r = 0; r < size; ++r) {
          for (int m = 0; m < size; ++m) A[r * size + m] = ipow(xs[m] - off_x, r);
          b[r] = r == 0 ? 1. : 0.;
        }
        solve(size, A, b, w);
        st.add(marker, ov * w[0]);
        for (int m = 0; m < size - 1; ++m) {
          Col c = oc;
          c.set(d, start + m);
          st.add(c, ov * w[m + 1]);
        }
      } else { // first derivative on the boundary from the off-grid point and size - 1 on-grid points, solved for the former (:406-452)
        xs[0] = off_x;
        for (int m = 0; m < size - 1; ++m) xs[m + 1] = g.coordinate(d, start + m, face, gxs, gxm + gxe, hp, hn);
        for (int r = 0; r < size; ++r) {
          for (int m = 0; m < size; ++m) A[r * size + m] = ipow(xs[m] - bx, r);
          b[r] = r == 1 ? 1. : 0.;
        }
        solve(size, A, b, w);
        if (std::fabs(w[0]) < COEFF_ATOL) throw Error(FL_ERR_ARG, "Neumann BC coefficient for off-grid point is too small");
        st.add(marker, ov / w[0]);
        for (int m = 0; m < size - 1; ++m) {
          Col c = oc;
          c.set(d, start + m);
          st.add(c, -ov * w[m + 1] / w[0]);
        }
      }
    } else throw Error(FL_ERR_ARG, "Unsupported boundary condition type");
  }
  if (iter >= 100) throw Error(FL_ERR_INTERNAL, "Failed to remove all off-grid points");
  st.remove_zero();
}

void Op::stencil(int i, int j, int k, Stencil &st)
{
  stencil_raw(i, j, k, st);
  remove_off_grid(*this, st);
  st.remove_zero();
}

// ------------------------------------------------------------------ FLUCAFDDERIVATIVE, derivative.c:16-150
struct Derivative : Op {
  int    dir = 0, deriv_order = 1, accu_order = 1;
  int    size = 0, offset = 0, gxs = 0, gxm = 0, gxe = 0, v_start = 0, v_end = 0;
  bool   fin = false, fout = false;
  double hp = 0., hn = 0.;
  std::vector<std::vector<double>> w; // weights for indices v_start - 1 .. v_end (the ends are v_prev / v_next)
  void setup() override
  {
    if (dir >= g->dim) throw Error(FL_ERR_ARG, "Cannot compute derivative in that direction on this DM");
    if (deriv_order < 0 || accu_order < 1) throw Error(FL_ERR_ARG, "Order of derivative must be non-negative and order of accuracy positive");
    fin = use_face(input_loc, dir), fout = use_face(output_loc, dir);
    bool valid = fin != fout;
    for (int d = 0; d < g->dim; ++d)
      if (d != dir && use_face(input_loc, d) != use_face(output_loc, d)) valid = false;
    if (!(input_loc == output_loc || valid)) throw Error(FL_ERR_ARG, "Cannot compute derivative between these stencil locations");
    size = deriv_order + accu_order;
    if (size > MAXS) throw Error(FL_ERR_ARG, "Required stencil size exceeds maximum");
    offset = -((size - 1) / 2);
    if (!fin && fout) offset -= 1;
    g->ghost_corners(dir, fin, gxs, gxm, gxe);
    v_start = gxs - (offset + size - 1);
    v_end   = gxs + gxm + gxe - offset;
    g->end_spacings(dir, gxs, gxm, hp, hn);
    w.assign(v_end - v_start + 2, std::vector<double>());
    for (int i = v_start - 1; i < v_end + 1; ++i) {
      double       A[MAXS * MAXS], b[MAXS];
      const double out_x = g->coordinate(dir, i, fout, gxs, gxm, hp, hn);
      for (int c = 0; c < size; ++c) {
        const double h = g->coordinate(dir, i + offset + c, fin, gxs, gxm + gxe, hp, hn) - out_x;
        for (int r = 0; r < size; ++r) A[r * size + c] = ipow(h, r);
      }
      double fact = 1.;
      for (int o = 1; o <= deriv_order; ++o) fact *= o;
      for (int c = 0; c < size; ++c) b[c] = c == deriv_order ? fact : 0.;
      std::vector<double> &x = w[i - (v_start - 1)];
      x.resize(size);
      solve(size, A, b, x.data());
    }
    Term t;
    t.deriv[dir] = deriv_order, t.accu[dir] = accu_order, t.input_loc = input_loc, t.input_c = input_c;
    terms.assign(1, t);
  }
  void raw(int i, int j, int k, Stencil &st) override
  {
    const int idx = dir == 0 ? i : (dir == 1 ? j : k);
    const int q   = std::min(std::max(idx, v_start - 1), v_end) - (v_start - 1);
    for (int c = 0; c < size; ++c) {
      Col p = {i, j, k, input_loc, input_c};
      p.set(dir, idx + offset + c);
      st.col[st.n] = p, st.v[st.n] = w[q][c], ++st.n;
    }
  }
};

// ------------------------------------------------------------------ FLUCAFDSUM, sum.c:3-52
struct Sum : Op {
  std::vector<Op *> ops;
  ~Sum() override;
  void setup() override
  {
    if (ops.empty()) throw Error(FL_ERR_ARG, "No operands set");
    if (input_loc != output_loc || input_c != output_c) throw Error(FL_ERR_ARG, "Cannot change location / component");
    for (Op *o : ops)
      if (o->output_loc != output_loc || o->output_c != output_c) throw Error(FL_ERR_ARG, "All operands must have the same output stencil location and component");
    terms.clear();
    for (Op *o : ops) merge_terms(terms, o->terms);
  }
  void raw(int i, int j, int k, Stencil &st) override
  {
    for (Op *o : ops) {
      Stencil s;
      o->stencil_raw(i, j, k, s);
      for (int q = 0; q < s.n; ++q) st.add(s.col[q], s.v[q]);
    }
  }
};

// compact host field of one location / component: [nz + ez][ny + ey][nx + ex], e = 1 on a non-periodic face direction
struct HostField {
  std::vector<double> a;
  int                 ext[3] = {1, 1, 1};
  bool                set = false;
  void assign(const Grid &g, int loc, const double *src)
  {
    size_t n = 1;
    for (int d = 0; d < 3; ++d) {
      ext[d] = d < g.dim ? g.N[d] + ((use_face(loc, d) && !g.per[d]) ? 1 : 0) : 1;
      n *= (size_t)ext[d];
    }
    a.assign(src, src + n);
    set = true;
  }
  // periodic directions wrap (the reference reads its ghosted local array); beyond a non-periodic end that array holds 0
  double at(const Grid &g, int i, int j, int k) const
  {
    int p[3] = {i, j, k};
    for (int d = 0; d < g.dim; ++d) {
      if (g.per[d]) p[d] = ((p[d] % ext[d]) + ext[d]) % ext[d];
      else if (p[d] < 0 || p[d] >= ext[d]) return 0.;
    }
    return a[(size_t)p[0] + (size_t)ext[0] * ((size_t)(g.dim > 1 ? p[1] : 0) + (size_t)ext[1] * (size_t)(g.dim > 2 ? p[2] : 0))];
  }
};

// ------------------------------------------------------------------ FLUCAFDSCALE, scale.c:18-93
struct Scale : Op {
  Op       *operand = nullptr;
  bool      is_constant = true;
  double    constant = 1.;
  HostField vec;
  int       vec_loc = LOC_ELEMENT, vec_c = 0;
  ~Scale() override;
  void setup() override
  {
    if (!operand) throw Error(FL_ERR_ARG, "Operand not set");
    if (!(operand->output_c == input_c && input_c == output_c)) throw Error(FL_ERR_ARG, "Cannot change component");
    if (!(operand->output_loc == input_loc && input_loc == output_loc)) throw Error(FL_ERR_ARG, "Cannot change location");
    if (!is_constant) {
      if (!vec.set) throw Error(FL_ERR_ARG, "Neither constant nor vector scale specified");
      if (operand->output_loc != vec_loc) throw Error(FL_ERR_ARG, "Operand and vector must have the same location");
    }
    terms = operand->terms;
  }
  void raw(int i, int j, int k, Stencil &st) override
  {
    operand->stencil_raw(i, j, k, st);
    const double s = is_constant ? constant : vec.at(*g, i, j, k);
    for (int q = 0; q < st.n; ++q) st.v[q] *= s;
  }
};

// ------------------------------------------------------------------ FLUCAFDCOMPOSITION, composition.c:3-76
struct Composition : Op {
  Op *inner = nullptr, *outer = nullptr;
  ~Composition() override;
  void setup() override
  {
    if (!inner || !outer) throw Error(FL_ERR_ARG, "Inner / outer operator not set");
    if (inner->output_c != outer->input_c) throw Error(FL_ERR_ARG, "Inner output component must match outer input component");
    if (inner->output_loc != outer->input_loc) throw Error(FL_ERR_ARG, "Inner output location must match outer input location");
    terms.clear();
    for (const Term &ot : outer->terms)
      for (const Term &it : inner->terms) {
        Term t;
        for (int d = 0; d < 3; ++d) {
          if (it.deriv[d] == -1) t.deriv[d] = ot.deriv[d];
          else if (ot.deriv[d] == -1) t.deriv[d] = it.deriv[d];
          else t.deriv[d] = it.deriv[d] + ot.deriv[d];
          t.accu[d] = std::min(it.accu[d], ot.accu[d]);
        }
        t.input_loc = it.input_loc, t.input_c = it.input_c;
        merge_terms(terms, std::vector<Term>(1, t));
      }
  }
  void raw(int i, int j, int k, Stencil &st) override
  {
    Stencil so;
    outer->stencil_raw(i, j, k, so);
    for (int oc = 0; oc < so.n; ++oc) {
      if (so.col[oc].c < 0) { // constant or boundary marker of the outer operator passes through
        st.add(so.col[oc], so.v[oc]);
        continue;
      }
      Stencil si;
      inner->stencil_raw(so.col[oc].i, so.col[oc].j, so.col[oc].k, si);
      for (int ic = 0; ic < si.n; ++ic) st.add(si.col[ic], so.v[oc] * si.v[ic]);
    }
  }
};

// ------------------------------------------------------------------ FLUCAFDSECONDORDERTVD, secondordertvd.c:53-356
typedef double (*Limiter)(double);
static double lim_superbee(double r) { return std::max(0., std::max(std::min(2. * r, 1.), std::min(r, 2.))); }
static double lim_minmod(double r) { return std::max(0., std::min(r, 1.)); }
static double lim_mc(double r) { return std::max(0., std::min(std::min(2. * r, (1. + r) / 2.), 2.)); }
static double lim_vanleer(double r) { return (r + std::fabs(r)) / (1. + std::fabs(r)); }
static double lim_vanalbada(double r) { return r <= 0. ? 0. : (r * r + r) / (r * r + 1.); }
static double lim_barthjesperson(double r) { return r <= 0. ? 0. : (1. + r) / 2. * std::min(1., std::min(4. * r / (1. + r), 4. / (1. + r))); }
static double lim_venkatakrishnan(double r) { return r <= 0. ? 0. : (1. + r) / 2. * std::min(4. * r * (3. * r + 1.) / (11. * r * r + 4. * r + 1.), 4. * (r + 3.) / (r * r + 4. * r + 11.)); }
static double lim_koren(double r) { return std::max(0., std::min(std::min(2. * r, (1. + 2. * r) / 3.), 2.)); }
static double lim_upwind(double) { return 0.; }
static double lim_sou(double r) { return r; }
static double lim_quick(double r) { return (3. + r) / 4.; }
static Limiter find_limiter(const std::string &name) // FlucaFDLimiterRegisterAll, secondordertvd.c:19-36
{
  static const struct { const char *n; Limiter f; } tab[] = {{"superbee", lim_superbee}, {"minmod", lim_minmod}, {"mc", lim_mc}, {"vanleer", lim_vanleer}, {"vanalbada", lim_vanalbada}, {"barthjesperson", lim_barthjesperson}, {"venkatakrishnan", lim_venkatakrishnan}, {"koren", lim_koren}, {"upwind", lim_upwind}, {"sou", lim_sou}, {"quick", lim_quick}};
  for (const auto &e : tab)
    if (name == e.n) return e.f;
  throw Error(FL_ERR_ARG, "Unknown limiter type: " + name);
}

struct TVD : Op {
  int                         dir = 0;
  Limiter                     limiter = lim_superbee;
  HostField                   vel, phi; // face velocity (location of the output), element-centred current solution
  std::unique_ptr<Derivative> grad;
  void setup() override
  {
    static const int face_of[3] = {LOC_LEFT, LOC_DOWN, LOC_BACK};
    if (input_loc != LOC_ELEMENT) throw Error(FL_ERR_ARG, "Input location must be DMSTAG_ELEMENT for TVD interpolation");
    if (dir >= g->dim || output_loc != face_of[dir]) throw Error(FL_ERR_ARG, "Output location must match direction (LEFT for X, DOWN for Y, BACK for Z)");
    grad.reset(new Derivative()); // d phi / dx, element -> face, first order, this operator's boundary condition TYPES (:76-78)
    grad->g = g, grad->dir = dir, grad->deriv_order = 1, grad->accu_order = 1;
    grad->input_loc = input_loc, grad->input_c = input_c, grad->output_loc = output_loc, grad->output_c = 0;
    for (int b = 0; b < 6; ++b) grad->bc_type[b] = bc_type[b], grad->bc_value[b] = bc_value[b];
    grad->setup();
    grad->setupcalled = true;
    Term t;
    t.deriv[dir] = 0, t.accu[dir] = 2, t.input_loc = input_loc, t.input_c = input_c; // interpolation, second order (:131-139)
    terms.assign(1, t);
  }
  void alpha(int idx, double &plus, double &minus) const // :86-128
  {
    plus = minus = 0.5;
    if ((idx == 0 || idx == g->N[dir]) && !g->per[dir]) return;
    const double xf = g->array_coord(dir, idx, true), xl = g->array_coord(dir, idx - 1, false), xr = g->array_coord(dir, idx, false), dx = xr - xl;
    if (std::fabs(dx) > 1e-14) plus = (xf - xl) / dx, minus = (xr - xf) / dx;
  }
  double face_gradient(int i, int j, int k) // ComputeFaceCenteredGradient_Private, :150-185
  {
    Stencil s;
    grad->stencil(i, j, k, s);
    double r = 0.;
    for (int q = 0; q < s.n; ++q) {
      if (s.col[q].c >= 0) r += s.v[q] * phi.at(*g, s.col[q].i, s.col[q].j, s.col[q].k);
      else if (s.col[q].c >= -6) r += s.v[q] * bc_value[-s.col[q].c - 1]; // the CURRENT boundary values of this operator
      else throw Error(FL_ERR_ARG, "Unsupported stencil point");
    }
    return r;
  }
  void raw(int i, int j, int k, Stencil &st) override
  {
    if (!vel.set || !phi.set) throw Error(FL_ERR_ARG, "Velocity / current solution not set");
    int       p[3] = {i, j, k}, lo[3] = {i, j, k}, up[3] = {i, j, k};
    const int idx = p[dir];
    lo[dir] -= 1, up[dir] += 1;
    int gxs, gxm, gxe;
    g->ghost_corners(dir, true, gxs, gxm, gxe);
    if (!(gxs <= idx && idx < gxs + gxm + gxe)) throw Error(FL_ERR_ARG, "Face index out of range");
    const double u = vel.at(*g, i, j, k);
    const bool   at_prev = idx == 0 && !g->per[dir], at_next = idx == g->N[dir] && !g->per[dir];
    const Col    here = {p[0], p[1], p[2], input_loc, input_c}, below = {lo[0], lo[1], lo[2], input_loc, input_c}, cst = {0, 0, 0, LOC_ELEMENT, CONSTANT};
    double       ap, am;
    alpha(idx, ap, am);
    st.n = 2;
    if (u > 0) {
      if (at_prev) {
        st.col[0] = below, st.v[0] = 0.5, st.col[1] = here, st.v[1] = 0.5;
        return;
      }
      const double gfu = face_gradient(lo[0], lo[1], lo[2]), gfc = face_gradient(p[0], p[1], p[2]);
      const double psi = limiter(std::fabs(gfc) > 1e-30 ? gfu / gfc : 1.);
      st.col[0] = below, st.v[0] = 1.;
      st.col[1] = cst, st.v[1] = ap * psi * (phi.at(*g, p[0], p[1], p[2]) - phi.at(*g, lo[0], lo[1], lo[2]));
    } else {
      if (at_next) {
        st.col[0] = here, st.v[0] = 0.5, st.col[1] = below, st.v[1] = 0.5;
        return;
      }
      const double gfu = face_gradient(up[0], up[1], up[2]), gfc = face_gradient(p[0], p[1], p[2]);
      const double psi = limiter(std::fabs(gfc) > 1e-30 ? gfu / gfc : 1.);
      st.col[0] = here, st.v[0] = 1.;
      st.col[1] = cst, st.v[1] = am * psi * (phi.at(*g, lo[0], lo[1], lo[2]) - phi.at(*g, p[0], p[1], p[2]));
    }
  }
};

// reference counting as PETSc's (operands are referenced by their parents; the user's destroy drops one reference)
static void unref(Op *o);
Sum::~Sum()
{
  for (Op *o : ops) unref(o);
}
Scale::~Scale() { unref(operand); }
Composition::~Composition()
{
  unref(inner);
  unref(outer);
}
static void unref(Grid *g)
{
  if (g && --g->refs == 0) delete g;
}
static void unref(Op *o)
{
  if (!o || --o->refs > 0) return;
  Grid *g = o->g;
  delete o;
  unref(g);
}

} // namespace fd
} // namespace fluca

using namespace fluca;
using namespace fluca::fd;

struct fluca_b200_fd_grid {
  Grid *g;
};
namespace fluca {
namespace fd {
struct DevicePlan; // tables of the generated apply kernel on the device (end of this file)
void free_device_plan(DevicePlan *p);
} // namespace fd
} // namespace fluca
struct fluca_b200_fd {
  Op                    *op;
  fluca::fd::DevicePlan *plan = nullptr; // built at the first apply; dropped when locations or boundary data change
};

static thread_local std::string g_fd_err;
// bumped whenever a field an operator depends on changes (vector of a scale, velocity / current solution of a TVD operator):
// an assembled device plan of a field-dependent tree is rebuilt when it is older than this
static long g_fd_field_generation = 0;
extern "C" const char *fluca_b200_fd_last_error(void) { return g_fd_err.c_str(); }

#define FD_BEGIN try {
#define FD_END \
  } \
  catch (const fluca::Error &e) \
  { \
    g_fd_err = e.what(); \
    return e.code; \
  } \
  catch (const std::exception &e) \
  { \
    g_fd_err = e.what(); \
    return FLUCA_B200_ERR_INTERNAL; \
  } \
  return FLUCA_B200_OK;

static fluca_b200_fd *wrap(Op *o, Grid *g)
{
  o->g = g;
  ++g->refs;
  fluca_b200_fd *h = new fluca_b200_fd();
  h->op            = o;
  return h;
}
static Op *operand_of(const fluca_b200_fd *h)
{
  if (!h || !h->op) throw Error(FL_ERR_ARG, "null operator");
  if (!h->op->setupcalled) throw Error(FL_ERR_ARG, "Operand FlucaFD is not set up. Call fluca_b200_fd_setup() on the operand first");
  return h->op;
}

extern "C" int fluca_b200_fd_grid_create(int dim, const int n[3], const double *const xf[3], const double *const xc[3], const int periodic[3], int stencil_width, fluca_b200_fd_grid **out)
{
  FD_BEGIN
  if (dim < 1 || dim > 3 || !n || !xf || !out || stencil_width < 1) throw Error(FL_ERR_ARG, "bad grid arguments");
  std::unique_ptr<Grid> g(new Grid());
  g->dim = dim, g->sw = stencil_width;
  for (int d = 0; d < dim; ++d) {
    if (n[d] < 2 || !xf[d]) throw Error(FL_ERR_ARG, "a direction needs at least 2 elements and its face coordinates");
    g->N[d]   = n[d];
    g->per[d] = periodic ? (periodic[d] != 0) : 0;
    g->xf[d].assign(xf[d], xf[d] + n[d] + 1);
    g->xc[d].resize(n[d]);
    for (int i = 0; i < n[d]; ++i) g->xc[d][i] = (xc && xc[d]) ? xc[d][i] : 0.5 * (xf[d][i] + xf[d][i + 1]);
  }
  fluca_b200_fd_grid *h = new fluca_b200_fd_grid;
  h->g                  = g.release();
  *out                  = h;
  FD_END
}
extern "C" int fluca_b200_fd_grid_destroy(fluca_b200_fd_grid *h)
{
  FD_BEGIN
  if (h) {
    unref(h->g);
    delete h;
  }
  FD_END
}

extern "C" int fluca_b200_fd_derivative_create(fluca_b200_fd_grid *grid, int dir, int deriv_order, int accu_order, int input_loc, int input_c, int output_loc, int output_c, fluca_b200_fd **out)
{
  FD_BEGIN
  if (!grid || !out) throw Error(FL_ERR_ARG, "null argument");
  if (!valid_location(input_loc) || !valid_location(output_loc)) throw Error(FL_ERR_ARG, "Invalid stencil location; only ELEMENT, LEFT, DOWN, BACK, and their combinations are allowed");
  Derivative *d  = new Derivative();
  d->dir         = dir, d->deriv_order = deriv_order, d->accu_order = accu_order;
  d->input_loc   = input_loc, d->input_c = input_c, d->output_loc = output_loc, d->output_c = output_c;
  *out           = wrap(d, grid->g);
  FD_END
}
extern "C" int fluca_b200_fd_sum_create(int n, fluca_b200_fd *const ops[], fluca_b200_fd **out)
{
  FD_BEGIN
  if (n < 1 || !ops || !out) throw Error(FL_ERR_ARG, "Number of operands must be positive");
  for (int q = 0; q < n; ++q) (void)operand_of(ops[q]); // all operands set up (sum.c:120-123) before any reference is taken
  Op  *first = operand_of(ops[0]);
  Sum *s     = new Sum();
  s->input_loc = s->output_loc = first->output_loc, s->input_c = s->output_c = first->output_c;
  for (int q = 0; q < n; ++q) {
    Op *o = operand_of(ops[q]);
    ++o->refs;
    s->ops.push_back(o);
  }
  *out = wrap(s, first->g);
  FD_END
}
extern "C" int fluca_b200_fd_scale_create_constant(fluca_b200_fd *operand, double constant, fluca_b200_fd **out)
{
  FD_BEGIN
  if (!out) throw Error(FL_ERR_ARG, "null argument");
  Op    *o = operand_of(operand);
  Scale *s = new Scale();
  s->operand = o, ++o->refs;
  s->is_constant = true, s->constant = constant;
  s->input_loc = s->output_loc = o->output_loc, s->input_c = s->output_c = o->output_c;
  *out = wrap(s, o->g);
  FD_END
}
extern "C" int fluca_b200_fd_scale_create_vector(fluca_b200_fd *operand, const double *field, int vec_loc, int vec_c, fluca_b200_fd **out)
{
  FD_BEGIN
  if (!out || !field) throw Error(FL_ERR_ARG, "null argument");
  if (!valid_location(vec_loc)) throw Error(FL_ERR_ARG, "Invalid stencil location");
  Op    *o = operand_of(operand);
  Scale *s = new Scale();
  s->operand = o, ++o->refs;
  s->is_constant = false, s->vec_loc = vec_loc, s->vec_c = vec_c;
  s->vec.assign(*o->g, vec_loc, field);
  s->input_loc = s->output_loc = o->output_loc, s->input_c = s->output_c = o->output_c;
  *out = wrap(s, o->g);
  FD_END
}
extern "C" int fluca_b200_fd_composition_create(fluca_b200_fd *inner, fluca_b200_fd *outer, fluca_b200_fd **out)
{
  FD_BEGIN
  if (!out) throw Error(FL_ERR_ARG, "null argument");
  Op          *in = operand_of(inner), *ou = operand_of(outer);
  Composition *c  = new Composition();
  c->inner = in, c->outer = ou, ++in->refs, ++ou->refs;
  c->input_loc = in->input_loc, c->input_c = in->input_c, c->output_loc = ou->output_loc, c->output_c = ou->output_c;
  *out = wrap(c, in->g);
  FD_END
}
extern "C" int fluca_b200_fd_tvd_create(fluca_b200_fd_grid *grid, int dir, int input_c, int output_c, fluca_b200_fd **out)
{
  FD_BEGIN
  if (!grid || !out || dir < 0 || dir > 2) throw Error(FL_ERR_ARG, "bad argument");
  static const int face_of[3] = {LOC_LEFT, LOC_DOWN, LOC_BACK};
  TVD *t = new TVD();
  t->dir = dir, t->input_loc = LOC_ELEMENT, t->input_c = input_c, t->output_loc = face_of[dir], t->output_c = output_c;
  *out = wrap(t, grid->g);
  FD_END
}
extern "C" int fluca_b200_fd_tvd_set_limiter(fluca_b200_fd *h, const char *name)
{
  FD_BEGIN
  TVD *t = h ? dynamic_cast<TVD *>(h->op) : nullptr;
  if (!t || !name) throw Error(FL_ERR_ARG, "not a secondordertvd operator");
  t->limiter = find_limiter(name);
  FD_END
}
extern "C" int fluca_b200_fd_tvd_set_velocity(fluca_b200_fd *h, const double *face_velocity)
{
  FD_BEGIN
  TVD *t = h ? dynamic_cast<TVD *>(h->op) : nullptr;
  if (!t || !face_velocity) throw Error(FL_ERR_ARG, "not a secondordertvd operator");
  t->vel.assign(*t->g, t->output_loc, face_velocity);
  ++g_fd_field_generation;
  FD_END
}
extern "C" int fluca_b200_fd_tvd_set_current_solution(fluca_b200_fd *h, const double *phi)
{
  FD_BEGIN
  TVD *t = h ? dynamic_cast<TVD *>(h->op) : nullptr;
  if (!t || !phi) throw Error(FL_ERR_ARG, "not a secondordertvd operator");
  t->phi.assign(*t->g, LOC_ELEMENT, phi);
  ++g_fd_field_generation;
  FD_END
}
extern "C" int fluca_b200_fd_set_locations(fluca_b200_fd *h, int input_loc, int input_c, int output_loc, int output_c)
{
  FD_BEGIN
  if (!h || !h->op) throw Error(FL_ERR_ARG, "null operator");
  if (!valid_location(input_loc) || !valid_location(output_loc)) throw Error(FL_ERR_ARG, "Invalid stencil location; only ELEMENT, LEFT, DOWN, BACK, and their combinations are allowed");
  h->op->input_loc = input_loc, h->op->input_c = input_c, h->op->output_loc = output_loc, h->op->output_c = output_c;
  h->op->setupcalled = false;
  free_device_plan(h->plan), h->plan = nullptr;
  FD_END
}
extern "C" int fluca_b200_fd_set_boundary_condition(fluca_b200_fd *h, int boundary, int type, double value)
{
  FD_BEGIN
  if (!h || !h->op || boundary < 0 || boundary > 5) throw Error(FL_ERR_ARG, "bad boundary");
  if (type != FLUCA_B200_FD_BC_NONE && type != FLUCA_B200_FD_BC_DIRICHLET && type != FLUCA_B200_FD_BC_NEUMANN) throw Error(FL_ERR_ARG, "Unsupported boundary condition type");
  h->op->bc_type[boundary] = type, h->op->bc_value[boundary] = value;
  free_device_plan(h->plan), h->plan = nullptr; // boundary values are folded into the kernel's per-variant constants
  FD_END
}
extern "C" int fluca_b200_fd_setup(fluca_b200_fd *h)
{
  FD_BEGIN
  if (!h || !h->op) throw Error(FL_ERR_ARG, "null operator");
  if (!h->op->setupcalled) {
    h->op->setup();
    h->op->setupcalled = true;
  }
  FD_END
}
extern "C" int fluca_b200_fd_get_stencil(fluca_b200_fd *h, int i, int j, int k, int *ncols, fluca_b200_fd_col col[FLUCA_B200_FD_MAX_STENCIL], double v[FLUCA_B200_FD_MAX_STENCIL])
{
  FD_BEGIN
  if (!h || !h->op || !ncols || !col || !v) throw Error(FL_ERR_ARG, "null argument");
  Stencil st;
  h->op->stencil(i, j, k, st);
  *ncols = st.n;
  for (int q = 0; q < st.n; ++q) col[q].i = st.col[q].i, col[q].j = st.col[q].j, col[q].k = st.col[q].k, col[q].loc = st.col[q].loc, col[q].c = st.col[q].c, v[q] = st.v[q];
  FD_END
}
// ------------------------------------------------------------------ matrix-free device apply (FlucaFDApply, fdapply.c:47-121)
// v1: operators whose stencil depends on the output point only through its distance class from the non-periodic boundaries
// -- derivative / sum / constant scale / composition on uniform product coordinates.  The host generates the stencil of one
// representative point per class triple with the stencil layer above ((2R + 1)^dim variants, R = reach of the composed
// stencil), folds boundary values and constant terms into one number per variant, and the kernel evaluates
//   y(P) = const[variant(P)] + sum_taps w * x_slot(P + offset)            (periodic directions wrap)
// straight from the input fields: nothing is assembled per point.  Field-dependent operators (vector scale, TVD) and
// non-uniform coordinates are rejected, not approximated.  Status: parity-tested in the host-emulation build; the CUDA
// path has not run on a B200 yet.
namespace fluca {
namespace fd {

static const int MAX_SLOTS = 4;

struct ApplyFunctor {
  int           dim, R;
  int           E[3];         // output extents
  int           per[3];
  int           ncls[3];
  const int    *tap_start;    // [nvar + 1]
  const int    *tap_meta;     // [ntaps][4]: di, dj, dk, slot
  const double *tap_w;        // [ntaps]
  const double *var_const;    // [nvar]
  const double *in[MAX_SLOTS];
  int           in_ext[MAX_SLOTS][3];
  double       *out;
  // fast path of the points whose whole stencil is the interior variant and stays inside the arrays (no wrap): the taps
  // travel with the kernel arguments (constant bank) as precomputed linear offsets -- no table load, no index arithmetic per tap
  int           fast_n;      // 0: no fast path
  int           fast_lo[3], fast_hi[3]; // output indices [lo, hi) per direction that qualify
  int           fast_slot[FLUCA_B200_FD_MAX_STENCIL];
  int           fast_off[FLUCA_B200_FD_MAX_STENCIL];
  double        fast_w[FLUCA_B200_FD_MAX_STENCIL];
  double        fast_const;
  FL_HD int cls(int d, int idx) const
  {
    if (d >= dim || per[d]) return 0; // one class: a periodic direction is translation invariant
    if (idx < R) return idx;
    if (idx >= E[d] - R) return R + 1 + (idx - (E[d] - R));
    return R;
  }
  FL_HD void operator()(int i, int j, int k) const
  {
    const bool fast = fast_n > 0 && i >= fast_lo[0] && i < fast_hi[0] && j >= fast_lo[1] && j < fast_hi[1] && k >= fast_lo[2] && k < fast_hi[2];
    if (FL_WARP_ALL(fast)) {
      // index of P in every slot's array once per point (registers; the slot of a tap selects one), then one add, one load and
      // one fma per tap.  The arrays of this path hold fewer than 2^31 entries (checked when the plan is built).
      int base[MAX_SLOTS];
#pragma unroll
      for (int s = 0; s < MAX_SLOTS; ++s) base[s] = i + in_ext[s][0] * (j + in_ext[s][1] * k);
      double acc = fast_const;
      for (int t = 0; t < fast_n; ++t) {
        const int s = fast_slot[t];
        const int b = s == 0 ? base[0] : (s == 1 ? base[1] : (s == 2 ? base[2] : base[3]));
        acc += fast_w[t] * in[s][b + fast_off[t]];
      }
      out[i + E[0] * (j + E[1] * k)] = acc;
      return;
    }
    const int    v  = (cls(2, k) * ncls[1] + cls(1, j)) * ncls[0] + cls(0, i);
    double       acc = var_const[v];
    const int    p[3] = {i, j, k};
    for (int t = tap_start[v]; t < tap_start[v + 1]; ++t) {
      const int *m = tap_meta + 4 * t;
      const int  s = m[3];
      int        q[3];
      for (int d = 0; d < 3; ++d) {
        q[d] = p[d] + m[d];
        if (d < dim && per[d]) {
          const int n = in_ext[s][d];
          q[d]        = ((q[d] % n) + n) % n;
        }
      }
      acc += tap_w[t] * in[s][(long)q[0] + (long)in_ext[s][0] * ((long)q[1] + (long)in_ext[s][1] * (long)q[2])];
    }
    out[(long)i + (long)E[0] * ((long)j + (long)E[1] * (long)k)] = acc;
  }
};

static void field_extents(const Grid &g, int loc, int ext[3])
{
  for (int d = 0; d < 3; ++d) ext[d] = d < g.dim ? g.N[d] + ((use_face(loc, d) && !g.per[d]) ? 1 : 0) : 1;
}
static bool field_dependent(const Op *o)
{
  if (dynamic_cast<const TVD *>(o)) return true;
  if (const Scale *s = dynamic_cast<const Scale *>(o)) return !s->is_constant || field_dependent(s->operand);
  if (const Sum *s = dynamic_cast<const Sum *>(o)) {
    for (const Op *q : s->ops)
      if (field_dependent(q)) return true;
    return false;
  }
  if (const Composition *c = dynamic_cast<const Composition *>(o)) return field_dependent(c->inner) || field_dependent(c->outer);
  return false;
}

struct Plan {
  int                 R = 0, ncls[3] = {1, 1, 1}, E[3] = {1, 1, 1};
  std::vector<int>    tap_start, tap_meta;
  std::vector<double> tap_w, var_const;
  int                 nslots = 0, slot_loc[MAX_SLOTS], slot_c[MAX_SLOTS];
};

// representative output index of class c along direction d
static int representative(const Plan &p, const Grid &g, int d, int c)
{
  if (d >= g.dim) return 0;
  if (g.per[d]) return p.E[d] / 2;
  if (c < p.R) return c;
  if (c == p.R) return p.R;
  return p.E[d] - p.R + (c - p.R - 1);
}

static void build_plan(Op &op, Plan &p)
{
  const Grid &g = *op.g;
  if (!op.setupcalled) throw Error(FL_ERR_ARG, "FlucaFD not setup");
  if (field_dependent(&op)) throw Error(FL_ERR_ARG, "the device apply (v1) covers derivative / sum / constant scale / composition; vector scale and TVD operators are not supported yet");
  for (int d = 0; d < g.dim; ++d) { // uniform coordinates: the class table relies on translation invariance
    const double h = (g.xf[d][g.N[d]] - g.xf[d][0]) / g.N[d];
    for (int i = 0; i < g.N[d]; ++i)
      if (std::fabs((g.xf[d][i + 1] - g.xf[d][i]) - h) > 1e-12 * std::fabs(h) || std::fabs(g.xc[d][i] - 0.5 * (g.xf[d][i] + g.xf[d][i + 1])) > 1e-12 * std::fabs(h)) throw Error(FL_ERR_ARG, "the device apply (v1) needs uniform product coordinates");
  }
  field_extents(g, op.output_loc, p.E);
  // reach of the composed stencil: the widest tap offset seen at a mid-grid point, in any direction
  Stencil st;
  int     mid[3] = {0, 0, 0};
  for (int d = 0; d < g.dim; ++d) mid[d] = p.E[d] / 2;
  op.stencil_raw(mid[0], mid[1], mid[2], st);
  int reach = 1;
  for (int q = 0; q < st.n; ++q)
    if (st.col[q].c >= 0)
      for (int d = 0; d < g.dim; ++d) {
        const int r = std::abs(st.col[q].idx(d) - mid[d]) + 1;
        reach       = std::max(reach, r);
        // beyond its ghost elements the reference extrapolates even in a periodic direction (fdutils.c:306-356); the class
        // table treats periodic directions as translation invariant, which is only the same thing while taps stay inside
        if (g.per[d] && r - 1 > g.sw) throw Error(FL_ERR_ARG, "the stencil is wider than the DMStag stencil width in a periodic direction");
      }
  p.R = reach + 1; // one more: the one-sided closures of the first interior row still see the boundary
  for (int d = 0; d < 3; ++d) {
    p.ncls[d] = (d < g.dim && !g.per[d]) ? 2 * p.R + 1 : 1;
    if (d < g.dim && !g.per[d] && p.E[d] < 2 * p.R + 2) throw Error(FL_ERR_ARG, "grid too small for this operator's boundary classes");
  }
  const int nvar = p.ncls[0] * p.ncls[1] * p.ncls[2];
  p.tap_start.assign(1, 0);
  p.var_const.assign(nvar, 0.);
  p.nslots = 0;
  for (int cz = 0; cz < p.ncls[2]; ++cz)
    for (int cy = 0; cy < p.ncls[1]; ++cy)
      for (int cx = 0; cx < p.ncls[0]; ++cx) {
        const int cc[3] = {cx, cy, cz};
        int       r[3];
        for (int d = 0; d < 3; ++d) r[d] = representative(p, g, d, g.per[d] ? p.R : cc[d]);
        op.stencil(r[0], r[1], r[2], st);
        const int v = (cz * p.ncls[1] + cy) * p.ncls[0] + cx;
        for (int q = 0; q < st.n; ++q) {
          const Col &c = st.col[q];
          if (c.c == CONSTANT) p.var_const[v] += st.v[q];
          else if (c.c < 0) p.var_const[v] += st.v[q] * op.bc_value[-c.c - 1]; // fdapply.c:99-104
          else {
            int slot = -1;
            for (int s2 = 0; s2 < p.nslots; ++s2)
              if (p.slot_loc[s2] == c.loc && p.slot_c[s2] == c.c) slot = s2;
            if (slot < 0) {
              if (p.nslots >= MAX_SLOTS) throw Error(FL_ERR_ARG, "too many distinct input fields");
              slot = p.nslots++, p.slot_loc[slot] = c.loc, p.slot_c[slot] = c.c;
            }
            p.tap_meta.push_back(c.i - r[0]), p.tap_meta.push_back(c.j - r[1]), p.tap_meta.push_back(c.k - r[2]), p.tap_meta.push_back(slot);
            p.tap_w.push_back(st.v[q]);
          }
        }
        p.tap_start.push_back((int)p.tap_w.size());
      }
}

} // namespace fd
} // namespace fluca

namespace fluca {
namespace fd {
// ---- v2: the ASSEMBLED apply, for everything the class-table kernel above rejects (vector scale, second-order TVD with its
// limiters, non-uniform coordinates, periodic stencils wider than the ghost width).  The host evaluates FlucaFDGetStencil at
// every output point -- exactly what FlucaFDApply does per point (fdapply.c:85-106) and what FlucaFDGetOperator assembles
// into a Mat (fdapply.c:123-180) -- into an ELL table (column-major: entry t of every row is contiguous), boundary values and
// constant terms folded into one number per row, and the kernel computes y = const + sum_t w_t x_slot(idx_t), one thread per
// output point.  A tree with field-dependent nodes is re-assembled when one of its fields changed (TVD: every apply after
// SetCurrentSolution, as in the reference, whose stencils are re-derived per point on every FlucaFDApply).
struct Assembled {
  long                  nrows = 0;
  int                   width = 0; // entries per row (rows are padded with zero weights)
  std::vector<double>   val;       // [width][nrows]
  std::vector<unsigned> idx;       // [width][nrows]: (slot << 28) | linear index in the slot's compact array
  std::vector<double>   cst;       // [nrows]
  int                   nslots = 0, slot_loc[MAX_SLOTS], slot_c[MAX_SLOTS];
  int                   E[3] = {1, 1, 1};
};
static const unsigned ELL_IDX_BITS = 28, ELL_IDX_MASK = (1u << ELL_IDX_BITS) - 1u;

struct EllFunctor {
  long            nrows;
  int             width;
  const double   *val;
  const unsigned *idx;
  const double   *cst;
  const double   *in[MAX_SLOTS];
  double         *out;
  FL_HD void operator()(long r) const
  {
    double acc = cst[r];
    for (int t = 0; t < width; ++t) {
      const unsigned e = idx[(size_t)t * nrows + r];
      acc += val[(size_t)t * nrows + r] * in[e >> ELL_IDX_BITS][e & ELL_IDX_MASK];
    }
    out[r] = acc;
  }
};

static void assemble(Op &op, Assembled &a)
{
  const Grid &g = *op.g;
  if (!op.setupcalled) throw Error(FL_ERR_ARG, "FlucaFD not setup");
  field_extents(g, op.output_loc, a.E);
  a.nrows = (long)a.E[0] * a.E[1] * a.E[2];
  std::vector<Stencil> rows((size_t)a.nrows);
  int                  in_ext[MAX_SLOTS][3];
  a.nslots = 0, a.width = 1;
  a.cst.assign((size_t)a.nrows, 0.);
  // pass 1: stencils, slots, row width
  for (int k = 0; k < a.E[2]; ++k)
    for (int j = 0; j < a.E[1]; ++j)
      for (int i = 0; i < a.E[0]; ++i) {
        const long r = (long)i + (long)a.E[0] * ((long)j + (long)a.E[1] * (long)k);
        Stencil   &st = rows[(size_t)r];
        op.stencil(i, j, k, st);
        int nint = 0;
        for (int q = 0; q < st.n; ++q) {
          const Col &c = st.col[q];
          if (c.c == CONSTANT) a.cst[(size_t)r] += st.v[q];
          else if (c.c < 0) a.cst[(size_t)r] += st.v[q] * op.bc_value[-c.c - 1]; // fdapply.c:99-104
          else {
            ++nint;
            int slot = -1;
            for (int s2 = 0; s2 < a.nslots; ++s2)
              if (a.slot_loc[s2] == c.loc && a.slot_c[s2] == c.c) slot = s2;
            if (slot < 0) {
              if (a.nslots >= MAX_SLOTS) throw Error(FL_ERR_ARG, "too many distinct input fields");
              slot = a.nslots++, a.slot_loc[slot] = c.loc, a.slot_c[slot] = c.c;
              field_extents(g, c.loc, in_ext[slot]);
              if ((long)in_ext[slot][0] * in_ext[slot][1] * in_ext[slot][2] > (long)ELL_IDX_MASK) throw Error(FL_ERR_ARG, "input field too large for the assembled apply (2^28 entries)");
            }
          }
        }
        a.width = std::max(a.width, nint);
      }
  // pass 2: the ELL table
  a.val.assign((size_t)a.width * a.nrows, 0.);
  a.idx.assign((size_t)a.width * a.nrows, 0u);
  for (long r = 0; r < a.nrows; ++r) {
    const Stencil &st = rows[(size_t)r];
    int            t  = 0;
    for (int q = 0; q < st.n; ++q) {
      const Col &c = st.col[q];
      if (c.c < 0) continue;
      int slot = 0;
      for (int s2 = 0; s2 < a.nslots; ++s2)
        if (a.slot_loc[s2] == c.loc && a.slot_c[s2] == c.c) slot = s2;
      int qi[3] = {c.i, c.j, c.k};
      for (int d = 0; d < 3; ++d) {
        const int n = in_ext[slot][d];
        if (d < g.dim && g.per[d]) qi[d] = ((qi[d] % n) + n) % n; // ghost elements of a periodic direction (DMGlobalToLocal fills them)
        if (qi[d] < 0 || qi[d] >= n) throw Error(FL_ERR_INTERNAL, "stencil point outside the input field after off-grid removal");
      }
      const unsigned lin = (unsigned)((long)qi[0] + (long)in_ext[slot][0] * ((long)qi[1] + (long)in_ext[slot][1] * (long)qi[2]));
      a.val[(size_t)t * a.nrows + r] = st.v[q];
      a.idx[(size_t)t * a.nrows + r] = ((unsigned)slot << ELL_IDX_BITS) | lin;
      ++t;
    }
  }
}

struct DevicePlan {
  Exec                ex;
  Plan                p;
  ApplyFunctor        f; // table pointers filled in; in[] / out set per launch
  bool                assembled = false; // v2 (ELL) instead of the class table
  bool                field_dep = false; // the tree holds field-dependent nodes: rebuilt when a field changed
  long                generation = 0;
  Assembled           a;
  EllFunctor          ef;
  std::vector<void *> owned;
};
void free_device_plan(DevicePlan *dp)
{
  if (!dp) return;
  try {
    dp->ex.sync();
  } catch (...) {
  }
  for (void *d : dp->owned) dev_free(d);
  dp->ex.destroy();
  delete dp;
}
// can the class-table kernel (v1) serve this operator?  (build_plan throws FL_ERR_ARG with the reason otherwise)
static bool class_table_applies(Op &op, Plan &p)
{
  if (getenv("FLUCA_B200_FD_ASSEMBLED")) return false; // tests: force the assembled path
  try {
    build_plan(op, p);
    return true;
  } catch (const Error &e) {
    if (e.code != FL_ERR_ARG || !op.setupcalled) throw;
    return false;
  }
}

static DevicePlan *device_plan(fluca_b200_fd *h)
{
  if (h->plan && h->plan->field_dep && h->plan->generation != g_fd_field_generation) free_device_plan(h->plan), h->plan = nullptr;
  if (h->plan) return h->plan;
#ifndef FLUCA_HOSTEMU
  {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) throw Error(FL_ERR_NODEVICE, "fluca_b200 needs a CUDA device (sm_100a); there is no CPU fallback");
  }
#endif
  std::unique_ptr<DevicePlan> dp(new DevicePlan());
  dp->assembled = !class_table_applies(*h->op, dp->p);
  dp->field_dep = field_dependent(h->op), dp->generation = g_fd_field_generation;
  if (dp->assembled) assemble(*h->op, dp->a);
  dp->ex.init();
  const Grid &g = *h->op->g;
  const Plan &p = dp->p;
  auto        up = [&](const void *src, size_t bytes) {
    void *d = dev_alloc(bytes);
    dp->owned.push_back(d);
    copy_h2d(dp->ex, d, src, bytes);
    return d;
  };
  ApplyFunctor &f = dp->f;
  if (dp->assembled) {
    try {
      const Assembled &a = dp->a;
      EllFunctor      &e = dp->ef;
      e.nrows = a.nrows, e.width = a.width;
      e.val = (const double *)up(a.val.data(), sizeof(double) * a.val.size());
      e.idx = (const unsigned *)up(a.idx.data(), sizeof(unsigned) * a.idx.size());
      e.cst = (const double *)up(a.cst.data(), sizeof(double) * a.cst.size());
      for (int s2 = 0; s2 < MAX_SLOTS; ++s2) e.in[s2] = nullptr;
      e.out = nullptr;
      // the v1 bookkeeping the entry points read: slots and extents
      dp->p.nslots = a.nslots;
      for (int s2 = 0; s2 < a.nslots; ++s2) dp->p.slot_loc[s2] = a.slot_loc[s2], dp->p.slot_c[s2] = a.slot_c[s2], field_extents(g, a.slot_loc[s2], f.in_ext[s2]);
      for (int d = 0; d < 3; ++d) dp->p.E[d] = a.E[d];
      dp->ex.sync();
      std::vector<double>().swap(dp->a.val), std::vector<unsigned>().swap(dp->a.idx), std::vector<double>().swap(dp->a.cst); // the device holds them now
    } catch (...) {
      free_device_plan(dp.release());
      throw;
    }
    h->plan = dp.release();
    return h->plan;
  }
  try {
    f.dim = g.dim, f.R = p.R;
    for (int d = 0; d < 3; ++d) f.E[d] = p.E[d], f.per[d] = d < g.dim ? g.per[d] : 0, f.ncls[d] = p.ncls[d];
    f.tap_start = (const int *)up(p.tap_start.data(), sizeof(int) * p.tap_start.size());
    f.tap_meta  = (const int *)up(p.tap_meta.data(), sizeof(int) * std::max<size_t>(p.tap_meta.size(), 1));
    f.tap_w     = (const double *)up(p.tap_w.data(), sizeof(double) * std::max<size_t>(p.tap_w.size(), 1));
    f.var_const = (const double *)up(p.var_const.data(), sizeof(double) * p.var_const.size());
    for (int s2 = 0; s2 < MAX_SLOTS; ++s2) {
      f.in[s2] = nullptr;
      for (int d = 0; d < 3; ++d) f.in_ext[s2][d] = 1;
    }
    for (int s2 = 0; s2 < p.nslots; ++s2) field_extents(g, p.slot_loc[s2], f.in_ext[s2]);
    f.out = nullptr;
    {
      // interior variant: class R in every non-periodic direction (class 0 in periodic ones)
      int cv[3];
      for (int d = 0; d < 3; ++d) cv[d] = (d < g.dim && !g.per[d]) ? p.R : 0;
      const int v = (cv[2] * p.ncls[1] + cv[1]) * p.ncls[0] + cv[0];
      const int t0 = p.tap_start[v], t1 = p.tap_start[v + 1];
      f.fast_n = 0;
      bool small = (long)p.E[0] * p.E[1] * p.E[2] < (1L << 31);
      for (int s2 = 0; s2 < p.nslots; ++s2) small = small && (long)f.in_ext[s2][0] * f.in_ext[s2][1] * f.in_ext[s2][2] < (1L << 31);
      if (small && t1 - t0 <= FLUCA_B200_FD_MAX_STENCIL) {
        f.fast_n     = t1 - t0;
        f.fast_const = p.var_const[v];
        int reach[3] = {0, 0, 0};
        for (int t = t0; t < t1; ++t) {
          const int *m = &p.tap_meta[4 * (size_t)t];
          const int  sl = m[3];
          f.fast_slot[t - t0] = sl;
          f.fast_off[t - t0]  = m[0] + f.in_ext[sl][0] * (m[1] + f.in_ext[sl][1] * m[2]);
          f.fast_w[t - t0]    = p.tap_w[t];
          for (int d = 0; d < 3; ++d) reach[d] = std::max(reach[d], std::abs(m[d]) + 1); // + 1: face inputs have one more entry
        }
        for (int d = 0; d < 3; ++d) {
          if (d >= g.dim) f.fast_lo[d] = 0, f.fast_hi[d] = 1;
          else if (g.per[d]) f.fast_lo[d] = reach[d], f.fast_hi[d] = p.E[d] - reach[d]; // taps must not wrap
          else f.fast_lo[d] = p.R, f.fast_hi[d] = p.E[d] - p.R;                           // the interior class
        }
      }
    }
    dp->ex.sync(); // the tables were staged from this function's host vectors
  } catch (...) {
    free_device_plan(dp.release());
    throw;
  }
  h->plan = dp.release();
  return h->plan;
}
} // namespace fd
} // namespace fluca

extern "C" int fluca_b200_fd_apply_inputs(fluca_b200_fd *h, int *ninputs, int loc[4], int c[4])
{
  FD_BEGIN
  if (!h || !h->op || !ninputs || !loc || !c) throw Error(FL_ERR_ARG, "null argument");
  Plan p;
  if (!class_table_applies(*h->op, p)) { // v2: the slots come from the assembled table (same discovery order as apply uses)
    Assembled a;
    assemble(*h->op, a);
    p.nslots = a.nslots;
    for (int s = 0; s < a.nslots; ++s) p.slot_loc[s] = a.slot_loc[s], p.slot_c[s] = a.slot_c[s];
  }
  *ninputs = p.nslots;
  for (int s = 0; s < p.nslots; ++s) loc[s] = p.slot_loc[s], c[s] = p.slot_c[s];
  FD_END
}

// device-resident form: inputs / output are device pointers (compact layout); the launch is asynchronous on the operator's
// stream (fluca_b200_fd_stream), ordered with nothing else -- the caller synchronises (fluca_b200_fd_sync or its own events)
extern "C" int fluca_b200_fd_apply_device(fluca_b200_fd *h, int ninputs, const double *const dev_inputs[], double *dev_output)
{
  FD_BEGIN
  if (!h || !h->op || !dev_inputs || !dev_output) throw Error(FL_ERR_ARG, "null argument");
  DevicePlan *dp = device_plan(h);
  if (ninputs != dp->p.nslots) throw Error(FL_ERR_ARG, "number of input fields does not match fluca_b200_fd_apply_inputs");
  for (int s = 0; s < ninputs; ++s)
    if (!dev_inputs[s]) throw Error(FL_ERR_ARG, "null input field");
  if (dp->assembled) {
    EllFunctor e = dp->ef;
    for (int s = 0; s < ninputs; ++s) e.in[s] = dev_inputs[s];
    e.out = dev_output;
    for_range(dp->ex, e.nrows, e);
  } else {
    ApplyFunctor f = dp->f;
    for (int s = 0; s < ninputs; ++s) f.in[s] = dev_inputs[s];
    f.out = dev_output;
    Box b = {dp->p.E[0], dp->p.E[1], dp->p.E[2]};
    for_box(dp->ex, b, f);
  }
  FD_END
}
extern "C" int fluca_b200_fd_stream(fluca_b200_fd *h, void **stream)
{
  FD_BEGIN
  if (!h || !h->op || !stream) throw Error(FL_ERR_ARG, "null argument");
  *stream = (void *)device_plan(h)->ex.stream;
  FD_END
}
extern "C" int fluca_b200_fd_sync(fluca_b200_fd *h)
{
  FD_BEGIN
  if (!h || !h->op) throw Error(FL_ERR_ARG, "null argument");
  if (h->plan) h->plan->ex.sync();
  FD_END
}

extern "C" int fluca_b200_fd_apply(fluca_b200_fd *h, int ninputs, const double *const inputs[], double *output)
{
  FD_BEGIN
  if (!h || !h->op || !inputs || !output) throw Error(FL_ERR_ARG, "null argument");
  DevicePlan *dp = device_plan(h);
  const Plan &p  = dp->p;
  if (ninputs != p.nslots) throw Error(FL_ERR_ARG, "number of input fields does not match fluca_b200_fd_apply_inputs");
  std::vector<void *> tmp;
  try {
    ApplyFunctor f = dp->f;
    for (int s = 0; s < p.nslots; ++s) {
      if (!inputs[s]) throw Error(FL_ERR_ARG, "null input field");
      const size_t bytes = sizeof(double) * (size_t)f.in_ext[s][0] * f.in_ext[s][1] * f.in_ext[s][2];
      void        *d     = dev_alloc(bytes);
      tmp.push_back(d);
      copy_h2d(dp->ex, d, inputs[s], bytes);
      f.in[s] = (const double *)d;
    }
    const size_t nout = (size_t)p.E[0] * p.E[1] * p.E[2];
    f.out             = (double *)dev_alloc(sizeof(double) * nout);
    tmp.push_back(f.out);
    if (dp->assembled) {
      EllFunctor e = dp->ef;
      for (int s = 0; s < p.nslots; ++s) e.in[s] = f.in[s];
      e.out = f.out;
      for_range(dp->ex, e.nrows, e);
    } else {
      Box b = {p.E[0], p.E[1], p.E[2]};
      for_box(dp->ex, b, f);
    }
    copy_d2h(dp->ex, output, f.out, sizeof(double) * nout);
    dp->ex.sync();
  } catch (...) {
    try {
      dp->ex.sync();
    } catch (...) {
    }
    for (void *d : tmp) dev_free(d);
    throw;
  }
  for (void *d : tmp) dev_free(d);
  FD_END
}

// FlucaFDGetOperator (fdapply.c:123-180): the matrix of the operator -- interior stencil points only, boundary and constant terms
// left out, as the reference leaves them out of the Mat -- as host CSR.  Two calls: with rowptr == NULL it reports the sizes.
extern "C" int fluca_b200_fd_get_operator(fluca_b200_fd *h, long *nrows, long *nnz, long *rowptr, fluca_b200_fd_col *cols, double *vals)
{
  FD_BEGIN
  if (!h || !h->op || !nrows || !nnz) throw Error(FL_ERR_ARG, "null argument");
  Op &op = *h->op;
  if (!op.setupcalled) throw Error(FL_ERR_ARG, "FlucaFD not setup");
  int E[3];
  field_extents(*op.g, op.output_loc, E);
  const long nr = (long)E[0] * E[1] * E[2];
  long       at = 0;
  Stencil    st;
  for (int k = 0; k < E[2]; ++k)
    for (int j = 0; j < E[1]; ++j)
      for (int i = 0; i < E[0]; ++i) {
        const long r = (long)i + (long)E[0] * ((long)j + (long)E[1] * (long)k);
        if (rowptr) rowptr[r] = at;
        op.stencil(i, j, k, st);
        for (int q = 0; q < st.n; ++q) {
          if (st.col[q].c < 0) continue;
          if (rowptr) {
            if (at >= *nnz) throw Error(FL_ERR_ARG, "the arrays are smaller than the operator (call with rowptr = NULL for the sizes)");
            cols[at].i = st.col[q].i, cols[at].j = st.col[q].j, cols[at].k = st.col[q].k, cols[at].loc = st.col[q].loc, cols[at].c = st.col[q].c;
            vals[at] = st.v[q];
          }
          ++at;
        }
      }
  if (rowptr) rowptr[nr] = at;
  *nrows = nr, *nnz = at;
  FD_END
}

extern "C" int fluca_b200_fd_destroy(fluca_b200_fd *h)
{
  FD_BEGIN
  if (h) {
    free_device_plan(h->plan);
    unref(h->op);
    delete h;
  }
  FD_END
}
