/*
 * oracle/fluca_oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C + OpenMP, no PETSc) of the Navier-Stokes time step of
 * thecasterian/fluca: NS type "cnlinear" + PC "abf" on a MeshCart mesh.  It follows
 *   fluca/src/ns/utils/cartdiscret.c            (24 closed-form stencil formulas)
 *   fluca/src/ns/impl/linearcn/cnlinearcart2d.c (operators, BC vectors, RHS, step; 2-D)
 *   fluca/src/ns/impl/linearcn/cnlinearcart3d.c (same, 3-D)
 *   fluca/src/ns/utils/abfpc/abfpc.c:48-182     (PCSetUp_ABF / PCApply_ABF = the fractional step)
 *   fluca/src/ns/interface/nsbasic.c:215-251, nssol.c:13-30 (null space, zero guess, tolerances)
 * of /root/reference.  The arithmetic the reference delegates to PETSc (>= 3.23, un-vendored,
 * unpinned: fluca/CMakeLists.txt:9-11) -- MatMult, MatMatMult, GMRES(30), ILU(0)/block-Jacobi --
 * is restated in oracle/src/sparse.c from the published algorithms.
 *
 * PARITY STATUS: PINNED to the reference's own code for everything the reference computes itself, UNPINNED for what it
 * delegates to PETSc.  The reference registers no NS test and stores no NS golden output and cannot be built as a
 * whole here (needs PETSc, MPI, HDF5, CGNS) -- but its NS sources (cartdiscret.c, cnlinear.c, cnlinearcart2d.c, cnlinearcart3d.c,
 * abfpc.c) compile, from where they lie under /root/reference, against a single-rank model of the PETSc API subset they use
 * (oracle/ref_model/, `make -C oracle ref` -> oracle/_ref/libfluca_ref_ns.so).  tests/test_oracle_vs_reference.py: every assembled
 * operator of this oracle equals the reference's entry for entry (stored zeros included), and right-hand side, PCABF application
 * (ID / DIAG / ROWSUM), the state after K steps in both solve modes and the outer GMRES residual history agree to round-off on 2-D
 * and 3-D cases with every boundary type, uniform and stretched; the reference's outputs are committed as
 * tests/golden/ns_reference.npz.  Also pinned: the stencil coefficients against the reference's fd golden outputs
 * (tests/golden/fd_coefficients.json) and the Taylor-Green analytic solution.  NOT pinned: PETSc's own arithmetic (inner GMRES /
 * ILU(0) iteration histories; the model's solves are exact) and the immersed-boundary section (the reference has no IBM code).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * use this library.  The product path (fluca_b200/) never links or loads it.
 */
#pragma once
#ifdef __cplusplus
extern "C" {
#endif

/* same numeric values as NSBoundaryConditionType, fluca/include/flucansbc.h:5-11 */
enum { ORC_BC_NONE = 0, ORC_BC_VELOCITY = 1, ORC_BC_PRESSURE_OUTLET = 2, ORC_BC_PERIODIC = 3, ORC_BC_SYMMETRY = 4 };

/* same signature as NSBoundaryConditionFunction, flucansbc.h:14 (PetscErrorCode -> int) */
typedef int (*OrcBCFn)(int dim, double t, const double x[], double val[], void *ctx);

/* same fields as NSBoundaryCondition, flucansbc.h:16-22 */
typedef struct {
  int     type;
  OrcBCFn velocity;
  void   *ctx_velocity;
  OrcBCFn pressure;
  void   *ctx_pressure;
} OrcBC;

typedef struct Orc Orc;

typedef struct {
  int    mode;            /* 0 = Mode A: outer GMRES(30) on the coupled system, PC = ABF (reference default);
                             1 = Mode B: one ABF application (-ns_ksp_type preonly) = classical fractional step */
  double outer_rtol;      /* nssol.c:24: 1e-5 */
  int    outer_maxit;     /* PETSc default 10000 */
  double mom_rtol;        /* inner KSP "abf_momentum_": PETSc default 1e-5 */
  double schur_rtol;      /* inner KSP "abf_schur_":    PETSc default 1e-5 */
  int    inner_maxit;     /* PETSc default 10000 */
  int    ilu_blocks;      /* 1 = serial ILU(0); P = bjacobi+ILU(0) of a P-rank run */
  int    exact_schur;     /* 0: S by sparse products as abfpc.c:151-170; 1: S = -D*Gst assembled directly */
  int    quirk_bcg_scale; /* 1 (default): 3-D RHS scales the outlet gradient BC vector by 1 like
                             cnlinearcart3d.c:2977; 0: use dt/rho as the 2-D file does */
  int    schur_ainv;      /* -ns_pc_abf_schur_ainv_type: 0 ID (default, abfpc.c:328), 1 DIAG, 2 ROWSUM (abfpc.c:151-168) */
  int    upper_ainv;      /* -ns_pc_abf_upper_ainv_type: same values (abfpc.c:80-94) */
} OrcOptions;

typedef struct {
  int    outer_its, mom_its, schur_its; /* totals over the step */
  int    abf_applies;
  int    converged;
  double outer_rnorm0, outer_rnorm;
  int    nhist;
  double hist[512]; /* outer KSP true-residual history */
} OrcStepInfo;

void orc_default_options(OrcOptions *o);

/* n[d] cells, periodic[d] flags, xf[d] = n[d]+1 face coordinates (centres are face midpoints,
 * cart.c:497).  bcs ordered LEFT,RIGHT,DOWN,UP,BACK,FRONT (cart.c:564-591). */
/* 1 (default): objects created from now on form the 3-D upper-outlet rows of T as cnlinearcart3d.c:1996,2055,2114 do (weights
   -1/3, 4/3); 0: as the 2-D file does (-1/8, 9/8 on a uniform mesh).  See interp_row in src/ns.c. */
void orc_set_t_outlet_quirk(int on);
Orc *orc_create(int dim, const int n[3], const int periodic[3], const double *const xf[3], double rho, double mu, double dt, const OrcBC bcs[6]);
void orc_destroy(Orc *o);

/* sizes: ncell, nface[d] */
void orc_sizes(const Orc *o, long *ncell, long nface[3]);

/* state, SoA: v[c*ncell + cell]; U_d[face]; cell = i + nx*(j + ny*k);
 * x-face = i + nfx*(j + ny*k), y-face = i + nx*(j + nfy*k), z-face = i + nx*(j + ny*k) */
void orc_set_state(Orc *o, const double *v, const double *const U[3], const double *p, const double *phalf, int step, double t);
void orc_get_state(const Orc *o, double *v, double *const U[3], double *p, double *phalf, int *step, double *t);

/* one NSStep (nsbasic.c:276-299 + cnlinearcart{2,3}d.c NSStep_CNLinear_*_Internal) */
int orc_step(Orc *o, const OrcOptions *opt, OrcStepInfo *info);

/* ---- operator-level access, for tests ---- */
/* names: "G" (scaled dt/rho), "L", "T", "B", "D", "Gst" (scaled dt/rho), "A", "C", "S", "negR".
 * A, C, S refer to the operators of the most recent orc_prepare_step / orc_step. Returns 0 or -1. */
int  orc_matrix(const Orc *o, const char *name, int *nrows, int *ncols, long *nnz, const int **ptr, const int **idx, const double **val);
/* builds v0interp, A and the RHS b = (r_mom, r_int, r_con) for the current state (NSFormFunction +
 * NSFormJacobian(UPDATE)); rhs has length dim*ncell + sum(nface) + ncell */
void orc_prepare_step(Orc *o, const OrcOptions *opt, double *rhs);
/* one PCApply_ABF (abfpc.c:48-111): x = ABF(b); sizes as rhs */
void orc_abf_apply(Orc *o, const OrcOptions *opt, const double *b, double *x, OrcStepInfo *info);
/* 1-D stencil formulas of cartdiscret.c, by name, on an explicit coordinate tuple; returns ncols */
int orc_formula(const char *name, const double *xs, double h, double vf, double w[4], int off[4]);

/* ---- immersed boundary: NOT a restatement (the reference has no IBM code); it defines the
 * direct-forcing coupling the CUDA path implements (see the IBM section of src/ns.c).  PARITY UNPINNED.
 * X, Ud, Um, F: dim consecutive blocks of n; dV: n.  npts: 4 (Peskin) or 3 (Roma). */
void orc_set_markers(Orc *o, long n, const double *X, const double *Ud, const double *dV, int npts);
void orc_set_ibm_iterations(Orc *o, int n); /* multi-direct forcing passes per step (default 1) */
void orc_ibm_interpolate(const Orc *o, const double *v, double *Um);
void orc_ibm_spread(const Orc *o, const double *Fm, double *f); /* f += spread(Fm * dV / vol) */
void orc_get_marker_forces(const Orc *o, double *F, double *Um); /* of the last step; either may be NULL */

/* built-in BC callbacks for timing runs (ctx = double[3] constant value) */
int orc_bc_constant(int dim, double t, const double x[], double val[], void *ctx);
/* pressure flavour: ctx = double[1] */
int orc_bc_constant_pressure(int dim, double t, const double x[], double val[], void *ctx);

/* OpenMP threads of the timed CPU legs: set (n > 0) and report what the runtime will use / actually spawns */
int orc_set_threads(int n);
int orc_get_threads(void);

#ifdef __cplusplus
}
#endif
