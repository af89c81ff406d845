/*
 * fluca_b200.h -- C ABI of the B200-native Navier-Stokes time-step solver.
 *
 * This is the drop-in boundary for the hot path of thecasterian/fluca: a
 * new NS type (registered with NSRegister, fluca/include/flucans.h:92, selected with
 * -ns_type b200) calls these entry points from its NSOps (fluca/include/fluca/private/nsimpl.h:21-31).
 * Plain C types only; every function returns 0 on success or a FLUCA_B200_ERR_* code, and
 * fluca_b200_last_error() returns the message (the glue maps it to PETSC_ERR_LIB / ns->reason,
 * nsbasic.c:293-297, :425-436).  INTEGRATION.md shows the reference-side binding.
 *
 * There is NO CPU fallback: creating a solver without a CUDA device fails with
 * FLUCA_B200_ERR_NODEVICE.
 *
 * Field layout at this boundary (host or device memory, compact, no padding; one rank's z-slab):
 *   cell fields     [nzl][ny][nx]             velocity: dim consecutive component blocks (SoA)
 *   x-face field    [nzl][ny][nx + ex]        ex = 0 if x periodic else 1 (the extra RIGHT face)
 *   y-face field    [nzl][ny + ey][nx]
 *   z-face field    [nzl + ez][ny][nx]        ez = 1 only on the rank that holds the FRONT wall
 * i.e. every cell owns its LEFT/DOWN/BACK face, the last rank the extra faces, as DMStag does
 * (fluca/src/mesh/impl/cart/cart.c:85-120); the PETSc glue converts with DMStagVecGetArray +
 * DMStagGetLocationSlot.
 */
#ifndef FLUCA_B200_H
#define FLUCA_B200_H

#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

#define FLUCA_B200_OK 0
#define FLUCA_B200_ERR_ARG 1
#define FLUCA_B200_ERR_CUDA 2
#define FLUCA_B200_ERR_NCCL 3
#define FLUCA_B200_ERR_DIVERGED 4 /* -> ns->reason = NS_DIVERGED_NONLINEAR_SOLVE (flucans.h:18) */
#define FLUCA_B200_ERR_NODEVICE 5
#define FLUCA_B200_ERR_INTERNAL 6

/* boundary condition types: numeric values of NSBoundaryConditionType (flucansbc.h:5-11) */
#define FLUCA_B200_BC_NONE 0
#define FLUCA_B200_BC_VELOCITY 1
#define FLUCA_B200_BC_PRESSURE_OUTLET 2
#define FLUCA_B200_BC_PERIODIC 3
#define FLUCA_B200_BC_SYMMETRY 4

/* solve modes (SURVEY.md section 7 "parity definition") */
#define FLUCA_B200_MODE_COUPLED 0    /* reference default: outer GMRES on the coupled system, PC = ABF (nssol.c:21-29) */
#define FLUCA_B200_MODE_FRACTIONAL 1 /* one ABF application = classical fractional step (abfpc.c:48-111) */

typedef struct fluca_b200_solver fluca_b200_solver;
typedef struct fluca_b200_comm   fluca_b200_comm;

typedef struct {
  int           dim;        /* 2 or 3 (MeshGetDimension) */
  int           n[3];       /* global cells (MeshCartGetGlobalSizes) */
  const double *xf[3];      /* n[d]+1 face coordinates per direction (MeshCartGetCoordinateArraysRead, slot PREV) */
  int           bc_type[6]; /* LEFT, RIGHT, DOWN, UP, BACK, FRONT (cart.c:564-591; ns->bcs[b].type) */
  double        rho, mu, dt; /* ns->rho, ns->mu, ns->dt (nsimpl.h:45-47) */
  int           k0, nzl;    /* z-slab [k0, k0+nzl) of this rank (MeshCartGetCorners with -cart_ranks_z P); 2-D: 0, 1 */
  /* solver options; 0 / 0.0 selects the default in brackets */
  int    mode;            /* FLUCA_B200_MODE_* [COUPLED] */
  double outer_rtol;      /* [1e-5]  -ns_ksp_rtol, nssol.c:24 */
  int    outer_maxit;     /* [100] */
  int    outer_restart;   /* [30]    GMRES restart (PETSc default) */
  double mom_rtol;        /* [1e-5]  -ns_abf_momentum_ksp_rtol */
  double schur_rtol;      /* [1e-5]  -ns_abf_schur_ksp_rtol */
  int    inner_maxit;     /* [500] */
  int    mg_nu1, mg_nu2;  /* [2, 2]  Jacobi pre / post sweeps of the pressure V-cycle */
  int    mg_coarse_sweeps; /* [40] */
  int    no_bcg_quirk;    /* 0: 3-D scales the outlet-gradient BC vector by 1 as cnlinearcart3d.c:2977 does; 1: by dt/rho */
  int    no_t_outlet_quirk; /* 0: 3-D forms operator T at an UPPER pressure outlet as cnlinearcart3d.c:1996,2055,2114 do (cell weights
                               -1/3, 4/3: the wall coordinate is read from the slot of the partial element); 1: as cnlinearcart2d.c:1391
                               and operator B do (-1/8, 9/8 on a uniform mesh) */
} fluca_b200_desc;

typedef struct {
  int    outer_its;    /* outer KSP iterations (-ns_ksp_monitor count) */
  int    mom_its;      /* total momentum BiCGStab iterations */
  int    schur_its;    /* total pressure Krylov iterations */
  int    abf_applies;  /* PCApply_ABF count */
  int    converged;
  double outer_rnorm0, outer_rnorm;
  int    nhist;
  double hist[128];    /* outer true-residual history, the analogue of -ns_ksp_monitor */
  long   launches;     /* kernels + device copies launched by this step */
  double mom_last_rel, schur_last_rel;
  int    inner_unconverged; /* inner solves of this step that ran into inner_maxit (PETSc's KSP inside PCApply_ABF would return
                               KSP_DIVERGED_ITS without an error too; the count makes a stalled inner solve visible) */
} fluca_b200_stats;

const char *fluca_b200_last_error(void);
/* 1 if the library was built as the CPU-only host-emulation test double (never the product build) */
int fluca_b200_is_host_emulation(void);

/* ---- communicator over the GPUs of one box (one process per GPU) ---- */
/* rank 0 calls unique_id and ships the bytes to the others (MPI_Bcast in the glue, torch.distributed in bench.py) */
int fluca_b200_comm_unique_id(void *out, int capacity, int *bytes);
int fluca_b200_comm_create_nccl(const void *unique_id, int bytes, int rank, int nranks, fluca_b200_comm **out);
/* host-callback flavour (MPI host without NCCL bootstrap; gloo in the CPU tests) */
typedef int (*fluca_b200_halo_fn)(void *ctx, const double *send_down, double *recv_down, const double *send_up, double *recv_up, long count, int periodic);
typedef int (*fluca_b200_allsum_fn)(void *ctx, double *vals, int n);
typedef int (*fluca_b200_allgather_fn)(void *ctx, const double *send, double *recv, long count);
int fluca_b200_comm_create_callbacks(int rank, int nranks, fluca_b200_halo_fn, fluca_b200_allsum_fn, fluca_b200_allgather_fn, void *ctx, fluca_b200_comm **out);
/* a communicator handed to fluca_b200_create is owned (and destroyed) by the solver */

/* ---- solver life cycle: NSCreate_<type> / ops->setup / ops->destroy ---- */
int fluca_b200_create(const fluca_b200_desc *desc, fluca_b200_comm *comm /* NULL = single GPU */, fluca_b200_solver **out);
int fluca_b200_destroy(fluca_b200_solver *s);

/* ---- PCABF variants: PCABFSetSchurComplementAinvType / PCABFSetUpperTriangularAinvType (flucans.h:99-107,
 * abfpc.c:300-318; options -ns_pc_abf_schur_ainv_type / -ns_pc_abf_upper_ainv_type, abfpc.c:246-247).  The approximation
 * of A^-1 used in the Schur complement S = D((-T) A~^-1 G~ - (-R)) (abfpc.c:151-170) and in the upper triangular factor
 * (abfpc.c:80-99): ID (default), the reciprocal diagonal of A, or the reciprocal row sums of A, rebuilt matrix-free with A
 * at every step.  Takes effect from the next fluca_b200_prepare_step / fluca_b200_step. */
#define FLUCA_B200_AINV_ID 0     /* PC_ABF_AINV_ID */
#define FLUCA_B200_AINV_DIAG 1   /* PC_ABF_AINV_DIAG */
#define FLUCA_B200_AINV_ROWSUM 2 /* PC_ABF_AINV_ROWSUM */
int fluca_b200_set_abf_ainv_types(fluca_b200_solver *s, int schur_type, int upper_type);

/* ---- inner residual histories: the analogue of -ns_abf_momentum_ksp_monitor / -ns_abf_schur_ksp_monitor on the two KSPs of PCABF
 * (abfpc.c:33-46,72,77).  fn is called on the host, from the thread that calls fluca_b200_step, once per residual norm the inner
 * solvers evaluate anyway (no extra reduction, no extra synchronisation): which = 0 momentum, 1 Schur complement; it = 0 for the
 * initial residual of a solve, then the iteration number; rnorm = 2-norm of the residual (all ranks see the same values).
 * fn = NULL switches it off.  The callback must not call into the library. */
typedef void (*fluca_b200_inner_monitor_fn)(void *ctx, int which, int it, double rnorm);
int fluca_b200_set_inner_monitor(fluca_b200_solver *s, fluca_b200_inner_monitor_fn fn, void *ctx);

/* ---- state: ns->sol sub-vectors Velocity / FaceNormalVelocity / Pressure + "PressureHalfStep" (cnlinear.c:54) ---- */
/* host pointers; any argument may be NULL to skip that field */
int fluca_b200_set_state(fluca_b200_solver *s, const double *v, const double *const U[3], const double *p, const double *phalf);
int fluca_b200_get_state(fluca_b200_solver *s, double *v, double *const U[3], double *p, double *phalf);

/* page-locked host memory for the caller's staging buffers (full-rate DMA for set_state / get_state / boundary planes) without
 * the caller needing the CUDA headers: the PETSc glue allocates its upload buffers here */
int fluca_b200_host_alloc(size_t bytes, void **ptr);
int fluca_b200_host_free(void *ptr);
/* ---- asynchronous solution views: what a monitor, NSViewSolution or the CGNS writer needs (NSViewSolution nssol.c:130-174,
 * VecView_Cart cartvec.c:4-25) without stalling the time loop.
 * stage_state  enqueues a copy of the current state into library-owned PINNED host buffers on a second CUDA stream, ordered
 *              after the work already submitted, and returns at once.  Steps issued afterwards run concurrently with the
 *              copy: the solver orders only its next overwrite of the copied fields after it.
 * staged_state blocks the HOST until that copy is complete and returns the buffers (compact layout of get_state; U[d] is
 *              NULL for d >= dim).  They stay valid, and unchanged, until the next stage_state or destroy. */
int fluca_b200_stage_state(fluca_b200_solver *s);
int fluca_b200_staged_state(fluca_b200_solver *s, const double **v, const double *U[3], const double **p, const double **phalf);

/* ---- boundary data: values of ns->bcs[b].velocity / .pressure (flucansbc.h:14) evaluated by the host at the
 * boundary-face centres of this rank's slab: x boundaries [nzl][ny], y boundaries [nzl][nx], z boundaries [ny][nx];
 * velocity: dim consecutive component blocks.
 * velocity slots: 0 = t^n, 1 = t^n + dt.   pressure slots: 0 = t_q (t^n at step 0, else t^n - dt/2), 1 = t^n + dt/2 */
int fluca_b200_set_boundary_velocity(fluca_b200_solver *s, int boundary, int slot, const double *values);
int fluca_b200_set_boundary_pressure(fluca_b200_solver *s, int boundary, int slot, const double *values);

/* ---- immersed boundary.  The reference advertises the method (README.md:14) but has no code
 * for it (THEORY_GUIDE.md:130-132 is a TODO), so these entry points replace no reference interface; they are what an
 * IBM-enabled NS type would add next to NSSetBoundaryCondition.  Every rank passes the same global list; a rank then works
 * only on the markers whose support touches its slab and exchanges the partial sums of the markers that straddle a slab face
 * with that neighbour (the "marker ownership exchange").  With several ranks set_markers, get_marker_forces and
 * ibm_interpolate are COLLECTIVE: every rank calls them, every rank gets the complete arrays.  X, Ud, F, Um: dim consecutive blocks of n doubles; dV: n doubles (Lagrangian volume weights).
 * delta_points: 4 (Peskin, default for 0) or 3 (Roma).  n = 0 removes the markers.
 * Coupling: direct forcing with an implicit predictor, added to the momentum right-hand side of the step. */
int fluca_b200_set_markers(fluca_b200_solver *s, long n, const double *X, const double *Ud, const double *dV, int delta_points);
/* multi-direct forcing: passes of (interpolate, spread) per step, each acting on the corrected predictor [1] */
int fluca_b200_set_ibm_iterations(fluca_b200_solver *s, int passes);
/* force of every marker on the fluid, rho (Ud - Um) dV / dt, and the interpolated predictor velocity Um of the last step */
int fluca_b200_get_marker_forces(fluca_b200_solver *s, double *F, double *Um);
/* ownership of this rank: info = {markers it works on, shared with the lower slab, shared with the upper slab, 1 if the
 * neighbour exchange is in use (0: every rank walks every marker and one allreduce completes the sums)} */
int fluca_b200_ibm_info(fluca_b200_solver *s, long info[4]);
/* operator-level (tests): Um = interpolation of the cell field v; f = spreading of Fm (f is overwritten) */
int fluca_b200_ibm_interpolate(fluca_b200_solver *s, const double *v, double *Um);
int fluca_b200_ibm_spread(fluca_b200_solver *s, const double *Fm, double *f);

/* ---- ops->step: one NSStep_CNLinear_Cart{2,3}d_Internal (cnlinearcart3d.c:2807-2863) ---- */
int fluca_b200_step(fluca_b200_solver *s, double t, int step_index, fluca_b200_stats *stats);

/* ---- operator-level entry points (parity tests; ops->formfunction / ops->formjacobian analogues) ---- */
/* sol0 <- sol, builds b = (r_mom, r_int, r_con) of NSFormFunction (cnlinearcart3d.c:2945-3043) */
int fluca_b200_prepare_step(fluca_b200_solver *s, double t, int step_index);
/* the right-hand side of the prepared step; also valid after fluca_b200_step, then it is the b the solve of that step saw
 * (immersed-boundary forcing added to r_mom; mean of r_con removed when no pressure outlet exists, nsbasic.c:133-144) */
int fluca_b200_get_rhs(fluca_b200_solver *s, double *rmom, double *const rint[3], double *rcon);
/* y = A x, A = I + dt C - (nu dt/2) L of NSFormJacobian(UPDATE) (cnlinearcart3d.c:2930-2941); needs prepare_step */
int fluca_b200_apply_momentum(fluca_b200_solver *s, const double *x, double *y);
/* y = S p, S = -(dt/rho) D Gst (abfpc.c:151-170); with a DIAG / ROWSUM Schur type S depends on A: needs prepare_step */
int fluca_b200_apply_schur(fluca_b200_solver *s, const double *p, double *y);
/* y = M x of the coupled 3x3 block system (MatNest J, nsbasic.c:203-207) */
int fluca_b200_apply_coupled(fluca_b200_solver *s, const double *xv, const double *const xU[3], const double *xp, double *yv, double *const yU[3], double *yp);
/* x = PCApply_ABF(b) (abfpc.c:48-111) */
int fluca_b200_apply_abf(fluca_b200_solver *s, const double *bv, const double *const bU[3], const double *bp, double *xv, double *const xU[3], double *xp, fluca_b200_stats *stats);

/* z = one multigrid V-cycle applied to r (the pressure preconditioner alone), for tests */
int fluca_b200_apply_vcycle(fluca_b200_solver *s, const double *r, double *z);

/* ---- device-resident access (HBM-resident benchmarking, GPU-side consumers) ---- */
/* copies the current state into / from a second device snapshot without touching the host */
int fluca_b200_snapshot_save(fluca_b200_solver *s);
int fluca_b200_snapshot_restore(fluca_b200_solver *s);
/* padded device layout of every field: index(i,j,kl) = i + px*(j + py*(kl+1)) */
int fluca_b200_device_layout(fluca_b200_solver *s, int *px, int *py, long *plane, long *nalloc);
/* names: "v0".."v2", "U0".."U2", "p", "phalf" -> device pointer (double*) of the live state */
int fluca_b200_device_field(fluca_b200_solver *s, const char *name, void **ptr);
/* CUDA stream the solver launches on (cudaStream_t), for event timing */
int fluca_b200_stream(fluca_b200_solver *s, void **stream);
/* total kernel / device-copy launches since creation */
long fluca_b200_launch_count(fluca_b200_solver *s);
/* algorithmic bytes of one step with the given iteration counts (SURVEY.md 8d model, actual padded sizes) */
double fluca_b200_step_model_bytes(fluca_b200_solver *s, const fluca_b200_stats *stats);

/* ---- live per-class kernel timing (CUDA event pairs on the solver stream around each launch) ---- */
#define FLUCA_B200_KT_MOMENTUM_APPLY 0 /* y = A x fused with two dot products */
#define FLUCA_B200_KT_MOMENTUM_VEC 1
#define FLUCA_B200_KT_POISSON_APPLY 2  /* q = P p fused with <p, q> */
#define FLUCA_B200_KT_POISSON_VEC 3
#define FLUCA_B200_KT_MG_SMOOTH 4      /* damped-Jacobi sweep, all levels */
#define FLUCA_B200_KT_MG_TRANSFER 5    /* residual+restriction, prolongation+correction */
#define FLUCA_B200_KT_RHS_PROJECT 6
#define FLUCA_B200_KT_OUTER 7
#define FLUCA_B200_KT_HALO 8
#define FLUCA_B200_KT_IBM 9          /* marker interpolation / spreading */
#define FLUCA_B200_KT_NCLASS 10
int fluca_b200_kernel_timing(fluca_b200_solver *s, int enable);
/* the model of fluca_b200_step_model_bytes split by kernel class (the roofline numerator of each class) */
int fluca_b200_step_model_bytes_split(fluca_b200_solver *s, const fluca_b200_stats *stats, double bytes[FLUCA_B200_KT_NCLASS]);
/* accumulated milliseconds and launch counts per class since the last reset */
int fluca_b200_kernel_times(fluca_b200_solver *s, double ms[FLUCA_B200_KT_NCLASS], long counts[FLUCA_B200_KT_NCLASS], int reset);

/* ---- kernel-level timing hooks for bench.py's roofline (times with CUDA events on the solver stream) ---- */
/* runs `reps` launches of the named kernel on scratch fields and returns the mean duration in ms and the
 * algorithmic bytes of one launch.  names: "momentum_apply", "poisson_apply", "mg_smooth", "vector_update" */
int fluca_b200_time_kernel(fluca_b200_solver *s, const char *name, int reps, double *ms, double *algorithmic_bytes);

/* ==== FlucaFD: finite-difference operator family (fluca/include/flucafd.h), stencil layer ====
 * Host-side entry points (no device needed): the composed stencil the reference's FlucaFDGetStencil returns at an output
 * point, from which a matrix-free device apply is generated.  One rank.  Names and argument meaning follow flucafd.h:
 *   grid_create          the DMStag facts FlucaFDSetUp reads (fdbasic.c:165-186): sizes, periodicity, stencil width and the
 *                        product coordinates (xf[d]: n[d]+1 faces; xc[d]: n[d] centres, NULL = face midpoints)
 *   *_create             FlucaFDDerivativeCreate / SumCreate / ScaleCreateConstant / ScaleCreateVector / CompositionCreate /
 *                        SecondOrderTVDCreate (flucafd.h:82-109); operands must be set up first, as in the reference
 *   set_locations        FlucaFDSetInputLocation + FlucaFDSetOutputLocation (flucafd.h:64-65; -flucafd_input_loc ...)
 *   set_boundary_condition  FlucaFDSetBoundaryConditions for one of LEFT, RIGHT, DOWN, UP, BACK, FRONT (flucafd.h:66)
 *   setup                FlucaFDSetUp (flucafd.h:59)
 *   get_stencil          FlucaFDGetStencil (flucafd.h:74): raw stencil, off-grid points removed per boundary condition
 * loc arguments are DMStagStencilLocation values (ELEMENT 14, LEFT 13, DOWN 11, BACK 5, DOWN_LEFT 10, BACK_LEFT 4,
 * BACK_DOWN 2, BACK_DOWN_LEFT 1); col.c < 0 marks a boundary value (-1 left ... -6 front, flucafd.h:44-50) or the constant
 * term (-7, flucafd.h:52).  Host fields are compact [nz+ez][ny+ey][nx+ex] arrays of one location (e = 1 on a non-periodic
 * face direction).  Errors: return code + fluca_b200_fd_last_error(). */
#define FLUCA_B200_FD_MAX_STENCIL 32 /* FLUCAFD_MAX_STENCIL_SIZE */
#define FLUCA_B200_FD_BC_NONE 0      /* FlucaFDBoundaryConditionType, flucafd.h:31-36 */
#define FLUCA_B200_FD_BC_DIRICHLET 1
#define FLUCA_B200_FD_BC_NEUMANN 2
typedef struct fluca_b200_fd_grid fluca_b200_fd_grid;
typedef struct fluca_b200_fd      fluca_b200_fd;
typedef struct {
  int i, j, k, loc, c; /* DMStagStencil */
} fluca_b200_fd_col;
const char *fluca_b200_fd_last_error(void);
int fluca_b200_fd_grid_create(int dim, const int n[3], const double *const xf[3], const double *const xc[3], const int periodic[3], int stencil_width, fluca_b200_fd_grid **out);
int fluca_b200_fd_grid_destroy(fluca_b200_fd_grid *grid);
int fluca_b200_fd_derivative_create(fluca_b200_fd_grid *grid, int dir, int deriv_order, int accu_order, int input_loc, int input_c, int output_loc, int output_c, fluca_b200_fd **out);
int fluca_b200_fd_sum_create(int n, fluca_b200_fd *const ops[], fluca_b200_fd **out);
int fluca_b200_fd_scale_create_constant(fluca_b200_fd *operand, double constant, fluca_b200_fd **out);
int fluca_b200_fd_scale_create_vector(fluca_b200_fd *operand, const double *field, int vec_loc, int vec_c, fluca_b200_fd **out);
int fluca_b200_fd_composition_create(fluca_b200_fd *inner, fluca_b200_fd *outer, fluca_b200_fd **out);
int fluca_b200_fd_tvd_create(fluca_b200_fd_grid *grid, int dir, int input_c, int output_c, fluca_b200_fd **out);
int fluca_b200_fd_tvd_set_limiter(fluca_b200_fd *fd, const char *name);              /* FlucaFDSecondOrderTVDSetLimiter */
int fluca_b200_fd_tvd_set_velocity(fluca_b200_fd *fd, const double *face_velocity);  /* FlucaFDSecondOrderTVDSetVelocity */
int fluca_b200_fd_tvd_set_current_solution(fluca_b200_fd *fd, const double *phi);    /* FlucaFDSecondOrderTVDSetCurrentSolution */
int fluca_b200_fd_set_locations(fluca_b200_fd *fd, int input_loc, int input_c, int output_loc, int output_c);
int fluca_b200_fd_set_boundary_condition(fluca_b200_fd *fd, int boundary, int type, double value);
int fluca_b200_fd_setup(fluca_b200_fd *fd);
int fluca_b200_fd_get_stencil(fluca_b200_fd *fd, int i, int j, int k, int *ncols, fluca_b200_fd_col col[FLUCA_B200_FD_MAX_STENCIL], double v[FLUCA_B200_FD_MAX_STENCIL]);
/* FlucaFDApply (flucafd.h:75, fdapply.c:47-121) as a matrix-free device kernel generated from the stencil layer; v1 covers
 * derivative / sum / constant scale / composition on uniform product coordinates and rejects anything else with an error
 * (csrc/fd.cu) -- v2: those fall to the ASSEMBLED apply (every point's stencil evaluated on the host into an ELL table, one
 * thread per output point on the device; re-assembled when a field the tree depends on changed), so vector scale, second-order
 * TVD and non-uniform coordinates are applied too.  apply_inputs reports which input fields (location, component) the composed operator reads, in the order
 * apply expects them; inputs / output are host arrays in the compact layout above (copied to and from the device inside).
 * Needs a CUDA device (FLUCA_B200_ERR_NODEVICE otherwise). */
int fluca_b200_fd_apply_inputs(fluca_b200_fd *fd, int *ninputs, int loc[4], int c[4]);
int fluca_b200_fd_apply(fluca_b200_fd *fd, int ninputs, const double *const inputs[], double *output);
/* device-resident form: device pointers, asynchronous launch on the operator's own stream (cudaStream_t through
 * fluca_b200_fd_stream, for event timing); fluca_b200_fd_sync waits for it.  The kernel's tables are built at the first
 * apply and rebuilt after set_locations / set_boundary_condition. */
int fluca_b200_fd_apply_device(fluca_b200_fd *fd, int ninputs, const double *const dev_inputs[], double *dev_output);
int fluca_b200_fd_stream(fluca_b200_fd *fd, void **stream);
int fluca_b200_fd_sync(fluca_b200_fd *fd);
/* FlucaFDGetOperator (flucafd.h, fdapply.c:123-180): the operator's matrix as host CSR over the output points (row = i + nx (j + ny k)
 * of the output location), interior stencil points only -- boundary and constant terms are left out, as the reference leaves
 * them out of the Mat.  Call with rowptr = NULL for the sizes, then with arrays of nrows + 1 / nnz / nnz entries. */
int fluca_b200_fd_get_operator(fluca_b200_fd *fd, long *nrows, long *nnz, long *rowptr, fluca_b200_fd_col *cols, double *vals);
int fluca_b200_fd_destroy(fluca_b200_fd *fd);

#ifdef __cplusplus
}
#endif
#endif
