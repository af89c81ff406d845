/*
 * tests/c/ns_b200_glue_driver.c -- TEST INFRASTRUCTURE: an application in the style of the reference's NS drivers
 * (fluca/tests/cavity_flow/cavity_flow_2d.c, cavity_flow_3d.c, fluca/tests/taylor_green_vortex/taylor_green_vortex.c) that selects
 * the NS type "b200" of glue/nsb200.c with -ns_type and runs it through the NS API.  PETSc is absent from this image, so it links
 * with the single-rank functional model of the API in tests/petsc_stub/petsc_fluca_mock.c; the numerical library underneath is the
 * real one (CUDA on the B200, the host-emulation build on a CPU).  tests/test_glue_mock.py compares what lands in ns->sol -- read
 * back here through DMGlobalToLocal + DMStagVecGetArrayRead + DMStagGetLocationSlot, as the reference's drivers read it -- with
 * the oracle.
 *
 *   usage: ns_b200_glue_driver case=<name> out=<file> [n=a,b[,c]] [steps=K] [steps2=K] [scenario=<s>] [stretch=x] [pout=x] [init=zero|smooth]
 *                              [-petsc_option[=value]] ...
 *   cases (fluca_b200/workloads.py): cavity2d cavity3d cavity3d_full tgv tgv_periodic channel2d channel2d_t channel3d channel3d_pz
 *   scenarios: plain | formfunction | edit | restart | stage;  markers=N: an immersed sphere of N Fibonacci-lattice markers (D = 1.2 at
 *   (0.1, 0, 0.05), 4-point delta) through NSB200SetMarkers, marker forces through NSB200GetMarkerForces
 *   out: float64 [dim n0 n1 n2 nstate nf] + v (component-major, then z, y, x) U_x U_y [U_z] p phalf [+ the same blocks of f]
 *        [+ with markers: F (dim x N) and U_m (dim x N)]
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "petsc_fluca_mock.h"

#define CHK(call) \
  do { \
    PetscErrorCode e_ = (call); \
    if (e_) { \
      fprintf(stderr, "%s:%d: %s failed (%d)\n", __FILE__, __LINE__, #call, (int)e_); \
      exit(1); \
    } \
  } while (0)

static long bc_calls = 0;

/* ---- boundary callbacks (flucansbc.h:14), the functions of fluca_b200/workloads.py */
static PetscErrorCode bc_const(PetscInt dim, PetscReal t, const PetscReal x[], PetscScalar val[], void *ctx)
{
  PetscInt c;
  (void)t, (void)x;
  ++bc_calls;
  for (c = 0; c < dim; ++c) val[c] = ((const double *)ctx)[c];
  return PETSC_SUCCESS;
}
static PetscErrorCode bc_const_pressure(PetscInt dim, PetscReal t, const PetscReal x[], PetscScalar val[], void *ctx)
{
  (void)dim, (void)t, (void)x;
  ++bc_calls;
  val[0] = *(const double *)ctx;
  return PETSC_SUCCESS;
}
static PetscErrorCode bc_tgv(PetscInt dim, PetscReal t, const PetscReal x[], PetscScalar val[], void *ctx)
{
  const double nu = *(const double *)ctx, e = exp(-2. * nu * t); /* taylor_green_vortex.c:13-22 */
  (void)dim;
  ++bc_calls;
  val[0] = sin(x[0]) * cos(x[1]) * e;
  val[1] = -cos(x[0]) * sin(x[1]) * e;
  return PETSC_SUCCESS;
}
static int time_dependent = 0;
static PetscErrorCode bc_inflow2d(PetscInt dim, PetscReal t, const PetscReal x[], PetscScalar val[], void *ctx)
{
  const double a = 1. + (time_dependent ? 0.1 * sin(3. * t) : 0.);
  (void)dim, (void)ctx;
  ++bc_calls;
  val[0] = a * (1. + 0.2 * cos(2. * M_PI * x[1] / 4.));
  val[1] = 0.;
  return PETSC_SUCCESS;
}
static PetscErrorCode bc_outlet2d(PetscInt dim, PetscReal t, const PetscReal x[], PetscScalar val[], void *ctx)
{
  const double pout = *(const double *)ctx;
  (void)dim;
  ++bc_calls;
  val[0] = pout * (1. + (time_dependent ? 0.5 * sin(2. * t) : 0.)) * (1. + 0.1 * x[1]);
  return PETSC_SUCCESS;
}

/* ---- the case table */
typedef struct {
  int                 dim, n[3];
  double              lo[3], hi[3], rho, mu, dt, stretch;
  NSBoundaryCondition bc[6];
  int                 init_tgv;
} Case;
static double zero3[3] = {0., 0., 0.}, lid3[3] = {1., 0., 0.}, inflow3[3] = {1., 0., 0.}, tgv_nu, pout = 0.;

static NSBoundaryCondition vel(NSBoundaryConditionFunction f, void *ctx)
{
  NSBoundaryCondition bc = {.type = NS_BC_VELOCITY, .velocity = f, .ctx_velocity = ctx};
  return bc;
}
static NSBoundaryCondition outlet(NSBoundaryConditionFunction f, void *ctx)
{
  NSBoundaryCondition bc = {.type = NS_BC_PRESSURE_OUTLET, .pressure = f, .ctx_pressure = ctx};
  return bc;
}
static NSBoundaryCondition plain(NSBoundaryConditionType t)
{
  NSBoundaryCondition bc = {.type = t};
  return bc;
}
static int make_case(const char *name, const int *n_arg, int n_given, Case *c)
{
  int b, d;
  memset(c, 0, sizeof(*c));
  c->rho = 1.;
  if (!strcmp(name, "cavity2d")) {
    const int n = n_given ? n_arg[0] : 16;
    c->dim = 2, c->n[0] = c->n[1] = n, c->hi[0] = c->hi[1] = 1., c->mu = 1. / 100., c->dt = 0.5 / n;
    for (b = 0; b < 3; ++b) c->bc[b] = vel(bc_const, zero3);
    c->bc[3] = vel(bc_const, lid3);
  } else if (!strcmp(name, "cavity3d") || !strcmp(name, "cavity3d_full")) {
    const int full = !strcmp(name, "cavity3d_full");
    c->dim = 3;
    for (d = 0; d < 3; ++d) c->n[d] = n_given ? n_arg[d] : (d == 2 && !full ? 4 : 8), c->hi[d] = 1.;
    if (!full) c->hi[2] = 0.5;
    c->mu = full ? 1. / 400. : 1. / 100., c->dt = 0.5 / c->n[0];
    for (b = 0; b < 6; ++b) c->bc[b] = vel(bc_const, zero3);
    c->bc[3] = vel(bc_const, lid3);
    if (!full) c->bc[4] = plain(NS_BC_SYMMETRY); /* cavity_flow_3d.c:51-77 */
  } else if (!strcmp(name, "tgv") || !strcmp(name, "tgv_periodic")) {
    const int per = !strcmp(name, "tgv_periodic"), n = n_given ? n_arg[0] : 8;
    c->dim = 2, c->n[0] = c->n[1] = n, c->hi[0] = c->hi[1] = 2. * M_PI, c->mu = 1., c->dt = 0.1, c->init_tgv = 1;
    tgv_nu = c->mu / c->rho;
    for (b = 0; b < 4; ++b) {
      c->bc[b] = vel(bc_tgv, &tgv_nu);
      if (per) c->bc[b].type = NS_BC_PERIODIC;
    }
  } else if (!strcmp(name, "channel2d") || !strcmp(name, "channel2d_t")) {
    time_dependent = !strcmp(name, "channel2d_t");
    c->dim = 2, c->n[0] = n_given ? n_arg[0] : 24, c->n[1] = n_given ? n_arg[1] : 12;
    c->lo[0] = -2., c->lo[1] = -2., c->hi[0] = 6., c->hi[1] = 2., c->mu = 1. / 100., c->dt = 0.5 * 8. / c->n[0];
    c->bc[0] = vel(bc_inflow2d, NULL), c->bc[1] = outlet(bc_outlet2d, &pout);
    c->bc[2] = c->bc[3] = plain(NS_BC_SYMMETRY);
  } else if (!strcmp(name, "channel3d") || !strcmp(name, "channel3d_pz")) {
    const int pz = !strcmp(name, "channel3d_pz");
    c->dim = 3;
    c->n[0] = n_given ? n_arg[0] : 12, c->n[1] = n_given ? n_arg[1] : 8, c->n[2] = n_given ? n_arg[2] : 8;
    for (d = 0; d < 3; ++d) c->lo[d] = -2., c->hi[d] = d == 0 ? 4. : 2.;
    c->mu = 1. / 300., c->dt = 0.5 * 6. / c->n[0];
    c->bc[0] = vel(bc_const, inflow3), c->bc[1] = outlet(bc_const_pressure, &pout);
    c->bc[2] = c->bc[3] = plain(NS_BC_SYMMETRY);
    c->bc[4] = c->bc[5] = plain(pz ? NS_BC_PERIODIC : NS_BC_SYMMETRY);
  } else return 1;
  return 0;
}

/* ---- fields through the DMStag API */
typedef double (*PointFn)(int comp, const double x[3], void *ctx);
static double f_tgv_v(int c, const double x[3], void *ctx) { return (void)ctx, c == 0 ? sin(x[0]) * cos(x[1]) : -cos(x[0]) * sin(x[1]); }
static double f_tgv_p(int c, const double x[3], void *ctx) { return (void)c, *(double *)ctx / 4. * (cos(2. * x[0]) + cos(2. * x[1])); }
/* the "smooth" user edit: the same closed forms in tests/test_glue_mock.py */
static double f_smooth_v(int c, const double x[3], void *ctx) { return (void)ctx, 0.5 * sin(1.3 * x[0] + 0.7 * x[1] + 0.5 * x[2] + c); }
static double f_smooth_U(int c, const double x[3], void *ctx) { return (void)ctx, 0.5 * sin(0.8 * x[0] + 1.2 * x[1] + 0.6 * x[2] + 2. + c); }
static double f_smooth_p(int c, const double x[3], void *ctx) { return (void)c, (void)ctx, cos(0.9 * x[0] - 1.1 * x[1] + 0.3 * x[2]); }

/* visit the entries (loc, comp) of a global vector: set them from fn (fn != NULL) or append them to out, in z, y, x order.  The
 * number of points per direction comes from DMStagGetCorners: the extra face layer of a non-periodic direction included. */
static size_t visit(Mesh mesh, DM dm, Vec g, DMStagStencilLocation loc, int facedir, int comp, PointFn fn, void *ctx, double *out)
{
  PetscInt            dim, x, y, z, m, n, p, ex, ey, ez, slot, i, j, k, iprev, ielem;
  const PetscScalar **ax, **ay, **az = NULL;
  Vec                 l;
  size_t              cnt = 0;
  (void)mesh;
  CHK(DMGetDimension(dm, &dim));
  CHK(DMStagGetCorners(dm, &x, &y, &z, &m, &n, &p, &ex, &ey, &ez));
  CHK(DMStagGetLocationSlot(dm, loc, comp, &slot));
  CHK(DMStagGetProductCoordinateArraysRead(dm, &ax, &ay, &az));
  CHK(DMStagGetProductCoordinateLocationSlot(dm, DMSTAG_LEFT, &iprev));
  CHK(DMStagGetProductCoordinateLocationSlot(dm, DMSTAG_ELEMENT, &ielem));
  if (dim == 2) z = 0, p = 1, ez = 0;
  m += facedir == 0 ? ex : 0, n += facedir == 1 ? ey : 0, p += facedir == 2 ? ez : 0;
  CHK(DMGetLocalVector(dm, &l));
  CHK(DMGlobalToLocal(dm, g, INSERT_VALUES, l));
  {
    PetscScalar ***a2 = NULL, ****a3 = NULL;
    if (dim == 2) CHK(DMStagVecGetArray(dm, l, &a2));
    else CHK(DMStagVecGetArray(dm, l, &a3));
    for (k = z; k < z + p; ++k)
      for (j = y; j < y + n; ++j)
        for (i = x; i < x + m; ++i) {
          double *e = dim == 2 ? &a2[j][i][slot] : &a3[k][j][i][slot];
          if (fn) {
            const double xx[3] = {ax[i][facedir == 0 ? iprev : ielem], ay[j][facedir == 1 ? iprev : ielem], dim == 3 ? az[k][facedir == 2 ? iprev : ielem] : 0.};
            *e = fn(facedir >= 0 ? facedir : comp, xx, ctx);
          } else out[cnt] = *e;
          ++cnt;
        }
    if (dim == 2) CHK(DMStagVecRestoreArray(dm, l, &a2));
    else CHK(DMStagVecRestoreArray(dm, l, &a3));
  }
  if (fn) CHK(DMLocalToGlobal(dm, l, INSERT_VALUES, g));
  CHK(DMRestoreLocalVector(dm, &l));
  CHK(DMStagRestoreProductCoordinateArraysRead(dm, &ax, &ay, &az));
  return cnt;
}
static const DMStagStencilLocation face_loc[3] = {DMSTAG_LEFT, DMSTAG_DOWN, DMSTAG_BACK};

/* set (fv != NULL) or read the three fields of a nest laid out like ns->sol */
static size_t nest_fields(NS ns, Vec nest, PointFn fv, PointFn fU, PointFn fp, void *ctx, double *out)
{
  Mesh        mesh = ns->mesh;
  DM          sdm, vdm, Sdm;
  IS          is[3];
  Vec         v, V, p;
  PetscInt    dim, d;
  size_t      cnt = 0;
  const char *names[3] = {NS_FIELD_VELOCITY, NS_FIELD_FACE_NORMAL_VELOCITY, NS_FIELD_PRESSURE};
  CHK(MeshGetDimension(mesh, &dim));
  CHK(MeshGetDM(mesh, MESH_DM_SCALAR, &sdm));
  CHK(MeshGetDM(mesh, MESH_DM_VECTOR, &vdm));
  CHK(MeshGetDM(mesh, MESH_DM_STAG_SCALAR, &Sdm));
  for (d = 0; d < 3; ++d) CHK(NSGetField(ns, names[d], NULL, NULL, &is[d]));
  CHK(VecGetSubVector(nest, is[0], &v));
  for (d = 0; d < dim; ++d) cnt += visit(mesh, vdm, v, DMSTAG_ELEMENT, -1, (int)d, fv, ctx, out ? out + cnt : NULL);
  CHK(VecRestoreSubVector(nest, is[0], &v));
  CHK(VecGetSubVector(nest, is[1], &V));
  for (d = 0; d < dim; ++d) cnt += visit(mesh, Sdm, V, face_loc[d], (int)d, 0, fU, ctx, out ? out + cnt : NULL);
  CHK(VecRestoreSubVector(nest, is[1], &V));
  CHK(VecGetSubVector(nest, is[2], &p));
  cnt += visit(mesh, sdm, p, DMSTAG_ELEMENT, -1, 0, fp, ctx, out ? out + cnt : NULL);
  CHK(VecRestoreSubVector(nest, is[2], &p));
  return cnt;
}
/* p-half lives in the type's private data; what an application can see of it is what NSViewSolution writes */
static size_t read_phalf(NS ns, double *out)
{
  PetscViewer store;
  Vec         ph;
  DM          sdm;
  size_t      cnt;
  CHK(MockViewerStoreOpen(&store));
  CHK(NSViewSolution(ns, store));
  CHK(MeshGetDM(ns->mesh, MESH_DM_SCALAR, &sdm));
  CHK(MeshCreateGlobalVector(ns->mesh, MESH_DM_SCALAR, &ph));
  CHK(PetscObjectSetName((PetscObject)ph, "PressureHalfStep")); /* cnlinear.c:54 */
  CHK(FlucaVecLoad(ph, store));
  cnt = visit(ns->mesh, sdm, ph, DMSTAG_ELEMENT, -1, 0, NULL, NULL, out);
  CHK(VecDestroy(&ph));
  CHK(PetscViewerDestroy(&store));
  return cnt;
}

static NS new_ns(const Case *c, Mesh mesh)
{
  NS  ns;
  int b;
  CHK(NSCreate(0, &ns));
  CHK(NSSetMesh(ns, mesh));
  CHK(NSSetDensity(ns, c->rho));
  CHK(NSSetViscosity(ns, c->mu));
  CHK(NSSetTimeStepSize(ns, c->dt));
  for (b = 0; b < 2 * c->dim; ++b) CHK(NSSetBoundaryCondition(ns, b, c->bc[b]));
  CHK(NSSetFromOptions(ns)); /* -ns_type b200 selects the type, as for the reference's apps (nsopts.c:179-181) */
  CHK(NSSetUp(ns));
  return ns;
}

int main(int argc, char **argv)
{
  const char *casename = "cavity2d", *outname = NULL, *scenario = "plain", *init = "zero";
  int         n_arg[3] = {0, 0, 0}, n_given = 0, steps = 2, steps2 = 1, a, d, k, nmark = 0;
  double      stretch = 0.;
  Case        c;
  Mesh        mesh;
  NS          ns;
  PetscViewer ascii;

  CHK(MockOptionsSetValue("-ns_type", "b200"));
  for (a = 1; a < argc; ++a) {
    char *eq = strchr(argv[a], '=');
    if (argv[a][0] == '-') {
      if (eq) *eq = 0;
      CHK(MockOptionsSetValue(argv[a], eq ? eq + 1 : ""));
    } else if (!strncmp(argv[a], "case=", 5)) casename = argv[a] + 5;
    else if (!strncmp(argv[a], "out=", 4)) outname = argv[a] + 4;
    else if (!strncmp(argv[a], "scenario=", 9)) scenario = argv[a] + 9;
    else if (!strncmp(argv[a], "init=", 5)) init = argv[a] + 5;
    else if (!strncmp(argv[a], "steps=", 6)) steps = atoi(argv[a] + 6);
    else if (!strncmp(argv[a], "steps2=", 7)) steps2 = atoi(argv[a] + 7);
    else if (!strncmp(argv[a], "markers=", 8)) nmark = atoi(argv[a] + 8);
    else if (!strncmp(argv[a], "stretch=", 8)) stretch = atof(argv[a] + 8);
    else if (!strncmp(argv[a], "pout=", 5)) pout = atof(argv[a] + 5);
    else if (!strncmp(argv[a], "nestbump=", 9)) MockSetNestRestoreBumpsState(atoi(argv[a] + 9));
    else if (!strncmp(argv[a], "n=", 2)) n_given = sscanf(argv[a] + 2, "%d,%d,%d", &n_arg[0], &n_arg[1], &n_arg[2]);
    else {
      fprintf(stderr, "unknown argument %s\n", argv[a]);
      return 2;
    }
  }
  if (n_given == 1) n_arg[1] = n_arg[2] = n_arg[0];
  if (make_case(casename, n_arg, n_given, &c) || !outname) {
    fprintf(stderr, "unknown case or no out=\n");
    return 2;
  }
  c.stretch = stretch;

  CHK(PetscDLLibraryRegister_fluca_nsb200()); /* what -dll_append libfluca_nsb200.so triggers: NSRegister("b200", NSCreate_B200) */

  { /* MeshCartCreate + coordinates (fluca_b200/workloads.py Case.faces: uniform, or a smooth stretching that keeps the end points) */
    PetscInt  N[3];
    PetscBool per[3];
    double   *xf[3] = {NULL, NULL, NULL};
    for (d = 0; d < c.dim; ++d) {
      N[d]   = c.n[d];
      per[d] = c.bc[2 * d].type == NS_BC_PERIODIC ? PETSC_TRUE : PETSC_FALSE;
      xf[d]  = malloc(sizeof(double) * (size_t)(c.n[d] + 1));
      for (k = 0; k <= c.n[d]; ++k) {
        double s = (double)k / c.n[d];
        if (c.stretch != 0. && !per[d]) s += c.stretch * sin(2. * M_PI * s) / (2. * M_PI);
        xf[d][k] = c.lo[d] + (c.hi[d] - c.lo[d]) * s;
      }
    }
    CHK(MockMeshCartCreate(c.dim, N, per, (const double *const *)xf, &mesh));
    for (d = 0; d < c.dim; ++d) free(xf[d]);
  }
  ns = new_ns(&c, mesh);
  CHK(MockViewerASCIIOpen(stdout, &ascii));

  /* initial condition through the host Vec, like cavity_flow_2d.c:74-75 and taylor_green_vortex.c:113-178 */
  {
    Vec sol;
    CHK(NSGetSolution(ns, &sol));
    CHK(VecSet(sol, 0.));
    if (c.init_tgv) nest_fields(ns, sol, f_tgv_v, f_tgv_v, f_tgv_p, &c.rho, NULL);
    else if (!strcmp(init, "smooth")) nest_fields(ns, sol, f_smooth_v, f_smooth_U, f_smooth_p, NULL, NULL);
  }

  size_t  nstate = 0, nf = 0;
  double *state = NULL, *fbuf = NULL, *mark = NULL;
  if (nmark > 0) { /* fluca_b200/workloads.py sphere_markers: Fibonacci lattice, volume weight = area / N x h; arrays are [component][marker] */
    const double D = 1.2, ctr[3] = {0.1, 0., 0.05}, h = 4.0 / c.n[1];
    double      *X = malloc(sizeof(double) * 3 * (size_t)nmark), *Ud = calloc(3 * (size_t)nmark, sizeof(double)), *dV = malloc(sizeof(double) * (size_t)nmark);
    if (c.dim != 3) {
      fprintf(stderr, "markers= needs a 3-D case\n");
      return 2;
    }
    for (k = 0; k < nmark; ++k) {
      const double kk = k + 0.5, z = 1. - 2. * kk / nmark, r = sqrt(fmax(0., 1. - z * z)), ph = M_PI * (1. + sqrt(5.)) * kk;
      X[k] = ctr[0] + 0.5 * D * r * cos(ph), X[nmark + k] = ctr[1] + 0.5 * D * r * sin(ph), X[2 * nmark + k] = ctr[2] + 0.5 * D * z;
      dV[k] = M_PI * D * D / nmark * h;
    }
    CHK(NSB200SetMarkers(ns, nmark, X, Ud, dV, 4));
    free(X), free(Ud), free(dV);
    mark = malloc(sizeof(double) * 6 * (size_t)nmark);
  }
  /* sizes: count once with a scratch buffer large enough for any field set */
  {
    size_t cells = (size_t)(c.n[0] + 1) * (c.n[1] + 1) * (c.dim == 3 ? c.n[2] + 1 : 1);
    state = malloc(sizeof(double) * cells * 9);
    fbuf  = malloc(sizeof(double) * cells * 8);
  }

  for (k = 0; k < steps; ++k) {
    if (!strcmp(scenario, "stage") && k == steps - 1) CHK(NSB200StageSolution(ns)); /* state k starts flowing to the host; the step below overlaps it */
    CHK(NSStep(ns));
    printf("%d NS dt %g time %g reason %d\n", (int)ns->step, ns->dt, ns->t, (int)ns->reason);
  }

  if (!strcmp(scenario, "formfunction")) {
    Vec x, f;
    CHK(MockNSCreateVecs(ns, &x, &f));
    CHK(NSFormFunction(ns, x, f)); /* nsbasic.c:316-323 */
    nf = nest_fields(ns, f, NULL, NULL, NULL, NULL, fbuf);
    CHK(VecDestroy(&x));
    CHK(VecDestroy(&f));
    for (k = 0; k < steps2; ++k) CHK(NSStep(ns));
  } else if (!strcmp(scenario, "edit")) {
    Vec sol; /* a user edit of ns->sol between two steps (v, U, p; p-half keeps its value): the type has to notice and upload */
    CHK(NSGetSolution(ns, &sol));
    nest_fields(ns, sol, f_smooth_v, f_smooth_U, f_smooth_p, NULL, NULL);
    for (k = 0; k < steps2; ++k) CHK(NSStep(ns));
  } else if (!strcmp(scenario, "restart")) {
    PetscViewer store; /* write, destroy everything, read into a new object, continue (nssol.c:130-204, cnlinear.c:146-162) */
    CHK(MockViewerStoreOpen(&store));
    CHK(NSViewSolution(ns, store));
    CHK(NSDestroy(&ns));
    ns = new_ns(&c, mesh);
    CHK(NSLoadSolution(ns, store));
    CHK(PetscViewerDestroy(&store));
    for (k = 0; k < steps2; ++k) CHK(NSStep(ns));
  } else if (!strcmp(scenario, "stage")) {
    CHK(NSB200SyncSolution(ns)); /* ns->sol = the state staged before the last step */
  } else if (strcmp(scenario, "plain")) {
    fprintf(stderr, "unknown scenario %s\n", scenario);
    return 2;
  }

  CHK(NSView(ns, ascii));
  {
    Vec sol;
    CHK(NSGetSolution(ns, &sol));
    /* with -ns_b200_sync_interval 0 the host copy is refreshed by the first observer: NSViewSolution (inside read_phalf) */
    (void)read_phalf(ns, state);
    nstate = nest_fields(ns, sol, NULL, NULL, NULL, NULL, state);
    nstate += read_phalf(ns, state + nstate);
  }
  {
    FILE  *f      = fopen(outname, "wb");
    double hdr[6] = {c.dim, c.n[0], c.n[1], c.dim == 3 ? c.n[2] : 1, (double)nstate, (double)nf};
    if (!f) return 3;
    fwrite(hdr, sizeof(double), 6, f);
    fwrite(state, sizeof(double), nstate, f);
    fwrite(fbuf, sizeof(double), nf, f);
    if (nmark > 0) {
      CHK(NSB200GetMarkerForces(ns, mark, mark + 3 * (size_t)nmark));
      fwrite(mark, sizeof(double), 6 * (size_t)nmark, f);
    }
    fclose(f);
  }
  printf("STEP %d TIME %.17g\n", (int)ns->step, ns->t);
  printf("BCCALLS %ld\n", bc_calls);
  printf("G2L %ld\n", mock_ndm_global_to_local);
  CHK(PetscViewerDestroy(&ascii));
  CHK(NSDestroy(&ns));
  CHK(MeshDestroy(&mesh));
  free(state), free(fbuf), free(mark);
  printf("LIVE %ld\n", MockLiveAllocations());
  return 0;
}
