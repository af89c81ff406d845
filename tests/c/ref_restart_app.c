/*
 * tests/c/ref_restart_app.c -- TEST INFRASTRUCTURE: an application written against the REFERENCE's public API (flucameshcart.h,
 * flucans.h, flucaviewer.h; compiled with the reference's own headers and linked with the reference's own libraries on the PETSc
 * model, oracle/Makefile target ref_app) that exercises what the reference's drivers do not: writing the solution with
 * NSViewSolution, and continuing another run from it with NSLoadSolution (nssol.c:130-204).  The case is the one of
 * fluca/tests/cavity_flow/cavity_flow_3d.c.  The NS type comes from the options (-ns_type, -dll_append), so a file written by the
 * reference's cnlinear can be continued by b200 and the other way round: the extra state of both types travels under the name
 * "PressureHalfStep" (cnlinear.c:54,146-162).
 *
 *   ref_restart_app write <file> [options]   zero state, NSSolve to -ns_max_steps, NSViewSolution into <file>
 *   ref_restart_app read  <file> [options]   NSLoadSolution from <file>, NSSolve on to -ns_max_steps, NSViewSolution into <file>.out
 *   -stretch s : a non-uniform mesh through the reference's public coordinate API (MeshCartGetCoordinateArrays / Restore, cart.c:467-502):
 *                face k/n moves to k/n + s sin(2 pi k/n) / (2 pi) in every direction
 */
#include <flucameshcart.h>
#include <flucans.h>
#include <flucasys.h>
#include <flucaviewer.h>
#include <math.h>
#include <string.h>

static PetscErrorCode wall_velocity(PetscInt dim, PetscReal t, const PetscReal x[], PetscScalar val[], void *ctx)
{
  val[0] = val[1] = val[2] = 0.;
  return PETSC_SUCCESS;
}
static PetscErrorCode moving_wall_velocity(PetscInt dim, PetscReal t, const PetscReal x[], PetscScalar val[], void *ctx)
{
  val[0] = 1., val[1] = val[2] = 0.;
  return PETSC_SUCCESS;
}

int main(int argc, char **argv)
{
  Mesh        mesh;
  NS          ns;
  Vec         sol;
  PetscViewer viewer;
  PetscInt    b, step;
  PetscReal   t;
  char        out[4096];
  const int   reading = argc > 2 && !strcmp(argv[1], "read");

  PetscCall(FlucaInitialize(&argc, &argv, NULL, NULL));
  PetscCheck(argc > 2 && (reading || !strcmp(argv[1], "write")), PETSC_COMM_WORLD, PETSC_ERR_ARG_WRONG, "usage: ref_restart_app write|read <file> [options]");
  PetscCall(MeshCartCreate3d(PETSC_COMM_WORLD, MESHCART_BOUNDARY_NONE, MESHCART_BOUNDARY_NONE, MESHCART_BOUNDARY_NONE, 8, 8, 4, PETSC_DECIDE, PETSC_DECIDE, PETSC_DECIDE, NULL, NULL, NULL, &mesh));
  PetscCall(MeshSetFromOptions(mesh));
  PetscCall(MeshSetUp(mesh));
  PetscCall(MeshCartSetUniformCoordinates(mesh, 0., 1., 0., 1., 0., 0.5));
  {
    PetscReal     stretch = 0.;
    PetscScalar **a[3];
    PetscInt      N[3], iprev, d, i;
    PetscCall(PetscOptionsGetReal(NULL, NULL, "-stretch", &stretch, NULL));
    if (stretch != 0.) {
      const double len[3] = {1., 1., 0.5};
      PetscCall(MeshCartGetGlobalSizes(mesh, &N[0], &N[1], &N[2]));
      PetscCall(MeshCartGetCoordinateLocationSlot(mesh, MESHCART_PREV, &iprev));
      PetscCall(MeshCartGetCoordinateArrays(mesh, &a[0], &a[1], &a[2]));
      for (d = 0; d < 3; ++d)
        for (i = 0; i <= N[d]; ++i) {
          const double s = (double)i / N[d];
          a[d][i][iprev] = len[d] * (s + stretch * sin(2. * M_PI * s) / (2. * M_PI));
        }
      PetscCall(MeshCartRestoreCoordinateArrays(mesh, &a[0], &a[1], &a[2])); /* recomputes the centres */
    }
  }
  PetscCall(NSCreate(PETSC_COMM_WORLD, &ns));
  PetscCall(NSSetType(ns, NSCNLINEAR));
  PetscCall(NSSetMesh(ns, mesh));
  PetscCall(NSSetDensity(ns, 1.));
  PetscCall(NSSetViscosity(ns, 0.01));
  for (b = 0; b < 6; ++b) {
    NSBoundaryCondition bc = {.type = NS_BC_VELOCITY, .velocity = b == 3 ? moving_wall_velocity : wall_velocity};
    if (b == 4) bc.type = NS_BC_SYMMETRY, bc.velocity = NULL; /* BACK, as cavity_flow_3d.c:51-77 */
    PetscCall(NSSetBoundaryCondition(ns, b, bc));
  }
  PetscCall(NSSetFromOptions(ns));
  PetscCall(NSSetUp(ns));
  PetscCall(NSGetSolution(ns, &sol));
  PetscCall(VecSet(sol, 0.));
  if (reading) {
    PetscCall(PetscViewerFlucaCGNSOpen(PETSC_COMM_WORLD, argv[2], FILE_MODE_READ, &viewer));
    PetscCall(NSLoadSolution(ns, viewer));
    PetscCall(PetscViewerDestroy(&viewer));
  }
  PetscCall(NSSolve(ns));
  PetscCall(NSGetTimeStep(ns, &step));
  PetscCall(NSGetTime(ns, &t));
  PetscCall(PetscPrintf(PETSC_COMM_WORLD, "finished at step %d time %g\n", (int)step, (double)t));
  PetscCall(PetscSNPrintf(out, sizeof(out), "%s%s", argv[2], reading ? ".out" : ""));
  PetscCall(MeshSetOutputSequenceNumber(mesh, step, t)); /* what NSMonitor does before its monitors view anything (nsmon.c:42) */
  PetscCall(PetscViewerFlucaCGNSOpen(PETSC_COMM_WORLD, out, FILE_MODE_WRITE, &viewer));
  PetscCall(NSViewSolution(ns, viewer));
  PetscCall(PetscViewerDestroy(&viewer));
  PetscCall(MeshDestroy(&mesh));
  PetscCall(NSDestroy(&ns));
  PetscCall(FlucaFinalize());
  return 0;
}
