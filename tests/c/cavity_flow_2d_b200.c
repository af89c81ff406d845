/*
 * tests/c/cavity_flow_2d_b200.c -- plain-C driver of the C ABI (include/fluca_b200.h), no Python and no PETSc.
 *
 * The case of the reference's fluca/tests/cavity_flow/cavity_flow_2d.c:9-21,49-68 (2-D lid-driven cavity, unit square,
 * three no-slip walls and a lid moving at (1, 0), zero initial state) run through the entry points the PETSc glue
 * (glue/nsb200.c) calls, in the order it calls them: create -> boundary planes -> set_state -> step ... -> get_state.
 * It shows that the drop-in boundary is usable from the reference's own language with nothing but the header, and
 * tests/test_c_driver.py checks its output against the oracle.
 *
 *   usage: cavity_flow_2d_b200 [n] [Re] [steps] [mode 0 coupled | 1 fractional] [schur_ainv] [upper_ainv] [rtol]
 *   prints one line per step (the analogue of -ns_monitor / -ns_ksp_monitor counts) and a final line
 *     RESULT n steps sum_u sum_v sum_p u_centre_column...
 */
#include <fluca_b200.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#define CHECK(call) \
  do { \
    int rc_ = (call); \
    if (rc_ != FLUCA_B200_OK) { \
      fprintf(stderr, "%s failed (%d): %s\n", #call, rc_, fluca_b200_last_error()); \
      return 1; \
    } \
  } while (0)

int main(int argc, char **argv)
{
  const int    n     = argc > 1 ? atoi(argv[1]) : 32;
  const double Re    = argc > 2 ? atof(argv[2]) : 100.;
  const int    steps = argc > 3 ? atoi(argv[3]) : 5;
  const int    mode  = argc > 4 ? atoi(argv[4]) : FLUCA_B200_MODE_COUPLED;
  const int    sa    = argc > 5 ? atoi(argv[5]) : FLUCA_B200_AINV_ID;
  const int    ua    = argc > 6 ? atoi(argv[6]) : FLUCA_B200_AINV_ID;
  const double rtol  = argc > 7 ? atof(argv[7]) : 0.; /* 0: the reference defaults (1e-5); the parity test passes 1e-13 */
  const double dt    = 0.5 / n; /* dt = 0.5 h */
  int          i, j, b, slot, k;

  if (n < 4 || steps < 1) {
    fprintf(stderr, "need n >= 4 and steps >= 1\n");
    return 2;
  }

  /* MeshCartCreate2d + MeshCartSetUniformCoordinates(0, 1, 0, 1) (cart.c:458-465): n + 1 face coordinates per direction */
  double *xf = malloc(sizeof(double) * (size_t)(n + 1));
  for (i = 0; i <= n; ++i) xf[i] = (double)i / n;

  fluca_b200_desc desc = {0};
  desc.dim   = 2;
  desc.n[0] = desc.n[1] = n, desc.n[2] = 1;
  desc.xf[0] = xf, desc.xf[1] = xf, desc.xf[2] = NULL;
  for (b = 0; b < 4; ++b) desc.bc_type[b] = FLUCA_B200_BC_VELOCITY; /* LEFT, RIGHT, DOWN, UP */
  desc.rho = 1., desc.mu = 1. / Re, desc.dt = dt;
  desc.k0 = 0, desc.nzl = 1;
  desc.mode = mode; /* everything else 0: the defaults of nssol.c:22-25 */
  desc.outer_rtol = desc.mom_rtol = desc.schur_rtol = rtol;

  fluca_b200_solver *s = NULL;
  CHECK(fluca_b200_create(&desc, NULL, &s));
  if (sa || ua) CHECK(fluca_b200_set_abf_ainv_types(s, sa, ua));

  /* NSSetBoundaryCondition: wall_velocity = (0, 0) on LEFT/RIGHT/DOWN, moving_wall_velocity = (1, 0) on UP; the planes hold
   * [component][boundary point], n points per boundary in 2-D, the same values at t^n and t^n + dt */
  double *plane = malloc(sizeof(double) * 2 * (size_t)n);
  for (b = 0; b < 4; ++b) {
    for (k = 0; k < n; ++k) plane[k] = (b == 3) ? 1. : 0., plane[n + k] = 0.;
    for (slot = 0; slot < 2; ++slot) CHECK(fluca_b200_set_boundary_velocity(s, b, slot, plane));
  }

  /* VecSet(sol, 0.) (cavity_flow_2d.c:74-75) */
  const size_t ncell = (size_t)n * n, nfx = (size_t)(n + 1) * n, nfy = (size_t)n * (n + 1);
  double      *v = calloc(2 * ncell, sizeof(double)), *p = calloc(ncell, sizeof(double)), *ph = calloc(ncell, sizeof(double));
  double      *U[3] = {calloc(nfx, sizeof(double)), calloc(nfy, sizeof(double)), NULL};
  CHECK(fluca_b200_set_state(s, v, (const double *const *)U, p, ph));

  /* NSSolve: NSStep until max_steps (nsbasic.c:325-351) */
  fluca_b200_stats st;
  for (k = 0; k < steps; ++k) {
    CHECK(fluca_b200_step(s, k * dt, k, &st));
    printf("%d NS dt %g time %g  outer %d momentum %d schur %d  |r| %.3e -> %.3e\n", k + 1, dt, (k + 1) * dt, st.outer_its, st.mom_its, st.schur_its, st.outer_rnorm0, st.outer_rnorm);
  }
  CHECK(fluca_b200_get_state(s, v, U, p, ph));

  double su = 0., sv = 0., sp = 0., div = 0.;
  for (k = 0; k < (int)ncell; ++k) su += v[k], sv += v[ncell + k], sp += p[k];
  for (j = 0; j < n; ++j)
    for (i = 0; i < n; ++i) {
      const double d = (U[0][(size_t)j * (n + 1) + i + 1] - U[0][(size_t)j * (n + 1) + i]) * n + (U[1][(size_t)(j + 1) * n + i] - U[1][(size_t)j * n + i]) * n;
      if (fabs(d) > div) div = fabs(d);
    }
  printf("RESULT %d %d %.15e %.15e %.15e", n, steps, su, sv, sp);
  for (j = 0; j < n; ++j) printf(" %.15e", v[(size_t)j * n + n / 2]); /* u along the vertical line through cell column n/2 */
  printf("\nMAXDIV %.3e\n", div);

  CHECK(fluca_b200_destroy(s));
  free(xf), free(plane), free(v), free(p), free(ph), free(U[0]), free(U[1]);
  return 0;
}
