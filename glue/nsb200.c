/*
 * glue/nsb200.c -- NS type "b200": the reference-side binding of libfluca_b200.so.
 *
 * This translation unit belongs in the reference tree (fluca/src/ns/impl/b200/nsb200.c); it is plain
 * C on PETSc like every other file there and holds NO arithmetic: fields live on the GPU, all operators are
 * CUDA kernels behind include/fluca_b200.h.  It fills the nine NSOps of the reference's type interface
 * (fluca/include/fluca/private/nsimpl.h:21-31; model: NSCreate_CNLinear, fluca/src/ns/impl/linearcn/cnlinear.c:164-187)
 * and is registered with NSRegister (fluca/include/flucans.h:92), so apps select it with -ns_type b200.
 *
 * PETSc, MPI, HDF5 and CGNS are not installed in the build image of this repository, so this file
 * cannot be compiled there; tests/test_glue_syntax.py compiles it against a declaration-only stub of the PETSc /
 * Fluca API subset it uses (tests/petsc_stub/).  INTEGRATION.md shows the build line for a real PETSc tree.
 *
 * What stays on the host, as in the reference: options (nsopts.c:179-194), the time loop and failure policy
 * (nsbasic.c:276-351), monitors, CGNS output of ns->sol (nssol.c:130-204), evaluation of the user's boundary
 * callbacks (flucansbc.h:14).  Decomposition: 1 MPI rank <-> 1 GPU, z-slabs (-cart_ranks_x 1 -cart_ranks_y 1
 * -cart_ranks_z P), NCCL bootstrapped by broadcasting the unique id over the NS communicator.
 */
#include <fluca/private/nsimpl.h>
#include <flucameshcart.h>
#include <flucaviewer.h>
#include <petscdmstag.h>
#include <fluca_b200.h>
#include "flucansb200.h"

typedef struct {
  fluca_b200_solver *solver;
  Vec                phalf;        /* "PressureHalfStep", restart-compatible with cnlinear (cnlinear.c:54,146-162) */
  /* options */
  PetscInt  mode;                  /* 0 coupled (reference default semantics), 1 fractional (one PCApply_ABF) */
  PetscInt  restart, outer_maxit, inner_maxit;
  PetscReal outer_rtol, mom_rtol, schur_rtol;
  PetscInt  sync_interval;         /* download ns->sol every k steps (1 = every step; 0 = only on demand) */
  PCABFAinvType schur_ainv, upper_ainv; /* the PCABF variants (flucans.h:99-107): ID, DIAG or ROWSUM approximation of A^-1 */
  /* geometry of this rank's slab */
  PetscInt dim, M, N, P, k0, nzl;
  PetscBool lastz, per[3];
  PetscBool wallz[2];             /* this rank holds the BACK / FRONT boundary plane */
  /* host staging in the C-ABI layout: page-locked (fluca_b200_host_alloc), so uploads and downloads run at full DMA rate */
  double *hv, *hU[3], *hp, *hph, *hbc;
  size_t  ncell, nface[3];
  /* boundary planes: the last uploaded copy per (boundary, slot); a plane whose callback returns the same values is not sent
     again, and with -ns_b200_bc_time_independent the callbacks are evaluated once */
  double   *bccache[6][2];
  PetscBool bcvalid[6][2], bc_time_independent;
  PetscBool ksp_monitor;           /* -ns_ksp_monitor: print the outer residual history of every step in PETSc's format */
  PetscBool inner_monitor[2];      /* -ns_abf_momentum_ksp_monitor / -ns_abf_schur_ksp_monitor: every inner residual norm, as it is computed */
  MPI_Comm  comm;                  /* for the monitor callback */
  PetscBool no_bcg_quirk, no_t_outlet_quirk; /* the two places where the reference's 3-D file differs from its 2-D file (include/fluca_b200.h):
                                                default PETSC_FALSE = do what cnlinearcart3d.c does */
  /* coherence between ns->sol (host) and the device state */
  PetscObjectState solstate;
  PetscBool        device_current, host_current;
  PetscBool        staged;        /* fluca_b200_stage_state is in flight or complete and not yet unpacked into ns->sol */
  fluca_b200_stats stats;
} NS_B200;

/* PETSc log events of the type (-log_view): the base class brackets the whole step with its own NSStep event (nsbasic.c:284-286);
 * these split what this type adds around the device work -- the DMStag <-> compact-layout conversions with their host <-> device
 * copies, and the evaluation + upload of the boundary callbacks. */
static PetscLogEvent NSB200_HostToDevice = 0, NSB200_DeviceToHost = 0, NSB200_BoundaryData = 0, NSB200_DeviceStep = 0;
static PetscErrorCode B200RegisterEvents_Private(void)
{
  PetscFunctionBegin;
  if (!NSB200_DeviceStep) {
    PetscCall(PetscLogEventRegister("NSB200HostToDevice", NS_CLASSID, &NSB200_HostToDevice));
    PetscCall(PetscLogEventRegister("NSB200DeviceToHost", NS_CLASSID, &NSB200_DeviceToHost));
    PetscCall(PetscLogEventRegister("NSB200BoundaryData", NS_CLASSID, &NSB200_BoundaryData));
    PetscCall(PetscLogEventRegister("NSB200DeviceStep", NS_CLASSID, &NSB200_DeviceStep));
  }
  PetscFunctionReturn(PETSC_SUCCESS);
}

#define B200Call(ns, call) \
  do { \
    int rc_ = (call); \
    PetscCheck(rc_ == FLUCA_B200_OK, PetscObjectComm((PetscObject)(ns)), PETSC_ERR_LIB, "fluca_b200: %s", fluca_b200_last_error()); \
  } while (0)

/* ------------------------------------------------------------------ DMStag <-> C-ABI layout
 * Per-element entry order and the extra faces of the last rank are DMStag's business: everything goes through
 * DMStagVecGetArray + DMStagGetLocationSlot (SURVEY.md 8b "do not assume raw DMStag global ordering"). */
typedef struct {
  DMStagStencilLocation loc;        /* where the entry lives in its element */
  PetscInt              c;          /* component at that location */
  PetscInt              ex, ey, ez; /* one more entry in x / y / z than there are cells (the face layer of the last rank) */
  double               *host;       /* compact array of the C ABI (x fastest) */
} B200Slot;

/* One DMGlobalToLocal and one array access per DM and Vec, however many slots are read from it (velocity: dim components, face-normal
 * velocity: dim face locations): at 512^3 each pass over a local vector is gigabytes of host traffic. */
static PetscErrorCode B200Upload_Private(DM dm, Vec g, PetscInt nslots, const B200Slot sl[])
{
  PetscInt  x, y, z, m, n, p, dim, slot, i, j, k, q, slots[3];
  PetscBool same = PETSC_TRUE; /* all slots cover the same index range */
  Vec       l;

  PetscFunctionBegin;
  PetscCall(DMGetDimension(dm, &dim));
  PetscCall(DMStagGetCorners(dm, &x, &y, &z, &m, &n, &p, NULL, NULL, NULL));
  PetscCheck(nslots >= 1 && nslots <= 3, PetscObjectComm((PetscObject)dm), PETSC_ERR_ARG_WRONG, "1 to 3 slots");
  for (q = 0; q < nslots; ++q) {
    PetscCall(DMStagGetLocationSlot(dm, sl[q].loc, sl[q].c, &slots[q]));
    if (sl[q].ex != sl[0].ex || sl[q].ey != sl[0].ey || sl[q].ez != sl[0].ez) same = PETSC_FALSE;
  }
  PetscCall(DMGetLocalVector(dm, &l));
  PetscCall(DMGlobalToLocal(dm, g, INSERT_VALUES, l));
  if (dim == 2) {
    const PetscScalar ***a;
    PetscCall(DMStagVecGetArrayRead(dm, l, &a));
    for (q = 0; q < nslots; ++q) {
      const PetscInt mm = m + sl[q].ex, nn = n + sl[q].ey;
      double        *host = sl[q].host;
      slot = slots[q];
      for (j = 0; j < nn; ++j)
        for (i = 0; i < mm; ++i) host[i + (size_t)mm * j] = PetscRealPart(a[y + j][x + i][slot]);
    }
    PetscCall(DMStagVecRestoreArrayRead(dm, l, &a));
  } else {
    const PetscScalar ****a;
    PetscCall(DMStagVecGetArrayRead(dm, l, &a));
    if (same) { /* the components of one element are neighbours in memory: one pass over the local array for all of them */
      for (k = 0; k < p + sl[0].ez; ++k)
        for (j = 0; j < n + sl[0].ey; ++j)
          for (i = 0; i < m + sl[0].ex; ++i) {
            const PetscScalar *e = a[z + k][y + j][x + i];
            const size_t       o = i + (size_t)(m + sl[0].ex) * (j + (size_t)(n + sl[0].ey) * k);
            for (q = 0; q < nslots; ++q) sl[q].host[o] = PetscRealPart(e[slots[q]]);
          }
    } else
      for (q = 0; q < nslots; ++q) {
        const PetscInt mm = m + sl[q].ex, nn = n + sl[q].ey, pp = p + sl[q].ez;
        double        *host = sl[q].host;
        slot = slots[q];
        for (k = 0; k < pp; ++k)
          for (j = 0; j < nn; ++j)
            for (i = 0; i < mm; ++i) host[i + (size_t)mm * (j + (size_t)nn * k)] = PetscRealPart(a[z + k][y + j][x + i][slot]);
      }
    PetscCall(DMStagVecRestoreArrayRead(dm, l, &a));
  }
  PetscCall(DMRestoreLocalVector(dm, &l));
  PetscFunctionReturn(PETSC_SUCCESS);
}

/* The reverse: all slots of one Vec are written into ONE local vector, and one DMLocalToGlobal with INSERT_VALUES moves the owned
 * entries (every entry of these DMs is one of the slots, and is owned by exactly one rank: nothing of g keeps an old value). */
static PetscErrorCode B200Download_Private(DM dm, Vec g, PetscInt nslots, const B200Slot sl[])
{
  PetscInt  x, y, z, m, n, p, dim, slot, i, j, k, q, slots[3];
  PetscBool same = PETSC_TRUE; /* all slots cover the same index range */
  Vec       l;

  PetscFunctionBegin;
  PetscCall(DMGetDimension(dm, &dim));
  PetscCall(DMStagGetCorners(dm, &x, &y, &z, &m, &n, &p, NULL, NULL, NULL));
  PetscCheck(nslots >= 1 && nslots <= 3, PetscObjectComm((PetscObject)dm), PETSC_ERR_ARG_WRONG, "1 to 3 slots");
  for (q = 0; q < nslots; ++q) {
    PetscCall(DMStagGetLocationSlot(dm, sl[q].loc, sl[q].c, &slots[q]));
    if (sl[q].ex != sl[0].ex || sl[q].ey != sl[0].ey || sl[q].ez != sl[0].ez) same = PETSC_FALSE;
  }
  PetscCall(DMGetLocalVector(dm, &l));
  PetscCall(VecZeroEntries(l)); /* ghost entries and the entries a partial element does not have */
  if (dim == 2) {
    PetscScalar ***a;
    PetscCall(DMStagVecGetArray(dm, l, &a));
    for (q = 0; q < nslots; ++q) {
      const PetscInt mm = m + sl[q].ex, nn = n + sl[q].ey;
      const double  *host = sl[q].host;
      slot = slots[q];
      for (j = 0; j < nn; ++j)
        for (i = 0; i < mm; ++i) a[y + j][x + i][slot] = host[i + (size_t)mm * j];
    }
    PetscCall(DMStagVecRestoreArray(dm, l, &a));
  } else {
    PetscScalar ****a;
    PetscCall(DMStagVecGetArray(dm, l, &a));
    if (same) {
      for (k = 0; k < p + sl[0].ez; ++k)
        for (j = 0; j < n + sl[0].ey; ++j)
          for (i = 0; i < m + sl[0].ex; ++i) {
            PetscScalar *e = a[z + k][y + j][x + i];
            const size_t o = i + (size_t)(m + sl[0].ex) * (j + (size_t)(n + sl[0].ey) * k);
            for (q = 0; q < nslots; ++q) e[slots[q]] = sl[q].host[o];
          }
    } else
      for (q = 0; q < nslots; ++q) {
        const PetscInt mm = m + sl[q].ex, nn = n + sl[q].ey, pp = p + sl[q].ez;
        const double  *host = sl[q].host;
        slot = slots[q];
        for (k = 0; k < pp; ++k)
          for (j = 0; j < nn; ++j)
            for (i = 0; i < mm; ++i) a[z + k][y + j][x + i][slot] = host[i + (size_t)mm * (j + (size_t)nn * k)];
      }
    PetscCall(DMStagVecRestoreArray(dm, l, &a));
  }
  PetscCall(DMLocalToGlobal(dm, l, INSERT_VALUES, g));
  PetscCall(DMRestoreLocalVector(dm, &l));
  PetscFunctionReturn(PETSC_SUCCESS);
}

static const DMStagStencilLocation B200FaceLoc[3] = {DMSTAG_LEFT, DMSTAG_DOWN, DMSTAG_BACK};

static PetscErrorCode B200HostToDevice_Private(NS ns)
{
  NS_B200 *b = (NS_B200 *)ns->data;
  DM       sdm, vdm, Sdm;
  Vec      v, V, p;
  PetscInt d;

  PetscFunctionBegin;
  PetscCall(PetscLogEventBegin(NSB200_HostToDevice, (PetscObject)ns, 0, 0, 0));
  PetscCall(MeshGetDM(ns->mesh, MESH_DM_SCALAR, &sdm));
  PetscCall(MeshGetDM(ns->mesh, MESH_DM_VECTOR, &vdm));
  PetscCall(MeshGetDM(ns->mesh, MESH_DM_STAG_SCALAR, &Sdm));
  PetscCall(NSGetSolutionSubVector(ns, NS_FIELD_VELOCITY, &v));
  PetscCall(NSGetSolutionSubVector(ns, NS_FIELD_FACE_NORMAL_VELOCITY, &V));
  PetscCall(NSGetSolutionSubVector(ns, NS_FIELD_PRESSURE, &p));
  {
    B200Slot sv[3], sU[3], sp = {DMSTAG_ELEMENT, 0, 0, 0, 0, b->hp}, sph = {DMSTAG_ELEMENT, 0, 0, 0, 0, b->hph};
    for (d = 0; d < b->dim; ++d) {
      const B200Slot cv = {DMSTAG_ELEMENT, d, 0, 0, 0, b->hv + b->ncell * d};
      const B200Slot cU = {B200FaceLoc[d], 0, d == 0 && !b->per[0], d == 1 && !b->per[1], d == 2 && b->lastz, b->hU[d]};
      sv[d] = cv, sU[d] = cU;
    }
    PetscCall(B200Upload_Private(vdm, v, b->dim, sv));
    PetscCall(B200Upload_Private(Sdm, V, b->dim, sU));
    PetscCall(B200Upload_Private(sdm, p, 1, &sp));
    PetscCall(B200Upload_Private(sdm, b->phalf, 1, &sph));
  }
  PetscCall(NSRestoreSolutionSubVector(ns, NS_FIELD_VELOCITY, &v));
  PetscCall(NSRestoreSolutionSubVector(ns, NS_FIELD_FACE_NORMAL_VELOCITY, &V));
  PetscCall(NSRestoreSolutionSubVector(ns, NS_FIELD_PRESSURE, &p));
  {
    const double *U[3] = {b->hU[0], b->hU[1], b->hU[2]};
    B200Call(ns, fluca_b200_set_state(b->solver, b->hv, U, b->hp, b->hph));
  }
  b->device_current = PETSC_TRUE;
  PetscCall(PetscLogEventEnd(NSB200_HostToDevice, (PetscObject)ns, 0, 0, 0));
  PetscFunctionReturn(PETSC_SUCCESS);
}

static PetscErrorCode B200DeviceToHost_Private(NS ns)
{
  NS_B200 *b = (NS_B200 *)ns->data;
  DM       sdm, vdm, Sdm;
  Vec      v, V, p;
  PetscInt d;

  const double *hv, *hU[3], *hp, *hph;

  PetscFunctionBegin;
  PetscCall(PetscLogEventBegin(NSB200_DeviceToHost, (PetscObject)ns, 0, 0, 0));
  /* the library's own pinned buffers (fluca_b200_stage_state): full-rate DMA and no second host copy.  If the user staged
     the state earlier (NSB200StageSolution) the copy has been running behind the steps issued since. */
  if (!b->staged) B200Call(ns, fluca_b200_stage_state(b->solver));
  B200Call(ns, fluca_b200_staged_state(b->solver, &hv, hU, &hp, &hph));
  b->staged = PETSC_FALSE;
  PetscCall(MeshGetDM(ns->mesh, MESH_DM_SCALAR, &sdm));
  PetscCall(MeshGetDM(ns->mesh, MESH_DM_VECTOR, &vdm));
  PetscCall(MeshGetDM(ns->mesh, MESH_DM_STAG_SCALAR, &Sdm));
  PetscCall(NSGetSolutionSubVector(ns, NS_FIELD_VELOCITY, &v));
  PetscCall(NSGetSolutionSubVector(ns, NS_FIELD_FACE_NORMAL_VELOCITY, &V));
  PetscCall(NSGetSolutionSubVector(ns, NS_FIELD_PRESSURE, &p));
  {
    B200Slot sv[3], sU[3], sp = {DMSTAG_ELEMENT, 0, 0, 0, 0, (double *)hp}, sph = {DMSTAG_ELEMENT, 0, 0, 0, 0, (double *)hph}; /* read only */
    for (d = 0; d < b->dim; ++d) {
      const B200Slot cv = {DMSTAG_ELEMENT, d, 0, 0, 0, (double *)hv + b->ncell * d};
      const B200Slot cU = {B200FaceLoc[d], 0, d == 0 && !b->per[0], d == 1 && !b->per[1], d == 2 && b->lastz, (double *)hU[d]};
      sv[d] = cv, sU[d] = cU;
    }
    PetscCall(B200Download_Private(vdm, v, b->dim, sv));
    PetscCall(B200Download_Private(Sdm, V, b->dim, sU));
    PetscCall(B200Download_Private(sdm, p, 1, &sp));
    PetscCall(B200Download_Private(sdm, b->phalf, 1, &sph));
  }
  PetscCall(NSRestoreSolutionSubVector(ns, NS_FIELD_VELOCITY, &v));
  PetscCall(NSRestoreSolutionSubVector(ns, NS_FIELD_FACE_NORMAL_VELOCITY, &V));
  PetscCall(NSRestoreSolutionSubVector(ns, NS_FIELD_PRESSURE, &p));
  PetscCall(PetscObjectStateGet((PetscObject)ns->sol, &b->solstate));
  b->host_current = PETSC_TRUE;
  PetscCall(PetscLogEventEnd(NSB200_DeviceToHost, (PetscObject)ns, 0, 0, 0));
  PetscFunctionReturn(PETSC_SUCCESS);
}

/* ------------------------------------------------------------------ boundary callbacks -> boundary planes
 * The user's NSBoundaryConditionFunction (flucansbc.h:14) is an arbitrary host function of (t, x): evaluate it at the
 * boundary-face centres of this rank's slab, at the times the step needs (cnlinearcart3d.c:2967-3033), and upload. */
static PetscErrorCode B200UploadBoundaryData_Private(NS ns)
{
  NS_B200            *b = (NS_B200 *)ns->data;
  const PetscScalar **ax, **ay, **az = NULL;
  PetscInt            iprev, ielem, x, y, z, m, n, p, bnd, slot, i, j, c;
  DM                  sdm;
  const PetscReal     tq = ns->step == 0 ? ns->t : ns->t - 0.5 * ns->dt;

  PetscFunctionBegin;
  PetscCall(PetscLogEventBegin(NSB200_BoundaryData, (PetscObject)ns, 0, 0, 0));
  PetscCall(MeshGetDM(ns->mesh, MESH_DM_SCALAR, &sdm));
  PetscCall(DMStagGetCorners(sdm, &x, &y, &z, &m, &n, &p, NULL, NULL, NULL));
  PetscCall(DMStagGetProductCoordinateArraysRead(sdm, &ax, &ay, &az));
  PetscCall(DMStagGetProductCoordinateLocationSlot(sdm, DMSTAG_LEFT, &iprev));
  PetscCall(DMStagGetProductCoordinateLocationSlot(sdm, DMSTAG_ELEMENT, &ielem));
  if (b->dim == 2) p = 1, z = 0;
  for (bnd = 0; bnd < 2 * b->dim; ++bnd) {
    const NSBoundaryCondition *bc = &ns->bcs[bnd];
    const PetscInt d = bnd / 2, side = bnd % 2;
    const PetscInt n0 = d == 0 ? n : m, n1 = d == 2 ? n : p; /* fastest, slowest extent of the plane */
    const PetscInt gl[3] = {b->M, b->N, b->P};
    const size_t   np = (size_t)n0 * n1;
    if (bc->type != NS_BC_VELOCITY && bc->type != NS_BC_PRESSURE_OUTLET) continue;
    if (d == 2 && !b->wallz[side]) continue; /* only the first / last slab holds (and has coordinates for) a z boundary */
    for (slot = 0; slot < 2; ++slot) {
      const PetscReal t = bc->type == NS_BC_VELOCITY ? (slot ? ns->t + ns->dt : ns->t) : (slot ? ns->t + 0.5 * ns->dt : tq);
      const size_t    nval = np * (bc->type == NS_BC_VELOCITY ? (size_t)b->dim : 1);
      PetscBool       same = PETSC_FALSE;
      if (b->bc_time_independent && b->bcvalid[bnd][slot]) continue; /* the user promised constant data: evaluated once */
      for (j = 0; j < n1; ++j)
        for (i = 0; i < n0; ++i) {
          PetscReal   xb[3] = {0., 0., 0.};
          PetscScalar val[3] = {0., 0., 0.};
          PetscInt    ci[3];
          /* plane coordinates: x boundaries [k][j], y boundaries [k][i], z boundaries [j][i] (include/fluca_b200.h) */
          ci[0] = d == 0 ? 0 : x + i;
          ci[1] = d == 0 ? y + i : (d == 1 ? 0 : y + j);
          ci[2] = d == 2 ? 0 : z + j;
          xb[0] = d == 0 ? ax[side ? gl[0] : 0][iprev] : ax[ci[0]][ielem];
          xb[1] = d == 1 ? ay[side ? gl[1] : 0][iprev] : ay[ci[1]][ielem];
          if (b->dim == 3) xb[2] = d == 2 ? az[side ? gl[2] : 0][iprev] : az[ci[2]][ielem];
          if (bc->type == NS_BC_VELOCITY) {
            PetscCall(bc->velocity(b->dim, t, xb, val, bc->ctx_velocity));
            for (c = 0; c < b->dim; ++c) b->hbc[c * np + i + (size_t)n0 * j] = PetscRealPart(val[c]);
          } else {
            PetscCall(bc->pressure(b->dim, t, xb, val, bc->ctx_pressure));
            b->hbc[i + (size_t)n0 * j] = PetscRealPart(val[0]);
          }
        }
      /* ~1 M callback values per step at 512^3; most runs have steady boundary data: skip the upload of an unchanged plane */
      if (!b->bccache[bnd][slot]) PetscCall(PetscMalloc1(np * 3, &b->bccache[bnd][slot]));
      if (b->bcvalid[bnd][slot]) PetscCall(PetscArraycmp(b->hbc, b->bccache[bnd][slot], nval, &same));
      if (same) continue;
      if (bc->type == NS_BC_VELOCITY) B200Call(ns, fluca_b200_set_boundary_velocity(b->solver, (int)bnd, (int)slot, b->hbc));
      else B200Call(ns, fluca_b200_set_boundary_pressure(b->solver, (int)bnd, (int)slot, b->hbc));
      PetscCall(PetscArraycpy(b->bccache[bnd][slot], b->hbc, nval));
      b->bcvalid[bnd][slot] = PETSC_TRUE;
    }
  }
  PetscCall(DMStagRestoreProductCoordinateArraysRead(sdm, &ax, &ay, &az));
  PetscCall(PetscLogEventEnd(NSB200_BoundaryData, (PetscObject)ns, 0, 0, 0));
  PetscFunctionReturn(PETSC_SUCCESS);
}

/* ------------------------------------------------------------------ NSOps */
static PetscErrorCode NSSetFromOptions_B200(NS ns, PetscOptionItems PetscOptionsObject)
{
  NS_B200 *b = (NS_B200 *)ns->data;

  PetscFunctionBegin;
  PetscOptionsHeadBegin(PetscOptionsObject, "NSB200 Options"); /* pattern: cnlinear.c:5-12 */
  PetscCall(PetscOptionsInt("-ns_b200_mode", "0: coupled solve, PC = ABF (reference default); 1: one ABF application (classical fractional step)", "", b->mode, &b->mode, NULL));
  /* the reference's OWN names for the solver knobs first, so that a command line written for cnlinear keeps its meaning when only
   * -ns_type changes: the KSP of the SNES lives under "ns_" (nssol.c:13-25), the two KSPs of PCABF under "ns_abf_momentum_" and
   * "ns_abf_schur_" (abfpc.c:33-46).  PETSc hands the same values to those objects of the base class, which this type never solves
   * with.  The -ns_b200_* spellings below are read afterwards and win. */
  PetscCall(PetscOptionsReal("-ns_ksp_rtol", "outer relative tolerance", "KSPSetTolerances", b->outer_rtol, &b->outer_rtol, NULL));
  PetscCall(PetscOptionsInt("-ns_ksp_max_it", "outer iteration limit", "KSPSetTolerances", b->outer_maxit, &b->outer_maxit, NULL));
  PetscCall(PetscOptionsInt("-ns_ksp_gmres_restart", "restart of the outer GMRES", "KSPGMRESSetRestart", b->restart, &b->restart, NULL));
  PetscCall(PetscOptionsReal("-ns_abf_momentum_ksp_rtol", "momentum solve relative tolerance", "KSPSetTolerances", b->mom_rtol, &b->mom_rtol, NULL));
  PetscCall(PetscOptionsReal("-ns_abf_schur_ksp_rtol", "pressure solve relative tolerance", "KSPSetTolerances", b->schur_rtol, &b->schur_rtol, NULL));
  {
    PetscInt ma = 0, ms = 0; /* one limit for both inner solves in the library: the larger of the two asked for */
    PetscCall(PetscOptionsInt("-ns_abf_momentum_ksp_max_it", "momentum solve iteration limit", "KSPSetTolerances", ma, &ma, NULL));
    PetscCall(PetscOptionsInt("-ns_abf_schur_ksp_max_it", "pressure solve iteration limit", "KSPSetTolerances", ms, &ms, NULL));
    if (PetscMax(ma, ms) > 0) b->inner_maxit = PetscMax(ma, ms);
  }
  PetscCall(PetscOptionsInt("-ns_b200_gmres_restart", "restart of the outer GMRES ((restart + 1) x 7 fields of device memory)", "", b->restart, &b->restart, NULL));
  PetscCall(PetscOptionsReal("-ns_b200_outer_rtol", "outer relative tolerance (nssol.c:24 sets 1e-5)", "", b->outer_rtol, &b->outer_rtol, NULL));
  PetscCall(PetscOptionsReal("-ns_b200_momentum_rtol", "momentum solve relative tolerance", "", b->mom_rtol, &b->mom_rtol, NULL));
  PetscCall(PetscOptionsReal("-ns_b200_schur_rtol", "pressure solve relative tolerance", "", b->schur_rtol, &b->schur_rtol, NULL));
  PetscCall(PetscOptionsInt("-ns_b200_sync_interval", "copy the device state into ns->sol every k steps (0: only for viewers)", "", b->sync_interval, &b->sync_interval, NULL));
  /* the options of the PC "abf" the base class creates under the "ns_" prefix (abfpc.c:246-247, nssol.c:17): this type never
   * applies that PC, so it reads the same two names and hands the choice to the device-side ABF factors */
  PetscCall(PetscOptionsEnum("-ns_pc_abf_schur_ainv_type", "Type of approximation used in Schur complement", "PCABFSetSchurComplementAinvType", PCABFAinvTypes, (PetscEnum)b->schur_ainv, (PetscEnum *)&b->schur_ainv, NULL));
  PetscCall(PetscOptionsEnum("-ns_pc_abf_upper_ainv_type", "Type of approximation used in upper triangular matrix", "PCABFSetUpperTriangularAinvType", PCABFAinvTypes, (PetscEnum)b->upper_ainv, (PetscEnum *)&b->upper_ainv, NULL));
  PetscCall(PetscOptionsBool("-ns_b200_no_bcg_quirk", "3-D: scale the outlet-gradient BC vector by dt/rho as the 2-D file does (cnlinearcart3d.c:2977 uses 1)", "", b->no_bcg_quirk, &b->no_bcg_quirk, NULL));
  PetscCall(PetscOptionsBool("-ns_b200_no_t_outlet_quirk", "3-D: form operator T at an upper pressure outlet as the 2-D file does (cnlinearcart3d.c:1996,2055,2114 read the partial element's slot)", "", b->no_t_outlet_quirk, &b->no_t_outlet_quirk, NULL));
  PetscCall(PetscOptionsBool("-ns_b200_bc_time_independent", "the boundary callbacks do not depend on time: evaluate them once", "", b->bc_time_independent, &b->bc_time_independent, NULL));
  /* the reference prints residual histories through the KSP of its SNES (-ns_ksp_monitor); this type owns its
   * outer Krylov solver, reads the same option name and prints the same lines */
  PetscCall(PetscOptionsBool("-ns_ksp_monitor", "print the outer (coupled) residual history of every step", "KSPMonitorSet", b->ksp_monitor, &b->ksp_monitor, NULL));
  /* ... and the KSPs of its PCABF under -ns_abf_momentum_ksp_monitor / -ns_abf_schur_ksp_monitor (abfpc.c:33-46): same names, same
   * line format, printed as the inner solvers of the library go (fluca_b200_set_inner_monitor) */
  PetscCall(PetscOptionsBool("-ns_abf_momentum_ksp_monitor", "print the residual norms of every momentum solve", "KSPMonitorSet", b->inner_monitor[0], &b->inner_monitor[0], NULL));
  PetscCall(PetscOptionsBool("-ns_abf_schur_ksp_monitor", "print the residual norms of every pressure (Schur complement) solve", "KSPMonitorSet", b->inner_monitor[1], &b->inner_monitor[1], NULL));
  PetscOptionsHeadEnd();
  PetscFunctionReturn(PETSC_SUCCESS);
}

/* fluca_b200_inner_monitor_fn: KSPMonitorResidual's lines at the tab level of a KSP inside a PC inside the KSP of the SNES.  Every
 * rank gets the same calls with the same values; PetscPrintf prints on the first rank of the communicator. */
static void B200InnerMonitor_Private(void *ctx, int which, int it, double rnorm)
{
  NS_B200 *b = (NS_B200 *)ctx;
  if (!b->inner_monitor[which]) return;
  if (it == 0) (void)PetscPrintf(b->comm, "    Residual norms for ns_abf_%s_ solve.\n", which ? "schur" : "momentum");
  (void)PetscPrintf(b->comm, "    %3d KSP Residual norm %14.12e\n", it, rnorm);
}

/* NSSetUp builds a MatNest J and calls formjacobian(INIT) before ops->setup, then MatCreateVecs(J) (nsbasic.c:203-208).
 * A matrix-free type has nothing to assemble: identity placeholders on the diagonal give J valid row/column layouts.
 * The SNES / PCABF objects the base class builds on top of J are never used by ops->step. */
static PetscErrorCode NSFormJacobian_B200(NS ns, Vec x, Mat J, NSFormJacobianType type)
{
  PetscFunctionBegin;
  (void)x;
  if (type == NS_INIT_JACOBIAN) {
    const MeshDMType dmt[3] = {MESH_DM_VECTOR, MESH_DM_STAG_SCALAR, MESH_DM_SCALAR};
    PetscInt         f;
    for (f = 0; f < 3; ++f) {
      DM       dm;
      Mat      I;
      PetscInt entries;
      PetscCall(MeshGetDM(ns->mesh, dmt[f], &dm));
      PetscCall(DMStagGetEntries(dm, &entries));
      PetscCall(MatCreateConstantDiagonal(PetscObjectComm((PetscObject)ns), entries, entries, PETSC_DETERMINE, PETSC_DETERMINE, 1., &I));
      PetscCall(MatNestSetSubMat(J, f, f, I));
      PetscCall(MatDestroy(&I));
    }
    PetscCall(MatAssemblyBegin(J, MAT_FINAL_ASSEMBLY));
    PetscCall(MatAssemblyEnd(J, MAT_FINAL_ASSEMBLY));
  }
  PetscFunctionReturn(PETSC_SUCCESS);
}

static PetscErrorCode NSSetup_B200(NS ns)
{
  NS_B200        *b = (NS_B200 *)ns->data;
  MPI_Comm        comm;
  PetscMPIInt     rank, size;
  PetscInt        rx, ry, rz, x, y, z, m, n, p, d, nb, i, iprev;
  PetscBool       iscart, lx, ly, lz, fx, fy, fz;
  MeshCartBoundaryType bt[3] = {MESHCART_BOUNDARY_NONE, MESHCART_BOUNDARY_NONE, MESHCART_BOUNDARY_NONE};
  const PetscScalar **ax, **ay, **az = NULL;
  double         *xf[3] = {NULL, NULL, NULL};
  fluca_b200_desc desc;
  fluca_b200_comm *gcomm = NULL;

  PetscFunctionBegin;
  PetscCall(PetscObjectGetComm((PetscObject)ns, &comm));
  PetscCallMPI(MPI_Comm_rank(comm, &rank));
  PetscCallMPI(MPI_Comm_size(comm, &size));
  PetscCall(PetscObjectTypeCompare((PetscObject)ns->mesh, MESHCART, &iscart));
  PetscCheck(iscart, comm, PETSC_ERR_ARG_WRONG, "Unsupported Mesh type");
  PetscCall(MeshGetDimension(ns->mesh, &b->dim));
  PetscCall(MeshCartGetGlobalSizes(ns->mesh, &b->M, &b->N, &b->P));
  PetscCall(MeshCartGetNumRanks(ns->mesh, &rx, &ry, &rz));
  PetscCheck(rx == 1 && ry == 1, comm, PETSC_ERR_SUP, "NS type b200 partitions the mesh in z-slabs: run with -cart_ranks_x 1 -cart_ranks_y 1 -cart_ranks_z <ranks>");
  PetscCheck(b->dim == 3 || size == 1, comm, PETSC_ERR_SUP, "2-D meshes run on one rank");
  PetscCall(MeshCartGetCorners(ns->mesh, &x, &y, &z, &m, &n, &p));
  PetscCall(MeshCartGetIsLastRank(ns->mesh, &lx, &ly, &lz));
  PetscCall(MeshCartGetIsFirstRank(ns->mesh, &fx, &fy, &fz));
  PetscCall(MeshCartGetBoundaryTypes(ns->mesh, &bt[0], &bt[1], &bt[2]));
  for (d = 0; d < 3; ++d) b->per[d] = (PetscBool)(d < b->dim && bt[d] == MESHCART_BOUNDARY_PERIODIC);
  b->k0    = b->dim == 3 ? z : 0;
  b->nzl   = b->dim == 3 ? p : 1;
  b->lastz = (PetscBool)(b->dim == 3 && lz && !b->per[2]);
  b->wallz[0] = (PetscBool)(b->dim == 3 && fz && !b->per[2]);
  b->wallz[1] = b->lastz;

  /* global face coordinates per direction (slot PREV; cart.c:56-151): every rank holds x and y fully; z is gathered */
  PetscCall(MeshCartGetCoordinateArraysRead(ns->mesh, &ax, &ay, &az));
  PetscCall(MeshCartGetCoordinateLocationSlot(ns->mesh, MESHCART_PREV, &iprev));
  PetscCall(PetscMalloc1(b->M + 1, &xf[0]));
  PetscCall(PetscMalloc1(b->N + 1, &xf[1]));
  for (i = 0; i <= b->M; ++i) xf[0][i] = PetscRealPart(ax[i][iprev]);
  for (i = 0; i <= b->N; ++i) xf[1][i] = PetscRealPart(ay[i][iprev]);
  if (b->dim == 3) {
    PetscCall(PetscCalloc1(b->P + 1, &xf[2]));
    for (i = z; i < z + p + (lz ? 1 : 0); ++i) xf[2][i] = PetscRealPart(az[i][iprev]);
    if (b->per[2] && lz) xf[2][b->P] = PetscRealPart(az[b->P][iprev]);
    PetscCallMPI(MPI_Allreduce(MPI_IN_PLACE, xf[2], (PetscMPIInt)(b->P + 1), MPI_DOUBLE, MPI_SUM, comm)); /* every face is owned by one rank */
  }
  PetscCall(MeshCartRestoreCoordinateArraysRead(ns->mesh, &ax, &ay, &az));

  PetscCall(PetscMemzero(&desc, sizeof(desc)));
  desc.dim  = (int)b->dim;
  desc.n[0] = (int)b->M, desc.n[1] = (int)b->N, desc.n[2] = (int)(b->dim == 3 ? b->P : 1);
  for (d = 0; d < b->dim; ++d) desc.xf[d] = xf[d];
  PetscCall(MeshGetNumberBoundaries(ns->mesh, &nb));
  for (i = 0; i < nb; ++i) desc.bc_type[i] = (int)ns->bcs[i].type; /* same numeric values (flucansbc.h:5-11) */
  desc.rho = ns->rho, desc.mu = ns->mu, desc.dt = ns->dt;          /* baked in at setup, like the reference */
  desc.k0 = (int)b->k0, desc.nzl = (int)b->nzl;
  desc.mode          = (int)b->mode;
  desc.outer_rtol    = b->outer_rtol;
  desc.outer_restart = (int)b->restart;
  desc.outer_maxit   = (int)b->outer_maxit; /* 0: the library's defaults (100 outer, 500 inner) */
  desc.inner_maxit   = (int)b->inner_maxit;
  desc.mom_rtol      = b->mom_rtol;
  desc.schur_rtol    = b->schur_rtol;
  desc.no_bcg_quirk      = b->no_bcg_quirk ? 1 : 0;
  desc.no_t_outlet_quirk = b->no_t_outlet_quirk ? 1 : 0;

  if (size > 1) { /* NCCL over the GPUs of the box: rank 0 makes the id, MPI ships it */
    char id[256];
    int  nbytes = 0;
    if (rank == 0) B200Call(ns, fluca_b200_comm_unique_id(id, (int)sizeof(id), &nbytes));
    PetscCallMPI(MPI_Bcast(&nbytes, 1, MPI_INT, 0, comm));
    PetscCallMPI(MPI_Bcast(id, nbytes, MPI_BYTE, 0, comm));
    B200Call(ns, fluca_b200_comm_create_nccl(id, nbytes, (int)rank, (int)size, &gcomm));
  }
  B200Call(ns, fluca_b200_create(&desc, gcomm, &b->solver));
  if (b->schur_ainv != PC_ABF_AINV_ID || b->upper_ainv != PC_ABF_AINV_ID) B200Call(ns, fluca_b200_set_abf_ainv_types(b->solver, (int)b->schur_ainv, (int)b->upper_ainv)); /* same numeric values */
  b->comm = comm;
  if (b->inner_monitor[0] || b->inner_monitor[1]) B200Call(ns, fluca_b200_set_inner_monitor(b->solver, B200InnerMonitor_Private, b));
  for (d = 0; d < 3; ++d) PetscCall(PetscFree(xf[d]));

  b->ncell    = (size_t)m * n * b->nzl;
  b->nface[0] = (size_t)(m + (b->per[0] ? 0 : 1)) * n * b->nzl;
  b->nface[1] = (size_t)m * (n + (b->per[1] ? 0 : 1)) * b->nzl;
  b->nface[2] = b->dim == 3 ? (size_t)m * n * (b->nzl + (b->lastz ? 1 : 0)) : 0;
  B200Call(ns, fluca_b200_host_alloc(sizeof(double) * b->ncell * b->dim, (void **)&b->hv));
  for (d = 0; d < b->dim; ++d) B200Call(ns, fluca_b200_host_alloc(sizeof(double) * b->nface[d], (void **)&b->hU[d]));
  B200Call(ns, fluca_b200_host_alloc(sizeof(double) * b->ncell, (void **)&b->hp));
  B200Call(ns, fluca_b200_host_alloc(sizeof(double) * b->ncell, (void **)&b->hph));
  {
    size_t big = (size_t)PetscMax(PetscMax(m * n, m * b->nzl), n * b->nzl);
    B200Call(ns, fluca_b200_host_alloc(sizeof(double) * big * 3, (void **)&b->hbc));
  }
  PetscCall(MeshCreateGlobalVector(ns->mesh, MESH_DM_SCALAR, &b->phalf));
  PetscCall(PetscObjectSetName((PetscObject)b->phalf, "PressureHalfStep"));
  b->device_current = PETSC_FALSE;
  b->host_current   = PETSC_TRUE;
  b->solstate       = -1;
  PetscFunctionReturn(PETSC_SUCCESS);
}

static PetscErrorCode NSStep_B200(NS ns)
{
  NS_B200         *b = (NS_B200 *)ns->data;
  PetscObjectState st;
  int              rc;

  PetscFunctionBegin;
  /* initial conditions, restarts and user edits arrive through the host Vec (cavity_flow_2d.c:74-75,
     taylor_green_vortex.c:113-178, nssol.c:191-196): upload when ns->sol changed since the last download */
  PetscCall(PetscObjectStateGet((PetscObject)ns->sol, &st));
  if (!b->device_current || st != b->solstate) PetscCall(B200HostToDevice_Private(ns));
  PetscCall(B200UploadBoundaryData_Private(ns));

  PetscCall(PetscLogEventBegin(NSB200_DeviceStep, (PetscObject)ns, 0, 0, 0));
  rc = fluca_b200_step(b->solver, (double)ns->t, (int)ns->step, &b->stats);
  PetscCall(PetscLogEventEnd(NSB200_DeviceStep, (PetscObject)ns, 0, 0, 0));
  if (rc == FLUCA_B200_ERR_DIVERGED) {
    ns->reason = NS_DIVERGED_NONLINEAR_SOLVE; /* NSCheckDiverged (nsbasic.c:425-436); NSStep applies the failure policy (:293-297) */
    PetscFunctionReturn(PETSC_SUCCESS);
  }
  PetscCheck(rc == FLUCA_B200_OK, PetscObjectComm((PetscObject)ns), PETSC_ERR_LIB, "fluca_b200: %s", fluca_b200_last_error());
  PetscCall(PetscInfo(ns, "b200 step %" PetscInt_FMT ": outer its %d, momentum its %d, Schur its %d, ABF applications %d, |r|/|r0| %g\n", ns->step, b->stats.outer_its, b->stats.mom_its, b->stats.schur_its, b->stats.abf_applies, b->stats.outer_rnorm0 > 0 ? b->stats.outer_rnorm / b->stats.outer_rnorm0 : 0.));
  if (b->stats.inner_unconverged) PetscCall(PetscInfo(ns, "b200 step %" PetscInt_FMT ": %d inner solve(s) ran into their iteration limit (last |r|/|b|: momentum %g, Schur %g)\n", ns->step, b->stats.inner_unconverged, b->stats.mom_last_rel, b->stats.schur_last_rel));

  if (b->ksp_monitor) {
    /* PETSc's KSPMonitorResidual format, one block per step; the inner solves report counts and final relative residuals
       (their iterates come from a different preconditioner than the reference's ILU(0): histories are not comparable) */
    int it;
    PetscCall(PetscPrintf(PetscObjectComm((PetscObject)ns), "  Residual norms for ns_ solve.\n"));
    for (it = 0; it < b->stats.nhist; ++it) PetscCall(PetscPrintf(PetscObjectComm((PetscObject)ns), "  %3d KSP Residual norm %14.12e\n", it, b->stats.hist[it]));
    PetscCall(PetscPrintf(PetscObjectComm((PetscObject)ns), "    ns_abf_momentum_ solves: %d iterations, last |r|/|b| %g; ns_abf_schur_ solves: %d iterations, last |r|/|b| %g\n", b->stats.mom_its, b->stats.mom_last_rel, b->stats.schur_its, b->stats.schur_last_rel));
  }
  b->host_current = PETSC_FALSE;
  if (b->sync_interval > 0 && (ns->step + 1) % b->sync_interval == 0) PetscCall(B200DeviceToHost_Private(ns));
  else PetscCall(PetscObjectStateGet((PetscObject)ns->sol, &b->solstate)); /* the base class copied sol -> sol0 only */
  PetscFunctionReturn(PETSC_SUCCESS);
}

/* right-hand side b = (r_mom, r_int, r_con) of the current step, for callers of NSFormFunction.  As in the reference, whose
 * ops->formfunction fills f with b itself (cnlinearcart3d.c:2945-3043: the "b(x)" of SNESSetPicard, nsbasic.c:249; the residual
 * A x - b is formed by SNES, not by the type) */
static PetscErrorCode NSFormFunction_B200(NS ns, Vec x, Vec f)
{
  NS_B200 *b = (NS_B200 *)ns->data;
  DM       sdm, vdm, Sdm;
  IS       vis, Vis, pis;
  Vec      fv, fV, fp;
  PetscInt d;

  PetscFunctionBegin;
  (void)x; /* the system is linear: b does not depend on x */
  {
    /* same coherence rule as NSStep_B200: a user edit of ns->sol since the last download goes up first */
    PetscObjectState st;
    PetscCall(PetscObjectStateGet((PetscObject)ns->sol, &st));
    if (!b->device_current || st != b->solstate) {
      PetscCall(B200HostToDevice_Private(ns));
      PetscCall(PetscObjectStateGet((PetscObject)ns->sol, &b->solstate)); /* after the upload: restoring the sub-vectors of a VecNest bumps its state */
    }
  }
  PetscCall(B200UploadBoundaryData_Private(ns));
  B200Call(ns, fluca_b200_prepare_step(b->solver, (double)ns->t, (int)ns->step));
  {
    double *U[3] = {b->hU[0], b->hU[1], b->hU[2]};
    B200Call(ns, fluca_b200_get_rhs(b->solver, b->hv, U, b->hp));
  }
  PetscCall(MeshGetDM(ns->mesh, MESH_DM_SCALAR, &sdm));
  PetscCall(MeshGetDM(ns->mesh, MESH_DM_VECTOR, &vdm));
  PetscCall(MeshGetDM(ns->mesh, MESH_DM_STAG_SCALAR, &Sdm));
  PetscCall(NSGetField(ns, NS_FIELD_VELOCITY, NULL, NULL, &vis));
  PetscCall(NSGetField(ns, NS_FIELD_FACE_NORMAL_VELOCITY, NULL, NULL, &Vis));
  PetscCall(NSGetField(ns, NS_FIELD_PRESSURE, NULL, NULL, &pis));
  PetscCall(VecGetSubVector(f, vis, &fv));
  PetscCall(VecGetSubVector(f, Vis, &fV));
  PetscCall(VecGetSubVector(f, pis, &fp));
  {
    B200Slot sv[3], sU[3], sp = {DMSTAG_ELEMENT, 0, 0, 0, 0, b->hp};
    for (d = 0; d < b->dim; ++d) {
      const B200Slot cv = {DMSTAG_ELEMENT, d, 0, 0, 0, b->hv + b->ncell * d};
      const B200Slot cU = {B200FaceLoc[d], 0, d == 0 && !b->per[0], d == 1 && !b->per[1], d == 2 && b->lastz, b->hU[d]};
      sv[d] = cv, sU[d] = cU;
    }
    PetscCall(B200Download_Private(vdm, fv, b->dim, sv));
    PetscCall(B200Download_Private(Sdm, fV, b->dim, sU));
    PetscCall(B200Download_Private(sdm, fp, 1, &sp));
  }
  PetscCall(VecRestoreSubVector(f, vis, &fv));
  PetscCall(VecRestoreSubVector(f, Vis, &fV));
  PetscCall(VecRestoreSubVector(f, pis, &fp));
  /* fluca_b200_prepare_step has no side effect on the device state (the time-n fields alias the live state, as the
     reference's VecCopy(sol, sol0) is idempotent), so nothing is invalidated here */
  PetscFunctionReturn(PETSC_SUCCESS);
}

static PetscErrorCode NSDestroy_B200(NS ns)
{
  NS_B200 *b = (NS_B200 *)ns->data;
  PetscInt d;

  PetscFunctionBegin;
  if (b->solver) B200Call(ns, fluca_b200_destroy(b->solver));
  PetscCall(VecDestroy(&b->phalf));
  B200Call(ns, fluca_b200_host_free(b->hv));
  for (d = 0; d < 3; ++d) B200Call(ns, fluca_b200_host_free(b->hU[d]));
  B200Call(ns, fluca_b200_host_free(b->hp));
  B200Call(ns, fluca_b200_host_free(b->hph));
  B200Call(ns, fluca_b200_host_free(b->hbc));
  for (d = 0; d < 6; ++d) {
    PetscCall(PetscFree(b->bccache[d][0]));
    PetscCall(PetscFree(b->bccache[d][1]));
  }
  PetscCall(PetscFree(ns->data));
  PetscFunctionReturn(PETSC_SUCCESS);
}

static PetscErrorCode NSView_B200(NS ns, PetscViewer viewer)
{
  NS_B200  *b = (NS_B200 *)ns->data;
  PetscBool isascii;

  PetscFunctionBegin;
  PetscCall(PetscObjectTypeCompare((PetscObject)viewer, PETSCVIEWERASCII, &isascii));
  if (isascii) {
    PetscCall(PetscViewerASCIIPrintf(viewer, "  b200: mode %s, slab planes [%" PetscInt_FMT ", %" PetscInt_FMT "), outer restart %" PetscInt_FMT "\n", b->mode ? "fractional" : "coupled", b->k0, b->k0 + b->nzl, b->restart));
    PetscCall(PetscViewerASCIIPrintf(viewer, "  ABF factors: Schur complement A inverse type %s, upper triangular A inverse type %s\n", PCABFAinvTypes[b->schur_ainv], PCABFAinvTypes[b->upper_ainv])); /* as PCView_ABF, abfpc.c:265-266 */
    PetscCall(PetscViewerASCIIPrintf(viewer, "  last step: outer %d, momentum %d, Schur %d iterations\n", b->stats.outer_its, b->stats.mom_its, b->stats.schur_its));
  }
  PetscFunctionReturn(PETSC_SUCCESS);
}

/* the base class views every field of ns->sol before calling this (nssol.c:130-175): refresh the host copy first.
 * NSViewSolution has no pre-hook, so a run that writes output uses -ns_b200_sync_interval equal to its output interval. */
static PetscErrorCode NSViewSolution_B200(NS ns, PetscViewer viewer)
{
  NS_B200 *b = (NS_B200 *)ns->data;

  PetscFunctionBegin;
  if (!b->host_current) PetscCall(B200DeviceToHost_Private(ns));
  PetscCall(VecView(b->phalf, viewer)); /* cnlinear.c:146-152 */
  PetscFunctionReturn(PETSC_SUCCESS);
}

static PetscErrorCode NSLoadSolution_B200(NS ns, PetscViewer viewer)
{
  NS_B200 *b = (NS_B200 *)ns->data;

  PetscFunctionBegin;
  PetscCall(FlucaVecLoad(b->phalf, viewer)); /* cnlinear.c:154-162 */
  b->device_current = PETSC_FALSE;
  PetscFunctionReturn(PETSC_SUCCESS);
}

PetscErrorCode NSCreate_B200(NS ns)
{
  NS_B200 *b;

  PetscFunctionBegin;
  PetscCall(B200RegisterEvents_Private());
  PetscCall(PetscNew(&b));
  ns->data = (void *)b;
  b->mode          = FLUCA_B200_MODE_COUPLED;
  b->restart       = 30;   /* PETSc's GMRES default */
  b->outer_rtol    = 1e-5; /* nssol.c:22-25 */
  b->mom_rtol      = 1e-5;
  b->schur_rtol    = 1e-5;
  b->sync_interval = 1;
  b->schur_ainv    = PC_ABF_AINV_ID; /* abfpc.c:328-329 */
  b->upper_ainv    = PC_ABF_AINV_ID;

#ifdef FLUCA_NS_HAS_MATRIXFREE
  /* with glue/patches/0001-ns-matrix-free-type-hooks.patch applied to the reference: NSSetUp skips J / null space / SNES /
     PCABF and NSStep skips the 7.5 GB host copy sol -> sol0 per step.  formjacobian(INIT) is then never
     called; without the patch the identity placeholders of NSFormJacobian_B200 keep the unmodified base class working. */
  ns->matrixfree = PETSC_TRUE;
#endif
  ns->ops->setfromoptions = NSSetFromOptions_B200;
  ns->ops->setup          = NSSetup_B200;
  ns->ops->step           = NSStep_B200;
  ns->ops->formjacobian   = NSFormJacobian_B200;
  ns->ops->formfunction   = NSFormFunction_B200;
  ns->ops->destroy        = NSDestroy_B200;
  ns->ops->view           = NSView_B200;
  ns->ops->viewsolution   = NSViewSolution_B200;
  ns->ops->loadsolution   = NSLoadSolution_B200;
  PetscFunctionReturn(PETSC_SUCCESS);
}

/* One line in NSRegisterAll (nsreg.c:13-20) ...                         NSRegister(NSB200, NSCreate_B200);
 * ... or no change to the reference at all: ship this file as a shared library and run with
 * -dll_append libfluca_nsb200.so ; PETSc calls this hook when it loads the library. */
PETSC_EXTERN PetscErrorCode PetscDLLibraryRegister_fluca_nsb200(void)
{
  PetscFunctionBegin;
  PetscCall(NSRegister(NSB200, NSCreate_B200));
  PetscFunctionReturn(PETSC_SUCCESS);
}

/* immersed-boundary markers of the b200 type (the reference only plans IBM: README.md:14, THEORY_GUIDE.md:130-132) */
PetscErrorCode NSB200SetMarkers(NS ns, PetscInt n, const PetscReal X[], const PetscReal Ud[], const PetscReal dV[], PetscInt delta_points)
{
  NS_B200  *b;
  PetscBool match;

  PetscFunctionBegin;
  PetscCall(PetscObjectTypeCompare((PetscObject)ns, NSB200, &match));
  PetscCheck(match, PetscObjectComm((PetscObject)ns), PETSC_ERR_ARG_WRONG, "NS type is not b200");
  PetscCheck(ns->setupcalled, PetscObjectComm((PetscObject)ns), PETSC_ERR_ARG_WRONGSTATE, "This function must be called after NSSetUp()");
  b = (NS_B200 *)ns->data;
  B200Call(ns, fluca_b200_set_markers(b->solver, (long)n, X, Ud, dV, (int)delta_points));
  PetscFunctionReturn(PETSC_SUCCESS);
}

/* Asynchronous output for runs with -ns_b200_sync_interval 0.  NSViewSolution views ns->sol BEFORE it calls the
 * type's hook (nssol.c:143-149), so a device-authoritative type cannot refresh lazily from inside it; an application that
 * knows its output cadence instead brackets the steps it wants to overlap:
 *     NSStep(ns); NSB200StageSolution(ns);     -- state n starts flowing to pinned host memory on its own stream
 *     NSStep(ns); ...                           -- the time loop continues on the device
 *     NSB200SyncSolution(ns); NSViewSolution(ns, viewer);   -- ns->sol (+ p-half) = state n, then written as usual */
PetscErrorCode NSB200StageSolution(NS ns)
{
  NS_B200 *b = (NS_B200 *)ns->data;

  PetscFunctionBegin;
  PetscCheck(ns->setupcalled && b->solver, PetscObjectComm((PetscObject)ns), PETSC_ERR_ARG_WRONGSTATE, "This function must be called after NSSetUp()");
  B200Call(ns, fluca_b200_stage_state(b->solver));
  b->staged = PETSC_TRUE;
  PetscFunctionReturn(PETSC_SUCCESS);
}

PetscErrorCode NSB200SyncSolution(NS ns)
{
  NS_B200 *b = (NS_B200 *)ns->data;

  PetscFunctionBegin;
  PetscCheck(ns->setupcalled && b->solver, PetscObjectComm((PetscObject)ns), PETSC_ERR_ARG_WRONGSTATE, "This function must be called after NSSetUp()");
  PetscCall(B200DeviceToHost_Private(ns)); /* waits for the staged copy (or stages now) and unpacks through DMStag */
  /* ns->sol may now be older than the device state: the next NSStep must not mistake it for a user edit */
  PetscFunctionReturn(PETSC_SUCCESS);
}

/* collective on the NS communicator: every marker's entry comes from the rank whose slab reports it */
PetscErrorCode NSB200GetMarkerForces(NS ns, PetscReal F[], PetscReal Um[])
{
  NS_B200 *b = (NS_B200 *)ns->data;

  PetscFunctionBegin;
  B200Call(ns, fluca_b200_get_marker_forces(b->solver, F, Um));
  PetscFunctionReturn(PETSC_SUCCESS);
}
