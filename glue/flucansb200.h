/*
 * glue/flucansb200.h -- public header of the NS type "b200" (install next to the reference's flucans.h as fluca/include/flucansb200.h).
 *
 * Everything an application needs beyond the reference's own NS API (flucans.h): the type name for NSSetType / -ns_type, and the
 * four entry points that have no counterpart in the reference -- immersed-boundary markers (the reference only plans IBM:
 * README.md:14, THEORY_GUIDE.md:130-132) and the asynchronous solution view for runs that keep ns->sol on the device between outputs
 * (-ns_b200_sync_interval 0).  Implemented in glue/nsb200.c.
 */
#pragma once

#include <flucans.h>

#define NSB200 "b200"

/* n markers: X and Ud are [component][marker] (dim x n), dV the marker volumes; delta_points 3 (Roma) or 4 (Peskin).  Call after
 * NSSetUp; collective.  Without markers the step is exactly the reference's scheme. */
FLUCA_EXTERN PetscErrorCode NSB200SetMarkers(NS ns, PetscInt n, const PetscReal X[], const PetscReal Ud[], const PetscReal dV[], PetscInt delta_points);
/* forces on the fluid and interpolated marker velocities of the last step, [component][marker]; collective */
FLUCA_EXTERN PetscErrorCode NSB200GetMarkerForces(NS ns, PetscReal F[], PetscReal Um[]);
/* start copying the current device state to pinned host memory behind the steps that follow ... */
FLUCA_EXTERN PetscErrorCode NSB200StageSolution(NS ns);
/* ... and make ns->sol (and the type's PressureHalfStep) that state, e.g. right before NSViewSolution */
FLUCA_EXTERN PetscErrorCode NSB200SyncSolution(NS ns);
/* the type's constructor, for NSRegister(NSB200, NSCreate_B200) in NSRegisterAll (nsreg.c:13-20); the out-of-tree route needs no
 * declaration at all: -dll_append libfluca_nsb200.so calls PetscDLLibraryRegister_fluca_nsb200 */
FLUCA_EXTERN PetscErrorCode NSCreate_B200(NS ns);
