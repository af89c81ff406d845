/* tests/petsc_stub/petsc_fluca_mock.h -- TEST INFRASTRUCTURE ONLY.
 * What a driver needs on top of petsc_fluca_stub.h to build a Mesh and an NS object, step it and read the solution back
 * through the functional model in petsc_fluca_mock.c.  Function names are the reference's (flucamesh.h, flucameshcart.h,
 * flucans.h) where the reference has the function; "Mock" in the name marks a helper that has no counterpart. */
#ifndef PETSC_FLUCA_MOCK_H
#define PETSC_FLUCA_MOCK_H
#include <stdio.h>
#include "petsc_fluca_stub.h"

/* options database: MockOptionsSetValue("-ns_b200_mode", "1") before NSSetFromOptions */
PetscErrorCode MockOptionsSetValue(const char name[], const char value[]);
PetscErrorCode MockOptionsClear(void);
/* VecRestoreSubVector on a VecNest increases the state of the nest (PETSc does; 0 models an implementation that only does so when
 * the sub-vector changed).  The glue's coherence logic must be right under both. */
void MockSetNestRestoreBumpsState(int on);
/* counts calls of the boundary callbacks / uploads of the whole state, for the tests of the caches */
extern long mock_ndm_global_to_local;

/* Mesh: one rank, `dim` directions, N cells and N + 1 face coordinates per direction (cart.c:56-151 makes the same DMs) */
PetscErrorCode MockMeshCartCreate(PetscInt dim, const PetscInt N[], const PetscBool periodic[], const double *const xf[], Mesh *mesh);
PetscErrorCode MeshDestroy(Mesh *mesh);

/* NS base class (nsbasic.c, nsopts.c, nssol.c): the order of operations of NSSetUp / NSStep / NSViewSolution / NSLoadSolution */
PetscErrorCode NSCreate(MPI_Comm comm, NS *ns);
PetscErrorCode NSSetType(NS ns, const char type[]);
PetscErrorCode NSSetMesh(NS ns, Mesh mesh);
PetscErrorCode NSSetDensity(NS ns, PetscReal rho);
PetscErrorCode NSSetViscosity(NS ns, PetscReal mu);
PetscErrorCode NSSetTimeStepSize(NS ns, PetscReal dt);
PetscErrorCode NSSetBoundaryCondition(NS ns, PetscInt index, NSBoundaryCondition bc);
PetscErrorCode NSSetFromOptions(NS ns);
PetscErrorCode NSSetUp(NS ns);
PetscErrorCode NSStep(NS ns);
PetscErrorCode NSGetSolution(NS ns, Vec *sol);
PetscErrorCode NSFormFunction(NS ns, Vec x, Vec f);
PetscErrorCode MockNSCreateVecs(NS ns, Vec *x, Vec *f); /* MatCreateVecs(ns->J, ...): nests like ns->sol */
PetscErrorCode NSView(NS ns, PetscViewer viewer);
PetscErrorCode NSViewSolution(NS ns, PetscViewer viewer);
PetscErrorCode NSLoadSolution(NS ns, PetscViewer viewer);
PetscErrorCode NSDestroy(NS *ns);

/* the entry points of glue/nsb200.c that are not NSOps: the type's public header, and the plugin hook PETSc looks up by name */
#include "../../glue/flucansb200.h"
PetscErrorCode PetscDLLibraryRegister_fluca_nsb200(void);

/* Vec helpers a driver uses */
PetscErrorCode VecSet(Vec v, PetscScalar a);
PetscErrorCode VecCopy(Vec x, Vec y);
PetscErrorCode VecDuplicate(Vec x, Vec *y);

/* viewers: "ascii" prints to a FILE; "mockstore" keeps named copies of the vectors viewed into it (the CGNS file of a restart) */
PetscErrorCode MockViewerASCIIOpen(FILE *f, PetscViewer *viewer);
PetscErrorCode MockViewerStoreOpen(PetscViewer *viewer);
PetscErrorCode PetscViewerDestroy(PetscViewer *viewer);

/* live allocations of the mock (objects + PetscMalloc): a driver that destroyed everything sees 0 */
long MockLiveAllocations(void);
#endif
