/* tests/petsc_stub/petsc_fluca_stub.h -- TEST INFRASTRUCTURE ONLY.
 * Stand-in for the subset of the PETSc (>= 3.23) and Fluca C APIs that glue/nsb200.c uses, in an image without PETSc:
 * tests/test_glue_syntax.py runs `gcc -fsyntax-only` on the glue against these declarations, and tests/test_glue_mock.py
 * links the glue with the single-rank functional model of them in petsc_fluca_mock.c and RUNS it.  Signatures follow the
 * PETSc manual pages and fluca/include/ *.h. */
#ifndef PETSC_FLUCA_STUB_H
#define PETSC_FLUCA_STUB_H
#include <stddef.h>
#include <stdint.h>

typedef int    PetscErrorCode;
typedef int    PetscInt;
typedef int    PetscMPIInt;
typedef double PetscReal;
typedef double PetscScalar;
typedef int64_t PetscObjectState;
typedef enum { PETSC_FALSE, PETSC_TRUE } PetscBool;
typedef int    MPI_Comm;
typedef int    MPI_Datatype;
typedef int    MPI_Op;
/* the part of PETSc's object header the glue relies on (type name, state counter, communicator); first member of every mock object */
struct _p_PetscObject {
  const char      *class_name;
  char             type_name[32];
  char             name[64];
  PetscObjectState state;
  MPI_Comm         comm;
  int              refct;
};
typedef struct _p_PetscObject *PetscObject;
typedef struct _p_Vec *Vec;
typedef struct _p_Mat *Mat;
typedef struct _p_DM *DM;
typedef struct _p_IS *IS;
typedef struct _p_SNES *SNES;
typedef struct _p_KSP *KSP;
typedef struct _p_PC *PC;
typedef struct _p_PetscViewer *PetscViewer;
typedef struct _p_MatNullSpace *MatNullSpace;
typedef struct _p_PetscOptionItems *PetscOptionItems;
typedef struct _n_PetscFunctionList *PetscFunctionList;
typedef struct _p_PetscViewerAndFormat PetscViewerAndFormat;
typedef int PetscClassId;
typedef int PetscLogEvent;
#define PetscLogEventRegister(name, classid, e) (*(e) = 1, PETSC_SUCCESS) /* -log_view is PETSc's; the mock has no profiler */
#define PetscLogEventBegin(e, a, b, c, d) PETSC_SUCCESS
#define PetscLogEventEnd(e, a, b, c, d) PETSC_SUCCESS
typedef enum { INSERT_VALUES = 1, ADD_VALUES = 2 } InsertMode;
typedef enum { MAT_FLUSH_ASSEMBLY = 1, MAT_FINAL_ASSEMBLY = 0 } MatAssemblyType;
typedef enum { DMSTAG_NULL_LOCATION, DMSTAG_BACK_DOWN_LEFT, DMSTAG_BACK_DOWN, DMSTAG_BACK_DOWN_RIGHT, DMSTAG_BACK_LEFT, DMSTAG_BACK, DMSTAG_BACK_RIGHT, DMSTAG_BACK_UP_LEFT, DMSTAG_BACK_UP, DMSTAG_BACK_UP_RIGHT, DMSTAG_DOWN_LEFT, DMSTAG_DOWN, DMSTAG_DOWN_RIGHT, DMSTAG_LEFT, DMSTAG_ELEMENT, DMSTAG_RIGHT } DMStagStencilLocation;

#define PETSC_SUCCESS 0
#define PETSC_ERR_LIB 76
#define PETSC_ERR_SUP 56
#define PETSC_ERR_ARG_WRONG 62
#define PETSC_ERR_ARG_WRONGSTATE 73
#define PETSC_DETERMINE (-1)
#define PETSC_EXTERN extern
#define FLUCA_EXTERN extern
#define PETSCVIEWERASCII "ascii"
#define PetscInt_FMT "d"
#define MPI_IN_PLACE ((void *)1)
#define MPI_DOUBLE 1
#define MPI_INT 2
#define MPI_BYTE 3
#define MPI_SUM 1
#define PetscFunctionBegin
#define PetscFunctionReturn(x) return (x)
#define PetscCall(...) \
  do { \
    PetscErrorCode ierr_ = (__VA_ARGS__); \
    if (ierr_) return ierr_; \
  } while (0)
#define PetscCallMPI(...) PetscCall(__VA_ARGS__)
PetscErrorCode PetscErrorStub(MPI_Comm, int, const char *, ...);
#define PetscCheck(cond, comm, err, ...) \
  do { \
    if (!(cond)) return PetscErrorStub(comm, err, __VA_ARGS__); \
  } while (0)
#define PetscMax(a, b) ((a) > (b) ? (a) : (b))
#define PetscRealPart(a) (a)
PetscErrorCode PetscMallocStub(size_t, void *);
PetscErrorCode PetscCallocStub(size_t, void *);
#define PetscMalloc1(n, p) PetscMallocStub((size_t)(n) * sizeof(**(p)), (void *)(p))
#define PetscCalloc1(n, p) PetscCallocStub((size_t)(n) * sizeof(**(p)), (void *)(p))
#define PetscNew(p) PetscCallocStub(sizeof(**(p)), (void *)(p))
PetscErrorCode PetscFreeStub(void *);
#define PetscFree(p) (PetscFreeStub((void *)(p)) || ((p) = NULL, 0))
PetscErrorCode PetscMemzero(void *, size_t);
PetscErrorCode PetscInfoStub(void *, const char *, ...);
#define PetscInfo(obj, ...) PetscInfoStub((void *)(obj), __VA_ARGS__)
#define PetscOptionsHeadBegin(obj, head) (void)(obj)
#define PetscOptionsHeadEnd()
PetscErrorCode PetscOptionsIntStub(PetscOptionItems, const char *, const char *, const char *, PetscInt, PetscInt *, PetscBool *);
PetscErrorCode PetscOptionsRealStub(PetscOptionItems, const char *, const char *, const char *, PetscReal, PetscReal *, PetscBool *);
#define PetscOptionsInt(a, b, c, d, e, f) PetscOptionsIntStub(PetscOptionsObject, a, b, c, d, e, f)
#define PetscOptionsReal(a, b, c, d, e, f) PetscOptionsRealStub(PetscOptionsObject, a, b, c, d, e, f)
typedef int PetscEnum;
PetscErrorCode PetscOptionsEnumStub(PetscOptionItems, const char *, const char *, const char *, const char *const *, PetscEnum, PetscEnum *, PetscBool *);
#define PetscOptionsEnum(a, b, c, d, e, f, g) PetscOptionsEnumStub(PetscOptionsObject, a, b, c, d, e, f, g)

int MPI_Comm_rank(MPI_Comm, int *);
int MPI_Comm_size(MPI_Comm, int *);
int MPI_Bcast(void *, int, MPI_Datatype, int, MPI_Comm);
int MPI_Allreduce(const void *, void *, int, MPI_Datatype, MPI_Op, MPI_Comm);

MPI_Comm       PetscObjectComm(PetscObject);
PetscErrorCode PetscObjectGetComm(PetscObject, MPI_Comm *);
PetscErrorCode PetscObjectTypeCompare(PetscObject, const char[], PetscBool *);
PetscErrorCode PetscObjectStateGet(PetscObject, PetscObjectState *);
PetscErrorCode PetscObjectSetName(PetscObject, const char[]);
PetscErrorCode PetscViewerASCIIPrintf(PetscViewer, const char[], ...);

PetscErrorCode VecDestroy(Vec *);
PetscErrorCode VecZeroEntries(Vec);
PetscErrorCode VecScale(Vec, PetscScalar);
PetscErrorCode VecView(Vec, PetscViewer);
PetscErrorCode VecGetSubVector(Vec, IS, Vec *);
PetscErrorCode VecRestoreSubVector(Vec, IS, Vec *);
PetscErrorCode MatCreateConstantDiagonal(MPI_Comm, PetscInt, PetscInt, PetscInt, PetscInt, PetscScalar, Mat *);
PetscErrorCode MatNestSetSubMat(Mat, PetscInt, PetscInt, Mat);
PetscErrorCode MatDestroy(Mat *);
PetscErrorCode MatAssemblyBegin(Mat, MatAssemblyType);
PetscErrorCode MatAssemblyEnd(Mat, MatAssemblyType);

PetscErrorCode DMGetDimension(DM, PetscInt *);
PetscErrorCode DMGetLocalVector(DM, Vec *);
PetscErrorCode DMRestoreLocalVector(DM, Vec *);
PetscErrorCode DMGlobalToLocal(DM, Vec, InsertMode, Vec);
PetscErrorCode DMLocalToGlobal(DM, Vec, InsertMode, Vec);
PetscErrorCode DMStagGetCorners(DM, PetscInt *, PetscInt *, PetscInt *, PetscInt *, PetscInt *, PetscInt *, PetscInt *, PetscInt *, PetscInt *);
PetscErrorCode DMStagGetLocationSlot(DM, DMStagStencilLocation, PetscInt, PetscInt *);
PetscErrorCode DMStagGetEntries(DM, PetscInt *);
PetscErrorCode DMStagVecGetArray(DM, Vec, void *);
PetscErrorCode DMStagVecRestoreArray(DM, Vec, void *);
PetscErrorCode DMStagVecGetArrayRead(DM, Vec, void *);
PetscErrorCode DMStagVecRestoreArrayRead(DM, Vec, void *);
PetscErrorCode DMStagGetProductCoordinateArraysRead(DM, void *, void *, void *);
PetscErrorCode DMStagRestoreProductCoordinateArraysRead(DM, void *, void *, void *);
PetscErrorCode DMStagGetProductCoordinateLocationSlot(DM, DMStagStencilLocation, PetscInt *);

/* ---- Fluca (fluca/include/flucamesh.h, flucameshcart.h, flucans.h, flucansbc.h, flucaviewer.h, private/nsimpl.h) ---- */
typedef struct _p_Mesh *Mesh;
#define MESHCART "cart"
typedef enum { MESH_DM_SCALAR, MESH_DM_VECTOR, MESH_DM_STAG_SCALAR, MESH_DM_STAG_VECTOR } MeshDMType;
PetscErrorCode MeshGetDimension(Mesh, PetscInt *);
PetscErrorCode MeshGetDM(Mesh, MeshDMType, DM *);
PetscErrorCode MeshCreateGlobalVector(Mesh, MeshDMType, Vec *);
PetscErrorCode MeshGetNumberBoundaries(Mesh, PetscInt *);
typedef enum { MESHCART_BOUNDARY_NONE, MESHCART_BOUNDARY_PERIODIC } MeshCartBoundaryType;
typedef enum { MESHCART_PREV, MESHCART_NEXT } MeshCartCoordinateStencilLocation;
PetscErrorCode MeshCartGetGlobalSizes(Mesh, PetscInt *, PetscInt *, PetscInt *);
PetscErrorCode MeshCartGetNumRanks(Mesh, PetscInt *, PetscInt *, PetscInt *);
PetscErrorCode MeshCartGetBoundaryTypes(Mesh, MeshCartBoundaryType *, MeshCartBoundaryType *, MeshCartBoundaryType *);
PetscErrorCode MeshCartGetCoordinateArraysRead(Mesh, const PetscScalar ***, const PetscScalar ***, const PetscScalar ***);
PetscErrorCode MeshCartRestoreCoordinateArraysRead(Mesh, const PetscScalar ***, const PetscScalar ***, const PetscScalar ***);
PetscErrorCode MeshCartGetCoordinateLocationSlot(Mesh, MeshCartCoordinateStencilLocation, PetscInt *);
PetscErrorCode MeshCartGetCorners(Mesh, PetscInt *, PetscInt *, PetscInt *, PetscInt *, PetscInt *, PetscInt *);
PetscErrorCode MeshCartGetIsLastRank(Mesh, PetscBool *, PetscBool *, PetscBool *);
PetscErrorCode MeshCartGetIsFirstRank(Mesh, PetscBool *, PetscBool *, PetscBool *);

typedef enum { NS_BC_NONE, NS_BC_VELOCITY, NS_BC_PRESSURE_OUTLET, NS_BC_PERIODIC, NS_BC_SYMMETRY } NSBoundaryConditionType;
typedef PetscErrorCode (*NSBoundaryConditionFunction)(PetscInt, PetscReal, const PetscReal[], PetscScalar[], void *);
typedef struct {
  NSBoundaryConditionType     type;
  NSBoundaryConditionFunction velocity;
  void                       *ctx_velocity;
  NSBoundaryConditionFunction pressure;
  void                       *ctx_pressure;
} NSBoundaryCondition;

typedef struct _p_NS *NS;
typedef enum { NS_CONVERGED_ITERATING = 0, NS_CONVERGED_TIME = 1, NS_CONVERGED_ITS = 2, NS_DIVERGED_NONLINEAR_SOLVE = -1 } NSConvergedReason;
typedef enum { NS_INIT_JACOBIAN, NS_UPDATE_JACOBIAN } NSFormJacobianType;
#define NS_FIELD_VELOCITY "Velocity"
#define NS_FIELD_FACE_NORMAL_VELOCITY "FaceNormalVelocity"
#define NS_FIELD_PRESSURE "Pressure"
PetscErrorCode NSRegister(const char[], PetscErrorCode (*)(NS));
typedef enum { PC_ABF_AINV_ID, PC_ABF_AINV_DIAG, PC_ABF_AINV_ROWSUM } PCABFAinvType; /* flucans.h:99-103 */
extern const char *const PCABFAinvTypes[];                                             /* flucans.h:104 */
PetscErrorCode NSGetField(NS, const char[], PetscInt *, MeshDMType *, IS *);
PetscErrorCode NSGetSolutionSubVector(NS, const char[], Vec *);
PetscErrorCode NSRestoreSolutionSubVector(NS, const char[], Vec *);
PetscErrorCode FlucaVecLoad(Vec, PetscViewer);

struct _NSOps {
  PetscErrorCode (*setfromoptions)(NS, PetscOptionItems);
  PetscErrorCode (*setup)(NS);
  PetscErrorCode (*step)(NS);
  PetscErrorCode (*formjacobian)(NS, Vec, Mat, NSFormJacobianType);
  PetscErrorCode (*formfunction)(NS, Vec, Vec);
  PetscErrorCode (*destroy)(NS);
  PetscErrorCode (*view)(NS, PetscViewer);
  PetscErrorCode (*viewsolution)(NS, PetscViewer);
  PetscErrorCode (*loadsolution)(NS, PetscViewer);
};
/* the members of struct _p_NS that a type implementation touches (nsimpl.h:41-82) */
struct _p_NS {
  struct _p_PetscObject hdr;
  struct _NSOps        ops[1];
  PetscReal            rho, mu, dt, max_time;
  PetscInt             max_steps, step;
  PetscReal            t;
  Mesh                 mesh;
  NSBoundaryCondition *bcs;
  void                *data;
  Vec                  sol, sol0;
  NSConvergedReason    reason;
  PetscBool            setupcalled;
#ifdef FLUCA_NS_HAS_MATRIXFREE
  PetscBool            matrixfree; /* glue/patches/0001-ns-matrix-free-type-hooks.patch */
#endif
};
/* ---- additions for the hardened glue: option flags, formatted output, array comparison ---- */
PetscErrorCode PetscOptionsBoolStub(PetscOptionItems, const char *, const char *, const char *, PetscBool, PetscBool *, PetscBool *);
#define PetscOptionsBool(a, b, c, d, e, f) PetscOptionsBoolStub(PetscOptionsObject, a, b, c, d, e, f)
PetscErrorCode PetscPrintf(MPI_Comm, const char[], ...);
PetscErrorCode PetscArraycmpStub(const void *, const void *, size_t, PetscBool *);
#define PetscArraycmp(a, b, n, e) PetscArraycmpStub((a), (b), (size_t)(n) * sizeof(*(a)), (e))
PetscErrorCode PetscArraycpyStub(void *, const void *, size_t);
#define PetscArraycpy(a, b, n) PetscArraycpyStub((a), (b), (size_t)(n) * sizeof(*(a)))
#endif
