/* oracle/ref_model/include/petsc_model.h -- TEST INFRASTRUCTURE ONLY (never part of the product, never linked into it).
 *
 * A single-rank functional MODEL of the PETSc API subset that the reference's Navier-Stokes sources use, so that those sources --
 * fluca/src/ns/utils/cartdiscret.c, fluca/src/ns/impl/linearcn/{cnlinear,cnlinearcart2d,cnlinearcart3d}.c and
 * fluca/src/ns/utils/abfpc/abfpc.c, with the reference's own headers under fluca/include -- can be compiled FROM WHERE THEY LIE under
 * /root/reference into oracle/_ref/ (oracle/Makefile, target `ref`) and run as the checker of this repository's oracle.  PETSc itself
 * (>= 3.23) is absent from the image; nothing here is PETSc source.  Semantics follow the PETSc manual pages:
 *   DMStag        element-wise storage (vertex, edges, faces, element), partial elements at the upper end of non-periodic
 *                 directions, ghost elements of width 1 in periodic directions, local indices through DMStagStencilToIndexLocal,
 *                 local-to-global maps that wrap periodic ghosts and give -1 for entries that do not exist (ignored on insertion)
 *   Vec / VecNest dense arrays; sub-vectors of a nest selected by the index set of a field
 *   Mat           row lists (AIJ semantics: INSERT_VALUES replaces, ADD_VALUES accumulates, explicit zeros are kept), MatNest 3 x 3
 *   KSP           KSPSolve is an EXACT solve (dense LU with partial pivoting; a constant null space is handled by bordering), i.e.
 *                 the limit the reference's GMRES + ILU(0) converges to; PETSc's iteration histories are NOT modelled
 *   SNES          SNESSolve of the Picard form the base class sets up (nsbasic.c:249-253): zero guess, b = ops->formfunction,
 *                 J = ops->formjacobian(UPDATE), null space removed from b, then one of three linear solves (ref_driver.c)
 */
#ifndef PETSC_MODEL_H
#define PETSC_MODEL_H
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

/* ---- scalars, enums ---- */
typedef int     PetscErrorCode;
typedef int     PetscInt;
typedef int     PetscMPIInt;
typedef double  PetscReal;
typedef double  PetscScalar;
typedef int64_t PetscObjectState;
typedef int     PetscEnum;
typedef int     PetscClassId;
typedef int     PetscLogEvent;
typedef int     MPI_Comm;
typedef enum { PETSC_FALSE, PETSC_TRUE } PetscBool;
typedef enum { NOT_SET_VALUES, INSERT_VALUES, ADD_VALUES } InsertMode;
typedef enum { MAT_FLUSH_ASSEMBLY = 1, MAT_FINAL_ASSEMBLY = 0 } MatAssemblyType;
typedef enum { MAT_INITIAL_MATRIX, MAT_REUSE_MATRIX, MAT_IGNORE_MATRIX, MAT_INPLACE_MATRIX } MatReuse;
typedef enum { DIFFERENT_NONZERO_PATTERN, SUBSET_NONZERO_PATTERN, SAME_NONZERO_PATTERN, UNKNOWN_NONZERO_PATTERN } MatStructure;
typedef enum { MAT_DO_NOT_COPY_VALUES, MAT_COPY_VALUES, MAT_SHARE_NONZERO_PATTERN } MatDuplicateOption;
typedef enum { MAT_NEW_NONZERO_ALLOCATION_ERR = 19, MAT_NEW_NONZERO_LOCATION_ERR = 11 } MatOption;
typedef enum { DM_BOUNDARY_NONE, DM_BOUNDARY_GHOSTED, DM_BOUNDARY_MIRROR, DM_BOUNDARY_PERIODIC } DMBoundaryType;
typedef enum {
  DMSTAG_NULL_LOCATION = 0, DMSTAG_BACK_DOWN_LEFT, DMSTAG_BACK_DOWN, DMSTAG_BACK_DOWN_RIGHT, DMSTAG_BACK_LEFT, DMSTAG_BACK, DMSTAG_BACK_RIGHT, DMSTAG_BACK_UP_LEFT, DMSTAG_BACK_UP,
  DMSTAG_BACK_UP_RIGHT, DMSTAG_DOWN_LEFT, DMSTAG_DOWN, DMSTAG_DOWN_RIGHT, DMSTAG_LEFT, DMSTAG_ELEMENT, DMSTAG_RIGHT, DMSTAG_UP_LEFT, DMSTAG_UP, DMSTAG_UP_RIGHT, DMSTAG_FRONT_DOWN_LEFT,
  DMSTAG_FRONT_DOWN, DMSTAG_FRONT_DOWN_RIGHT, DMSTAG_FRONT_LEFT, DMSTAG_FRONT, DMSTAG_FRONT_RIGHT, DMSTAG_FRONT_UP_LEFT, DMSTAG_FRONT_UP, DMSTAG_FRONT_UP_RIGHT
} DMStagStencilLocation;
typedef struct {
  DMStagStencilLocation loc;
  PetscInt              i, j, k, c;
} DMStagStencil;
typedef const char *MatType;
typedef const char *VecType;
#define MATNEST "nest"
#define MATAIJ "aij"
#define VECNEST "nest"
#define PETSCVIEWERASCII "ascii"

#define PETSC_SUCCESS 0
#define PETSC_ERR_SUP 56
#define PETSC_ERR_ARG_WRONG 62
#define PETSC_ERR_ARG_WRONGSTATE 73
#define PETSC_ERR_ARG_OUTOFRANGE 63
#define PETSC_ERR_LIB 76
#define PETSC_ERR_PLIB 77
#define PETSC_ERR_ORDER 58
#define PETSC_ERROR_INITIAL 0
#define PETSC_DETERMINE (-1)
#define PETSC_DECIDE (-1)
#define PETSC_DEFAULT (-2)
#define PETSC_COMM_SELF 1
#define PETSC_COMM_WORLD 2
#define PETSC_MAX_REAL 1.7976931348623157e308
#define PETSC_FUNCTION_NAME __func__
#define PETSC_EXTERN extern
#define PETSC_INTERN extern
#define PetscInt_FMT "d"

/* ---- objects ---- */
struct composed_object;
struct composed_function;
struct _p_PetscObject {
  PetscClassId              classid;
  const char               *class_name;
  char                     *type_name;
  char                     *name;
  char                     *prefix;
  void                     *options;
  PetscObjectState          state;
  MPI_Comm                  comm;
  int                       refct;
  int                       tablevel;
  struct composed_object   *olist;
  struct composed_function *flist;
  PetscErrorCode (*destroy_model)(struct _p_PetscObject *);  /* how PetscObjectDereference frees a composed object of the model ... */
  PetscErrorCode (*destroy_public)(struct _p_PetscObject **); /* ... and one made by PetscHeaderCreate (its XxxDestroy(Xxx *)) */
};
typedef struct _p_PetscObject *PetscObject;
#define PETSCHEADER(ObjectOps) \
  struct _p_PetscObject hdr; \
  ObjectOps             ops[1]

typedef struct _p_Vec                    *Vec;
typedef struct _p_Mat                    *Mat;
typedef struct _p_DM                     *DM;
typedef struct _p_IS                     *IS;
typedef struct _p_SNES                   *SNES;
typedef struct _p_KSP                    *KSP;
typedef struct _p_PC                     *PC;
typedef struct _p_PetscViewer            *PetscViewer;
typedef struct _p_MatNullSpace           *MatNullSpace;
typedef struct _p_ISLocalToGlobalMapping *ISLocalToGlobalMapping;
typedef struct _p_PetscOptionItems       *PetscOptionItems;
typedef struct _n_PetscFunctionList      *PetscFunctionList;
typedef struct _p_PetscViewerAndFormat    PetscViewerAndFormat;
typedef struct _n_PetscOptions           *PetscOptions;
typedef int                               PetscViewerFormat;
typedef enum { FILE_MODE_UNDEFINED = -1, FILE_MODE_READ = 0, FILE_MODE_WRITE, FILE_MODE_APPEND, FILE_MODE_UPDATE, FILE_MODE_APPEND_UPDATE } PetscFileMode;

/* petsc/private/pcimpl.h: what a PC implementation touches */
struct _PCOps {
  PetscErrorCode (*setup)(PC);
  PetscErrorCode (*apply)(PC, Vec, Vec);
  PetscErrorCode (*reset)(PC);
  PetscErrorCode (*destroy)(PC);
  PetscErrorCode (*setfromoptions)(PC, PetscOptionItems);
  PetscErrorCode (*view)(PC, PetscViewer);
};
struct _p_PC {
  PETSCHEADER(struct _PCOps);
  Mat   mat, pmat;
  void *data;
  int   setupcalled;
};
extern PetscClassId PC_CLASSID, VEC_CLASSID, MAT_CLASSID;

/* ---- control flow, errors, memory ---- */
#define PetscFunctionBegin
#define PetscFunctionReturn(x) return (x)
#define PetscCall(...) \
  do { \
    PetscErrorCode ierr_q_ = (__VA_ARGS__); \
    if (ierr_q_) return ModelErrorTrace(ierr_q_, __FILE__, __LINE__, __func__); \
  } while (0)
#define PetscCallMPI(...) PetscCall(__VA_ARGS__)
PetscErrorCode ModelError(MPI_Comm, int, const char *, int, const char *, ...);
PetscErrorCode ModelErrorTrace(PetscErrorCode, const char *, int, const char *);
#define SETERRQ(comm, err, ...) return ModelError(comm, err, __FILE__, __LINE__, __VA_ARGS__)
#define PetscCheck(cond, comm, err, ...) \
  do { \
    if (!(cond)) return ModelError(comm, err, __FILE__, __LINE__, __VA_ARGS__); \
  } while (0)
#define PetscAssert(cond, comm, err, ...) PetscCheck(cond, comm, err, __VA_ARGS__)
#define PetscError(comm, line, fn, file, err, kind, ...) ModelError(comm, err, file, line, __VA_ARGS__)
#define PetscValidHeaderSpecific(obj, classid, arg) (void)(obj)
#define PetscValidHeader(obj, arg) (void)(obj)
#define PetscAssertPointer(p, arg) (void)(p)
#define PetscCheckSameComm(a, ia, b, ib) (void)0
#define PetscMax(a, b) ((a) > (b) ? (a) : (b))
#define PetscMin(a, b) ((a) < (b) ? (a) : (b))
#define PetscRealPart(a) (a)
#define PetscSqrtReal(a) sqrt(a)
#define PetscAbsReal(a) fabs(a)
#define PetscAbsScalar(a) fabs(a)
#define PETSC_PI 3.14159265358979323846
PetscErrorCode ModelMalloc(size_t, int zero, void *);
PetscErrorCode ModelFree(void *);
#define PetscMalloc1(n, p) ModelMalloc((size_t)(n) * sizeof(**(p)), 0, (void *)(p))
#define PetscCalloc1(n, p) ModelMalloc((size_t)(n) * sizeof(**(p)), 1, (void *)(p))
#define PetscNew(p) ModelMalloc(sizeof(**(p)), 1, (void *)(p))
#define PetscMalloc2(n1, p1, n2, p2) (ModelMalloc((size_t)(n1) * sizeof(**(p1)), 0, (void *)(p1)) || ModelMalloc((size_t)(n2) * sizeof(**(p2)), 0, (void *)(p2)))
#define PetscFree(p) (ModelFree((void *)(p)) || ((p) = NULL, 0))
#define PetscFree2(p1, p2) (ModelFree((void *)(p1)) || ModelFree((void *)(p2)) || ((p1) = NULL, (p2) = NULL, 0))
PetscErrorCode PetscMemzero(void *, size_t);
PetscErrorCode PetscSNPrintf(char *, size_t, const char[], ...);
PetscErrorCode PetscStrcmp(const char[], const char[], PetscBool *);
PetscErrorCode PetscStrallocpy(const char[], char **);
PetscErrorCode PetscInfoModel(void *, const char *, ...);
#define PetscInfo(obj, ...) PetscInfoModel((void *)(obj), __VA_ARGS__)
/* log events: timed (wall clock, inclusive) when PETSC_MODEL_TIMING is set and listed by PetscFinalize, otherwise free */
PetscErrorCode ModelLogEventBegin(PetscLogEvent e);
PetscErrorCode ModelLogEventEnd(PetscLogEvent e);
#define PetscLogEventBegin(e, a, b, c, d) ModelLogEventBegin(e)
#define PetscLogEventEnd(e, a, b, c, d) ModelLogEventEnd(e)
#define PetscArraycpy(a, b, n) (memcpy((a), (b), (size_t)(n) * sizeof(*(a))), PETSC_SUCCESS)

PetscErrorCode PetscPrintf(MPI_Comm, const char[], ...);
#define PetscArraycmp(a, b, n, e) (*(e) = memcmp((a), (b), (size_t)(n) * sizeof(*(a))) ? PETSC_FALSE : PETSC_TRUE, PETSC_SUCCESS)
/* MPI: one rank */
typedef int MPI_Datatype;
typedef int MPI_Op;
#define MPI_IN_PLACE ((void *)1)
#define MPI_DOUBLE 1
#define MPI_INT 2
#define MPI_BYTE 3
#define MPI_SUM 1
int MPI_Comm_rank(MPI_Comm, int *);
int MPI_Comm_size(MPI_Comm, int *);
int MPI_Bcast(void *, int, MPI_Datatype, int, MPI_Comm);
int MPI_Allreduce(const void *, void *, int, MPI_Datatype, MPI_Op, MPI_Comm);

/* options database (filled from argv by PetscInitialize, or by ModelOptionsSetValue) */
PetscErrorCode ModelOptionsSetValue(const char name[], const char value[]);
PetscErrorCode ModelOptionsClear(void);
PetscErrorCode ModelOptionsReal(const char name[], PetscReal *val, PetscBool *set);
PetscErrorCode ModelOptionsInt(const char name[], PetscInt *val, PetscBool *set);
PetscErrorCode ModelOptionsBool(const char name[], PetscBool *val, PetscBool *set);
PetscErrorCode ModelOptionsString(const char name[], char *val, size_t len, PetscBool *set);
PetscErrorCode ModelOptionsEnum(const char name[], const char *const *list, PetscEnum *val, PetscBool *set);
void ModelOptionsPrefixPush(const char *prefix);
void ModelOptionsPrefixPop(void);
const char *ModelOptionsPrefixGet(void);
#define PetscObjectOptionsBegin(obj) \
  { \
    PetscOptionItems PetscOptionsObject = (PetscOptionItems)(obj); \
    (void)PetscOptionsObject; \
    ModelOptionsPrefixPush(((PetscObject)(obj))->prefix);
#define PetscOptionsEnd() \
  ModelOptionsPrefixPop(); \
  }
#define PetscOptionsHeadBegin(obj, head) (void)(obj)
#define PetscOptionsHeadEnd()
#define PetscOptionsEnum(name, text, man, list, cur, val, set) ModelOptionsEnum(name, list, val, set)
#define PetscOptionsInt(name, text, man, cur, val, set) ModelOptionsInt(name, val, set)
#define PetscOptionsBoundedInt(name, text, man, cur, val, set, bound) ModelOptionsInt(name, val, set)
#define PetscOptionsReal(name, text, man, cur, val, set) ModelOptionsReal(name, val, set)
#define PetscOptionsBool(name, text, man, cur, val, set) ModelOptionsBool(name, val, set)
#define PetscOptionsFList(name, text, man, list, def, val, len, set) ModelOptionsString(name, val, len, set)
#define PetscOptionsGetReal(opts, pre, name, val, set) ModelOptionsReal(name, val, set)
#define PetscOptionsGetInt(opts, pre, name, val, set) ModelOptionsInt(name, val, set)
#define PetscOptionsGetBool(opts, pre, name, val, set) ModelOptionsBool(name, val, set)
#define PetscOptionsGetString(opts, pre, name, val, len, set) ModelOptionsString(name, val, len, set)
PetscErrorCode PetscStrInList(const char[], const char[], char, PetscBool *);
#define PETSC_MAX_PATH_LEN 4096
#define PETSC_MAX_OPTION_NAME 512
#define PETSC_INT_MAX 2147483647
#define PETSC_ERR_ARG_UNKNOWN_TYPE 86
#define PETSC_ERR_NOT_CONVERGED 82
#define PetscFunctionBeginUser
#define PetscSinReal(a) sin(a)
#define PetscCosReal(a) cos(a)
#define PetscExpReal(a) exp(a)
#include <math.h>

/* ---- program, packages, classes, function lists ---- */
PetscErrorCode PetscInitialize(int *, char ***, const char[], const char[]);
PetscErrorCode PetscInitialized(PetscBool *);
PetscErrorCode PetscFinalize(void);
PetscErrorCode PetscFinalized(PetscBool *);
PetscErrorCode PetscRegisterFinalize(PetscErrorCode (*)(void));
PetscErrorCode PetscClassIdRegister(const char[], PetscClassId *);
PetscErrorCode PetscLogEventRegister(const char[], PetscClassId, PetscLogEvent *);
PetscErrorCode PetscInfoProcessClass(const char[], PetscInt, PetscClassId[]);
PetscErrorCode PetscLogEventExcludeClass(PetscClassId);
PetscErrorCode ModelFunctionListAdd(PetscFunctionList *, const char[], void (*)(void));
PetscErrorCode ModelFunctionListFind(PetscFunctionList, const char[], void (**)(void));
PetscErrorCode PetscFunctionListDestroy(PetscFunctionList *);
#define PetscFunctionListAdd(list, name, f) ModelFunctionListAdd(list, name, (void (*)(void))(f))
#define PetscFunctionListFind(list, name, f) ModelFunctionListFind(list, name, (void (**)(void))(f))
PetscErrorCode ModelHeaderCreate(void *pobj, size_t size, PetscClassId, const char cls[], MPI_Comm, PetscErrorCode (*destroy)(struct _p_PetscObject **));
PetscErrorCode ModelHeaderDestroy(void *pobj);
#define PetscHeaderCreate(h, classid, class_name, descr, mansec, comm, destroy, view) ModelHeaderCreate((void *)&(h), sizeof(*(h)), classid, class_name, comm, (PetscErrorCode(*)(struct _p_PetscObject **))(destroy))
#define PetscHeaderDestroy(h) ModelHeaderDestroy((void *)(h))
PetscErrorCode PetscObjectChangeTypeName(PetscObject, const char[]);
PetscErrorCode PetscObjectGetName(PetscObject, const char *[]);
PetscErrorCode PetscObjectPrintClassNamePrefixType(PetscObject, PetscViewer);
PetscErrorCode PetscObjectDereference(PetscObject);
#define PetscUseTypeMethod(obj, method, ...) \
  do { \
    PetscCheck((obj)->ops->method, 0, PETSC_ERR_SUP, "No method %s for %s of type %s", #method, ((PetscObject)(obj))->class_name, ((PetscObject)(obj))->type_name); \
    PetscCall((*(obj)->ops->method)(obj __VA_OPT__(, ) __VA_ARGS__)); \
  } while (0)
#define PetscTryTypeMethod(obj, method, ...) \
  do { \
    if ((obj)->ops->method) PetscCall((*(obj)->ops->method)(obj __VA_OPT__(, ) __VA_ARGS__)); \
  } while (0)

/* ---- PetscObject ---- */
MPI_Comm       PetscObjectComm(PetscObject);
PetscErrorCode PetscObjectGetComm(PetscObject, MPI_Comm *);
PetscErrorCode PetscObjectTypeCompare(PetscObject, const char[], PetscBool *);
PetscErrorCode PetscObjectStateGet(PetscObject, PetscObjectState *);
PetscErrorCode PetscObjectSetName(PetscObject, const char[]);
PetscErrorCode PetscObjectReference(PetscObject);
PetscErrorCode PetscObjectCompose(PetscObject, const char[], PetscObject);
PetscErrorCode PetscObjectQuery(PetscObject, const char[], PetscObject *);
PetscErrorCode PetscObjectIncrementTabLevel(PetscObject, PetscObject, PetscInt);
PetscErrorCode PetscObjectSetOptions(PetscObject, void *);
PetscErrorCode PetscObjectGetOptionsPrefix(PetscObject, const char *[]);
PetscErrorCode PetscObjectComposeFunctionModel(PetscObject, const char[], void (*)(void));
PetscErrorCode PetscObjectQueryFunctionModel(PetscObject, const char[], void (**)(void));
#define PetscObjectComposeFunction(obj, name, f) PetscObjectComposeFunctionModel(obj, name, (void (*)(void))(f))
#define PetscTryMethod(obj, name, argtypes, args) \
  do { \
    PetscErrorCode(*f_q_) argtypes = NULL; \
    PetscCall(PetscObjectQueryFunctionModel((PetscObject)(obj), name, (void (**)(void)) & f_q_)); \
    if (f_q_) PetscCall((*f_q_)args); \
  } while (0)
#define PetscUseMethod(obj, name, argtypes, args) \
  do { \
    PetscErrorCode(*f_q_) argtypes = NULL; \
    PetscCall(PetscObjectQueryFunctionModel((PetscObject)(obj), name, (void (**)(void)) & f_q_)); \
    PetscCheck(f_q_, 0, PETSC_ERR_SUP, "no method %s", name); \
    PetscCall((*f_q_)args); \
  } while (0)

/* ---- viewer ---- */
PetscErrorCode PetscViewerASCIIPrintf(PetscViewer, const char[], ...);
PetscErrorCode PetscViewerASCIIPushTab(PetscViewer);
PetscErrorCode PetscViewerASCIIPopTab(PetscViewer);

struct _p_PetscViewerAndFormat {
  PetscViewer       viewer;
  PetscViewerFormat format;
  PetscInt          view_interval;
  void             *data;
};
typedef struct {
  int unused;
} *PetscSegBuffer;
typedef int PetscDataType;
PetscErrorCode PetscViewerAndFormatCreate(PetscViewer, PetscViewerFormat, PetscViewerAndFormat **);
PetscErrorCode PetscViewerAndFormatDestroy(PetscViewerAndFormat **);
PetscErrorCode PetscViewerPushFormat(PetscViewer, PetscViewerFormat);
PetscErrorCode PetscViewerPopFormat(PetscViewer);
PetscErrorCode PetscViewerFlush(PetscViewer);
PetscErrorCode PetscViewerDestroy(PetscViewer *);
PetscErrorCode PetscViewerCheckReadable(PetscViewer);
PetscErrorCode PetscViewerRegister(const char[], PetscErrorCode (*)(PetscViewer));
PetscErrorCode PetscViewerASCIIGetStdout(MPI_Comm, PetscViewer *);
PetscErrorCode PetscViewerASCIISynchronizedPrintf(PetscViewer, const char[], ...);
PetscErrorCode PetscViewerASCIIAddTab(PetscViewer, PetscInt);
PetscErrorCode PetscViewerASCIISubtractTab(PetscViewer, PetscInt);
PetscErrorCode PetscViewerASCIIPushSynchronized(PetscViewer);
PetscErrorCode PetscViewerASCIIPopSynchronized(PetscViewer);
PetscErrorCode PetscMonitorCompare(PetscErrorCode (*)(void), void *, PetscErrorCode (*)(void **), PetscErrorCode (*)(void), void *, PetscErrorCode (*)(void **), PetscBool *);

/* ---- Vec ---- */
typedef enum { NORM_1, NORM_2, NORM_INFINITY } NormType;
typedef enum { VECOP_VIEW = 33, VECOP_LOAD = 41 } VecOperation;
PetscErrorCode VecNorm(Vec, NormType, PetscReal *);
PetscErrorCode VecSetOperation(Vec, VecOperation, void (*)(void));
PetscErrorCode VecCreateNest(MPI_Comm, PetscInt, IS[], Vec[], Vec *);
PetscErrorCode VecGetDM(Vec, DM *);
PetscErrorCode ISDestroy(IS *);
PetscErrorCode VecDestroy(Vec *);
PetscErrorCode VecDuplicate(Vec, Vec *);
PetscErrorCode VecSet(Vec, PetscScalar);
PetscErrorCode VecZeroEntries(Vec);
PetscErrorCode VecCopy(Vec, Vec);
PetscErrorCode VecScale(Vec, PetscScalar);
PetscErrorCode VecAXPY(Vec y, PetscScalar a, Vec x);                                   /* y += a x */
PetscErrorCode VecAYPX(Vec y, PetscScalar a, Vec x);                                   /* y = x + a y */
PetscErrorCode VecWAXPY(Vec w, PetscScalar a, Vec x, Vec y);                           /* w = a x + y */
PetscErrorCode VecAXPBYPCZ(Vec z, PetscScalar a, PetscScalar b, PetscScalar c, Vec x, Vec y); /* z = a x + b y + c z */
PetscErrorCode VecReciprocal(Vec);
PetscErrorCode VecPointwiseMult(Vec w, Vec x, Vec y);
PetscErrorCode VecGetSize(Vec, PetscInt *);
PetscErrorCode VecGetSubVector(Vec, IS, Vec *);
PetscErrorCode VecRestoreSubVector(Vec, IS, Vec *);
PetscErrorCode VecAssemblyBegin(Vec);
PetscErrorCode VecAssemblyEnd(Vec);
PetscErrorCode VecView(Vec, PetscViewer);

/* ---- Mat ---- */
PetscErrorCode MatCreate(MPI_Comm, Mat *);
PetscErrorCode MatSetSizes(Mat, PetscInt, PetscInt, PetscInt, PetscInt);
PetscErrorCode MatSetType(Mat, MatType);
PetscErrorCode MatSetUp(Mat);
PetscErrorCode MatSetLocalToGlobalMapping(Mat, ISLocalToGlobalMapping, ISLocalToGlobalMapping);
PetscErrorCode MatSetOption(Mat, MatOption, PetscBool);
PetscErrorCode MatSetValuesLocal(Mat, PetscInt, const PetscInt[], PetscInt, const PetscInt[], const PetscScalar[], InsertMode);
PetscErrorCode MatAssemblyBegin(Mat, MatAssemblyType);
PetscErrorCode MatAssemblyEnd(Mat, MatAssemblyType);
PetscErrorCode MatDestroy(Mat *);
PetscErrorCode MatZeroEntries(Mat);
PetscErrorCode MatScale(Mat, PetscScalar);
PetscErrorCode MatShift(Mat, PetscScalar);
PetscErrorCode MatAXPY(Mat Y, PetscScalar a, Mat X, MatStructure);
PetscErrorCode MatMult(Mat, Vec x, Vec y);
PetscErrorCode MatMultAdd(Mat, Vec x, Vec y, Vec z); /* z = y + A x */
PetscErrorCode MatMatMult(Mat A, Mat B, MatReuse, PetscReal fill, Mat *C);
PetscErrorCode MatDuplicate(Mat, MatDuplicateOption, Mat *);
PetscErrorCode MatDiagonalScale(Mat, Vec l, Vec r);
PetscErrorCode MatGetDiagonal(Mat, Vec);
PetscErrorCode MatGetRowSum(Mat, Vec);
PetscErrorCode MatCreateVecs(Mat, Vec *right, Vec *left);
PetscErrorCode MatCreateConstantDiagonal(MPI_Comm, PetscInt, PetscInt, PetscInt, PetscInt, PetscScalar, Mat *);
PetscErrorCode MatCreateSubMatrix(Mat, IS, IS, MatReuse, Mat *);
PetscErrorCode MatCreateNest(MPI_Comm, PetscInt, const IS[], PetscInt, const IS[], const Mat[], Mat *);
PetscErrorCode MatNestSetSubMat(Mat, PetscInt, PetscInt, Mat);
PetscErrorCode MatNestGetSubMat(Mat, PetscInt, PetscInt, Mat *);
PetscErrorCode MatNestGetSize(Mat, PetscInt *, PetscInt *);
PetscErrorCode MatNestGetISs(Mat, IS[], IS[]);
PetscErrorCode MatNestSetVecType(Mat, VecType);
PetscErrorCode MatNullSpaceCreate(MPI_Comm, PetscBool has_cnst, PetscInt n, const Vec[], MatNullSpace *);
PetscErrorCode MatNullSpaceDestroy(MatNullSpace *);
PetscErrorCode MatNullSpaceRemove(MatNullSpace, Vec);
PetscErrorCode MatSetNullSpace(Mat, MatNullSpace);
PetscErrorCode MatGetNullSpace(Mat, MatNullSpace *);

/* ---- KSP (exact solves) ---- */
PetscErrorCode KSPCreate(MPI_Comm, KSP *);
PetscErrorCode KSPSetOperators(KSP, Mat, Mat);
PetscErrorCode KSPSolve(KSP, Vec b, Vec x);
PetscErrorCode KSPDestroy(KSP *);
PetscErrorCode KSPSetFromOptions(KSP);
PetscErrorCode KSPSetOptionsPrefix(KSP, const char[]);
PetscErrorCode KSPView(KSP, PetscViewer);
PetscErrorCode SNESSolve(SNES, Vec b, Vec x);
typedef enum { KSP_NORM_DEFAULT = -1, KSP_NORM_NONE, KSP_NORM_PRECONDITIONED, KSP_NORM_UNPRECONDITIONED, KSP_NORM_NATURAL } KSPNormType;
typedef enum { SNES_CONVERGED_ITERATING = 0, SNES_CONVERGED_FNORM_RELATIVE = 3, SNES_DIVERGED_LINEAR_SOLVE = -3 } SNESConvergedReason;
extern const char *const *SNESConvergedReasons;
typedef const char       *PCType;
PetscErrorCode KSPGetPC(KSP, PC *);
PetscErrorCode KSPSetTolerances(KSP, PetscReal, PetscReal, PetscReal, PetscInt);
PetscErrorCode KSPSetNormType(KSP, KSPNormType);
PetscErrorCode PCSetType(PC, PCType);
PetscErrorCode PCRegister(const char[], PetscErrorCode (*)(PC));
PetscErrorCode SNESCreate(MPI_Comm, SNES *);
PetscErrorCode SNESDestroy(SNES *);
PetscErrorCode SNESGetKSP(SNES, KSP *);
PetscErrorCode SNESSetTolerances(SNES, PetscReal, PetscReal, PetscReal, PetscInt, PetscInt);
PetscErrorCode SNESSetOptionsPrefix(SNES, const char[]);
PetscErrorCode SNESAppendOptionsPrefix(SNES, const char[]);
PetscErrorCode SNESSetFromOptions(SNES);
PetscErrorCode SNESSetPicard(SNES, Vec, PetscErrorCode (*)(SNES, Vec, Vec, void *), Mat, Mat, PetscErrorCode (*)(SNES, Vec, Mat, Mat, void *), void *);
PetscErrorCode SNESSetFunction(SNES, Vec, PetscErrorCode (*)(SNES, Vec, Vec, void *), void *);
PetscErrorCode SNESSetComputeInitialGuess(SNES, PetscErrorCode (*)(SNES, Vec, void *), void *);
PetscErrorCode SNESPicardComputeFunction(SNES, Vec, Vec, void *);
PetscErrorCode SNESMonitorCancel(SNES);
PetscErrorCode SNESGetConvergedReason(SNES, SNESConvergedReason *);

/* ---- DM / DMStag ---- */
typedef enum { DMSTAG_STENCIL_NONE, DMSTAG_STENCIL_STAR, DMSTAG_STENCIL_BOX } DMStagStencilType;
typedef const char *DMType;
#define DMPRODUCT "product"
#define DMSTAG "stag"
PetscErrorCode DMStagCreate2d(MPI_Comm, DMBoundaryType, DMBoundaryType, PetscInt, PetscInt, PetscInt, PetscInt, PetscInt, PetscInt, PetscInt, DMStagStencilType, PetscInt, const PetscInt[], const PetscInt[], DM *);
PetscErrorCode DMStagCreate3d(MPI_Comm, DMBoundaryType, DMBoundaryType, DMBoundaryType, PetscInt, PetscInt, PetscInt, PetscInt, PetscInt, PetscInt, PetscInt, PetscInt, PetscInt, PetscInt, DMStagStencilType, PetscInt, const PetscInt[], const PetscInt[], const PetscInt[], DM *);
PetscErrorCode DMStagCreateCompatibleDMStag(DM, PetscInt, PetscInt, PetscInt, PetscInt, DM *);
PetscErrorCode DMSetUp(DM);
PetscErrorCode DMDestroy(DM *);
PetscErrorCode DMSetMatrixPreallocateOnly(DM, PetscBool);
PetscErrorCode DMStagSetRefinementFactor(DM, PetscInt, PetscInt, PetscInt);
PetscErrorCode DMStagGetNumRanks(DM, PetscInt *, PetscInt *, PetscInt *);
PetscErrorCode DMStagGetOwnershipRanges(DM, const PetscInt *[], const PetscInt *[], const PetscInt *[]);
PetscErrorCode DMStagGetLocalSizes(DM, PetscInt *, PetscInt *, PetscInt *);
PetscErrorCode DMStagSetUniformCoordinatesProduct(DM, PetscReal, PetscReal, PetscReal, PetscReal, PetscReal, PetscReal);
PetscErrorCode DMStagGetProductCoordinateArrays(DM, void *, void *, void *);
PetscErrorCode DMStagRestoreProductCoordinateArrays(DM, void *, void *, void *);
PetscErrorCode DMStagSetCoordinateDMType(DM, DMType);
PetscErrorCode DMGetCoordinateDM(DM, DM *);
PetscErrorCode DMSetCoordinateDM(DM, DM);
PetscErrorCode DMCompositeCreate(MPI_Comm, DM *);
PetscErrorCode DMCompositeAddDM(DM, DM);
PetscErrorCode DMCompositeGetGlobalISs(DM, IS *[]);
PetscErrorCode DMGetDimension(DM, PetscInt *);
PetscErrorCode DMGetLocalVector(DM, Vec *);
PetscErrorCode DMRestoreLocalVector(DM, Vec *);
PetscErrorCode DMGetGlobalVector(DM, Vec *);
PetscErrorCode DMRestoreGlobalVector(DM, Vec *);
PetscErrorCode DMCreateGlobalVector(DM, Vec *);
PetscErrorCode DMGlobalToLocal(DM, Vec, InsertMode, Vec);
PetscErrorCode DMLocalToGlobal(DM, Vec, InsertMode, Vec);
PetscErrorCode DMGetLocalToGlobalMapping(DM, ISLocalToGlobalMapping *);
PetscErrorCode DMGetMatType(DM, MatType *);
PetscErrorCode DMStagGetGlobalSizes(DM, PetscInt *, PetscInt *, PetscInt *);
PetscErrorCode DMStagGetCorners(DM, PetscInt *, PetscInt *, PetscInt *, PetscInt *, PetscInt *, PetscInt *, PetscInt *, PetscInt *, PetscInt *);
PetscErrorCode DMStagGetIsFirstRank(DM, PetscBool *, PetscBool *, PetscBool *);
PetscErrorCode DMStagGetIsLastRank(DM, PetscBool *, PetscBool *, PetscBool *);
PetscErrorCode DMStagGetEntries(DM, PetscInt *);
PetscErrorCode DMStagGetLocationSlot(DM, DMStagStencilLocation, PetscInt, PetscInt *);
PetscErrorCode DMStagStencilToIndexLocal(DM, PetscInt dim, PetscInt n, const DMStagStencil *, PetscInt *);
PetscErrorCode DMStagVecSetValuesStencil(DM, Vec, PetscInt, const DMStagStencil *, const PetscScalar *, InsertMode);
PetscErrorCode DMStagMatSetValuesStencil(DM, Mat, PetscInt, const DMStagStencil *, PetscInt, const DMStagStencil *, const PetscScalar *, InsertMode);
PetscErrorCode DMStagVecGetArray(DM, Vec, void *);
PetscErrorCode DMStagVecRestoreArray(DM, Vec, void *);
PetscErrorCode DMStagVecGetArrayRead(DM, Vec, void *);
PetscErrorCode DMStagVecRestoreArrayRead(DM, Vec, void *);
PetscErrorCode DMStagGetProductCoordinateArraysRead(DM, void *, void *, void *);
PetscErrorCode DMStagRestoreProductCoordinateArraysRead(DM, void *, void *, void *);
PetscErrorCode DMStagGetProductCoordinateLocationSlot(DM, DMStagStencilLocation, PetscInt *);
#endif
