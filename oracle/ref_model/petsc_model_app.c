/* oracle/ref_model/petsc_model_app.c -- TEST INFRASTRUCTURE ONLY: the part of the single-rank PETSc model that a whole PROGRAM of the
 * reference needs (its sys / mesh / ns-interface sources and its own test drivers, oracle/Makefile target `ref_app`): program start
 * and the options database (with PETSc's -dll_append plugin loading), class and function-list registration, PETSCHEADER objects,
 * DMStag creation with product coordinates, DMComposite / VecNest construction, the SNES -> KSP -> PC object chain with a generic
 * Picard solve, and viewers ("ascii" and a stand-in for the reference's CGNS viewer that dumps named vectors to a file).
 * Written from the PETSc manual pages; no PETSc source.  See include/petsc_model.h for the semantics that are and are not modelled. */
#include "petsc_model_impl.h"
#include <dlfcn.h>
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>

/* ------------------------------------------------------------------ options database */
#define MAXOPT 128
static struct {
  char *name, *value;
  int   used;
} opts[MAXOPT];
static int nopts = 0;

PetscErrorCode ModelOptionsSetValue(const char name[], const char value[])
{
  int i;
  for (i = 0; i < nopts; ++i)
    if (!strcmp(opts[i].name, name)) break;
  PetscCheck(i < MAXOPT, 0, PETSC_ERR_LIB, "too many options");
  if (i == nopts) opts[nopts++].name = strdup(name);
  else free(opts[i].value);
  opts[i].value = strdup(value ? value : ""), opts[i].used = 0;
  return PETSC_SUCCESS;
}
PetscErrorCode ModelOptionsClear(void)
{
  int i;
  for (i = 0; i < nopts; ++i) free(opts[i].name), free(opts[i].value);
  nopts = 0;
  return PETSC_SUCCESS;
}
/* PETSc prepends the options prefix of the object being configured to the names asked for inside its options block
   ("-pc_abf_schur_ainv_type" of the PC with prefix "ns_" is given as -ns_pc_abf_schur_ainv_type) */
static const char *current_prefix = NULL;
void ModelOptionsPrefixPush(const char *prefix) { current_prefix = prefix; }
void ModelOptionsPrefixPop(void) { current_prefix = NULL; }
const char *ModelOptionsPrefixGet(void) { return current_prefix; }
static const char *opt_find(const char *name)
{
  char full[600];
  int  i;
  if (current_prefix && current_prefix[0] && name[0] == '-') snprintf(full, sizeof(full), "-%s%s", current_prefix, name + 1);
  else snprintf(full, sizeof(full), "%s", name);
  for (i = 0; i < nopts; ++i)
    if (!strcmp(opts[i].name, full)) return opts[i].used = 1, opts[i].value;
  return NULL;
}
PetscErrorCode ModelOptionsReal(const char name[], PetscReal *val, PetscBool *set)
{
  const char *v = opt_find(name);
  if (v) *val = atof(v);
  if (set) *set = v ? PETSC_TRUE : PETSC_FALSE;
  return PETSC_SUCCESS;
}
PetscErrorCode ModelOptionsInt(const char name[], PetscInt *val, PetscBool *set)
{
  const char *v = opt_find(name);
  if (v) *val = (PetscInt)atol(v);
  if (set) *set = v ? PETSC_TRUE : PETSC_FALSE;
  return PETSC_SUCCESS;
}
PetscErrorCode ModelOptionsBool(const char name[], PetscBool *val, PetscBool *set)
{
  const char *v = opt_find(name);
  if (v) *val = (!v[0] || !strcmp(v, "1") || !strcmp(v, "true") || !strcmp(v, "yes")) ? PETSC_TRUE : PETSC_FALSE; /* a bare flag is true */
  if (set) *set = v ? PETSC_TRUE : PETSC_FALSE;
  return PETSC_SUCCESS;
}
PetscErrorCode ModelOptionsString(const char name[], char *val, size_t len, PetscBool *set)
{
  const char *v = opt_find(name);
  if (v) snprintf(val, len, "%s", v);
  if (set) *set = v ? PETSC_TRUE : PETSC_FALSE;
  return PETSC_SUCCESS;
}
static int same_nocase(const char *a, const char *b)
{
  for (; *a && *b; ++a, ++b)
    if ((*a | 32) != (*b | 32)) return 0;
  return !*a && !*b;
}
PetscErrorCode ModelOptionsEnum(const char name[], const char *const *list, PetscEnum *val, PetscBool *set)
{
  const char *v = opt_find(name);
  int         n = 0, i;
  while (list[n]) ++n;
  n -= 2; /* value names, then the enum's type name and prefix */
  if (v) {
    for (i = 0; i < n; ++i)
      if (same_nocase(v, list[i])) break;
    PetscCheck(i < n, 0, PETSC_ERR_ARG_WRONG, "Unknown value %s for option %s", v, name);
    *val = (PetscEnum)i;
  }
  if (set) *set = v ? PETSC_TRUE : PETSC_FALSE;
  return PETSC_SUCCESS;
}
PetscErrorCode PetscStrInList(const char s[], const char list[], char sep, PetscBool *found)
{
  const char *p = list;
  size_t      n = strlen(s);
  *found = PETSC_FALSE;
  while (p && *p) {
    const char *e = strchr(p, sep);
    size_t      l = e ? (size_t)(e - p) : strlen(p);
    if (l == n && !strncmp(p, s, n)) *found = PETSC_TRUE;
    p = e ? e + 1 : NULL;
  }
  return PETSC_SUCCESS;
}

/* ------------------------------------------------------------------ program start / end */
static int initialized = 0, finalized = 0, nfinalizers = 0;
static PetscErrorCode (*finalizers[32])(void);
static void *plugins[8];
static int   nplugins = 0;

PetscErrorCode PetscInitialize(int *argc, char ***argv, const char file[], const char help[])
{
  int a;
  (void)file, (void)help;
  for (a = 1; argc && a < *argc; ++a) { /* "-name value" or a bare "-flag" */
    const char *arg = (*argv)[a];
    if (arg[0] != '-' || (arg[1] >= '0' && arg[1] <= '9')) continue;
    if (a + 1 < *argc && ((*argv)[a + 1][0] != '-' || ((*argv)[a + 1][1] >= '0' && (*argv)[a + 1][1] <= '9') || (*argv)[a + 1][1] == '.')) PetscCall(ModelOptionsSetValue(arg, (*argv)[++a]));
    else PetscCall(ModelOptionsSetValue(arg, ""));
  }
  initialized = 1, finalized = 0;
  { /* -model_solvers iterative: every KSP nobody configures runs GMRES(30) (+ ILU(0) for the assembled inner ones), as serial PETSc does */
    const char *ms = opt_find("-model_solvers");
    if (ms) {
      PetscCheck(!strcmp(ms, "iterative") || !strcmp(ms, "exact"), 0, PETSC_ERR_SUP, "-model_solvers exact | iterative, not %s", ms);
      ModelKSPSetDefaults(!strcmp(ms, "iterative"));
    }
  }
  /* -dll_append <library>: PETSc opens the shared library and calls PetscDLLibraryRegister_<name>, <name> = the file name without
     directory, "lib" and suffix.  That is how a type implemented outside the reference (glue/nsb200.c) registers itself. */
  {
    const char *lib = opt_find("-dll_append");
    if (lib && lib[0]) {
      char        sym[512], base[256];
      const char *b = strrchr(lib, '/');
      char       *dot;
      void       *h = dlopen(lib, RTLD_NOW | RTLD_GLOBAL);
      PetscErrorCode (*reg)(void);
      PetscCheck(h, 0, PETSC_ERR_LIB, "-dll_append: cannot open %s: %s", lib, dlerror());
      snprintf(base, sizeof(base), "%s", b ? b + 1 : lib);
      if ((dot = strchr(base, '.'))) *dot = 0;
      snprintf(sym, sizeof(sym), "PetscDLLibraryRegister_%s", !strncmp(base, "lib", 3) ? base + 3 : base);
      reg = (PetscErrorCode(*)(void))dlsym(h, sym);
      PetscCheck(reg, 0, PETSC_ERR_LIB, "-dll_append: %s has no %s", lib, sym);
      PetscCheck(nplugins < 8, 0, PETSC_ERR_LIB, "too many plugins");
      plugins[nplugins++] = h;
      PetscCall(reg());
    }
  }
  return PETSC_SUCCESS;
}
PetscErrorCode PetscInitialized(PetscBool *b) { return *b = initialized ? PETSC_TRUE : PETSC_FALSE, PETSC_SUCCESS; }
PetscErrorCode PetscFinalized(PetscBool *b) { return *b = finalized ? PETSC_TRUE : PETSC_FALSE, PETSC_SUCCESS; }
PetscErrorCode PetscRegisterFinalize(PetscErrorCode (*f)(void))
{
  PetscCheck(nfinalizers < 32, 0, PETSC_ERR_LIB, "too many finalizers");
  finalizers[nfinalizers++] = f;
  return PETSC_SUCCESS;
}
static void events_report(FILE *f);
PetscErrorCode PetscFinalize(void)
{
  int i;
  while (nfinalizers > 0) PetscCall(finalizers[--nfinalizers]());
  for (i = 0; i < nopts; ++i)
    if (!opts[i].used && strcmp(opts[i].name, "-dll_append")) fprintf(stderr, "WARNING! There are options you set that were not used! Option left: name:%s value: %s\n", opts[i].name, opts[i].value);
  ModelOptionsClear();
  if (getenv("PETSC_MODEL_TIMING")) ModelTimingReport(stderr), events_report(stderr);
  initialized = 0, finalized = 1;
  return PETSC_SUCCESS;
}
static PetscClassId next_classid = 1000;
PetscErrorCode PetscClassIdRegister(const char name[], PetscClassId *id) { return (void)name, *id = next_classid++, PETSC_SUCCESS; }
enum { MAX_EVENTS = 64 };
static struct {
  int    id;
  char   name[48];
  double t0, sum;
  long   calls;
} events[MAX_EVENTS];
static int nevents = 0, events_on = -1;
PetscErrorCode PetscLogEventRegister(const char name[], PetscClassId c, PetscLogEvent *e)
{
  (void)c;
  *e = next_classid++;
  if (nevents < MAX_EVENTS) events[nevents].id = *e, snprintf(events[nevents].name, sizeof(events[nevents].name), "%s", name), ++nevents;
  return PETSC_SUCCESS;
}
static int event_slot(PetscLogEvent e)
{
  int i;
  if (events_on < 0) events_on = getenv("PETSC_MODEL_TIMING") ? 1 : 0;
  if (!events_on) return -1;
  for (i = 0; i < nevents; ++i)
    if (events[i].id == e) return i;
  return -1;
}
PetscErrorCode ModelLogEventBegin(PetscLogEvent e)
{
  const int i = event_slot(e);
  if (i >= 0) events[i].t0 = ModelWallTime();
  return PETSC_SUCCESS;
}
PetscErrorCode ModelLogEventEnd(PetscLogEvent e)
{
  const int i = event_slot(e);
  if (i >= 0) events[i].sum += ModelWallTime() - events[i].t0, ++events[i].calls;
  return PETSC_SUCCESS;
}
static void events_report(FILE *f)
{
  int i;
  for (i = 0; i < nevents; ++i)
    if (events[i].calls) fprintf(f, "[PETSc model timing] event %-22s %8ld calls %10.3f s\n", events[i].name, events[i].calls, events[i].sum);
}
PetscErrorCode PetscInfoProcessClass(const char n[], PetscInt k, PetscClassId ids[]) { return (void)n, (void)k, (void)ids, PETSC_SUCCESS; }
PetscErrorCode PetscLogEventExcludeClass(PetscClassId c) { return (void)c, PETSC_SUCCESS; }

/* ------------------------------------------------------------------ function lists, headers */
struct _n_PetscFunctionList {
  char *name;
  void (*f)(void);
  struct _n_PetscFunctionList *next;
};
PetscErrorCode ModelFunctionListAdd(PetscFunctionList *list, const char name[], void (*f)(void))
{
  PetscFunctionList e;
  for (e = *list; e; e = e->next)
    if (!strcmp(e->name, name)) return e->f = f, PETSC_SUCCESS;
  e       = (PetscFunctionList)calloc(1, sizeof(*e));
  e->name = strdup(name), e->f = f, e->next = *list, *list = e;
  return PETSC_SUCCESS;
}
PetscErrorCode ModelFunctionListFind(PetscFunctionList list, const char name[], void (**f)(void))
{
  *f = NULL;
  for (; list; list = list->next)
    if (!strcmp(list->name, name)) *f = list->f;
  return PETSC_SUCCESS;
}
PetscErrorCode PetscFunctionListDestroy(PetscFunctionList *list)
{
  while (*list) {
    PetscFunctionList e = *list;
    *list               = e->next;
    free(e->name), free(e);
  }
  return PETSC_SUCCESS;
}
PetscErrorCode ModelHeaderCreate(void *pobj, size_t size, PetscClassId classid, const char cls[], MPI_Comm comm, PetscErrorCode (*destroy)(struct _p_PetscObject **))
{
  void *o = calloc(1, size);
  PetscCheck(o, 0, 55, "out of memory");
  ModelHeaderInit(o, classid, cls, NULL, NULL);
  ((PetscObject)o)->comm = comm, ((PetscObject)o)->destroy_public = destroy;
  *(void **)pobj         = o;
  return PETSC_SUCCESS;
}
PetscErrorCode ModelHeaderDestroy(void *pobj)
{
  void *o = *(void **)pobj;
  if (o) ModelHeaderFree(o), free(o);
  *(void **)pobj = NULL;
  return PETSC_SUCCESS;
}
PetscErrorCode PetscObjectChangeTypeName(PetscObject o, const char t[])
{
  free(o->type_name);
  o->type_name = t ? strdup(t) : NULL;
  return PETSC_SUCCESS;
}
PetscErrorCode PetscObjectGetName(PetscObject o, const char *n[]) { return *n = o->name ? o->name : "", PETSC_SUCCESS; }
PetscErrorCode PetscObjectPrintClassNamePrefixType(PetscObject o, PetscViewer v) { return PetscViewerASCIIPrintf(v, "%s Object: %s\n  type: %s\n", o->class_name, o->prefix ? o->prefix : "", o->type_name ? o->type_name : "not yet set"); }
PetscErrorCode PetscObjectDereference(PetscObject o)
{
  if (o && o->destroy_model) return o->destroy_model(o);
  if (o && o->destroy_public) return o->destroy_public(&o);
  return PETSC_SUCCESS;
}

/* ------------------------------------------------------------------ viewers */
static PetscViewer viewer_stdout = NULL;
PetscErrorCode PetscViewerASCIIGetStdout(MPI_Comm c, PetscViewer *v)
{
  (void)c;
  if (!viewer_stdout) {
    viewer_stdout = (PetscViewer)calloc(1, sizeof(*viewer_stdout));
    ModelHeaderInit(viewer_stdout, 31, "PetscViewer", PETSCVIEWERASCII, NULL);
    viewer_stdout->f = stdout, viewer_stdout->hdr.refct = 1 << 20; /* never destroyed */
  }
  *v = viewer_stdout;
  return PETSC_SUCCESS;
}
/* what the reference's FlucaOptionsCreateViewer (viewer package, needs CGNS) does for "-name type[:file]": here "ascii" and
   "flucacgns:<file>" -- the latter a stand-in that dumps every vector viewed into it to <file> (see VecView_Cart_Local_CGNS) */
PetscErrorCode FlucaOptionsCreateViewer(MPI_Comm comm, PetscOptions o, const char pre[], const char name[], PetscViewer *viewer, PetscViewerFormat *format, PetscBool *set)
{
  const char *v = opt_find(name);
  (void)o, (void)pre;
  if (set) *set = v ? PETSC_TRUE : PETSC_FALSE;
  if (format) *format = 0;
  if (!v) return PETSC_SUCCESS;
  if (!v[0] || !strcmp(v, "ascii")) {
    PetscCall(PetscViewerASCIIGetStdout(comm, viewer));
    ++(*viewer)->hdr.refct;
    return PETSC_SUCCESS;
  }
  PetscCheck(!strncmp(v, "flucacgns:", 10) || !strncmp(v, "cgns:", 5), comm, PETSC_ERR_SUP, "the model opens ascii and flucacgns:<file> viewers, not %s", v);
  *viewer = (PetscViewer)calloc(1, sizeof(**viewer));
  ModelHeaderInit(*viewer, 31, "PetscViewer", "flucacgns", NULL);
  (*viewer)->path = strdup(strchr(v, ':') + 1);
  return PETSC_SUCCESS;
}
PetscErrorCode PetscViewerDestroy(PetscViewer *pv)
{
  PetscViewer v = *pv;
  if (!v) return PETSC_SUCCESS;
  *pv = NULL;
  if (--v->hdr.refct > 0) return PETSC_SUCCESS;
  while (v->store) {
    struct stored_vec *s = v->store;
    v->store             = s->next;
    free(s->name), free(s->a), free(s);
  }
  ModelHeaderFree(v);
  free(v->path), free(v);
  return PETSC_SUCCESS;
}
PetscErrorCode PetscViewerAndFormatCreate(PetscViewer v, PetscViewerFormat f, PetscViewerAndFormat **vf)
{
  *vf           = (PetscViewerAndFormat *)calloc(1, sizeof(**vf));
  (*vf)->viewer = v, (*vf)->format = f, (*vf)->view_interval = 1;
  if (v) ++v->hdr.refct;
  return PETSC_SUCCESS;
}
PetscErrorCode PetscViewerAndFormatDestroy(PetscViewerAndFormat **vf)
{
  if (*vf) {
    PetscCall(PetscViewerDestroy(&(*vf)->viewer));
    free(*vf);
  }
  *vf = NULL;
  return PETSC_SUCCESS;
}
PetscErrorCode PetscViewerPushFormat(PetscViewer v, PetscViewerFormat f) { return (void)v, (void)f, PETSC_SUCCESS; }
PetscErrorCode PetscViewerPopFormat(PetscViewer v) { return (void)v, PETSC_SUCCESS; }
PetscErrorCode PetscViewerFlush(PetscViewer v)
{
  if (v && v->f) fflush(v->f);
  return PETSC_SUCCESS;
}
PetscErrorCode PetscViewerCheckReadable(PetscViewer v) { return (void)v, PETSC_SUCCESS; }
PetscErrorCode PetscViewerRegister(const char n[], PetscErrorCode (*f)(PetscViewer)) { return (void)n, (void)f, PETSC_SUCCESS; }
PetscErrorCode PetscViewerASCIISynchronizedPrintf(PetscViewer v, const char fmt[], ...)
{
  va_list ap;
  if (!v || !v->f) return PETSC_SUCCESS;
  va_start(ap, fmt);
  vfprintf(v->f, fmt, ap);
  va_end(ap);
  return PETSC_SUCCESS;
}
PetscErrorCode PetscViewerASCIIAddTab(PetscViewer v, PetscInt n) { return (void)v, (void)n, PETSC_SUCCESS; }
PetscErrorCode PetscViewerASCIISubtractTab(PetscViewer v, PetscInt n) { return (void)v, (void)n, PETSC_SUCCESS; }
PetscErrorCode PetscViewerASCIIPushSynchronized(PetscViewer v) { return (void)v, PETSC_SUCCESS; }
PetscErrorCode PetscViewerASCIIPopSynchronized(PetscViewer v) { return (void)v, PETSC_SUCCESS; }
PetscErrorCode PetscMonitorCompare(PetscErrorCode (*nmon)(void), void *nmctx, PetscErrorCode (*nmdestroy)(void **), PetscErrorCode (*mon)(void), void *mctx, PetscErrorCode (*mdestroy)(void **), PetscBool *identical)
{
  *identical = (nmon == mon && nmctx == mctx && nmdestroy == mdestroy) ? PETSC_TRUE : PETSC_FALSE;
  return PETSC_SUCCESS;
}

/* the CGNS back ends of the reference's mesh and vector viewers (cartcgns.c needs the CGNS library): the stand-in writes every
   vector viewed into a "flucacgns" viewer to the viewer's file as  name\n n\n  followed by n doubles in the GLOBAL ordering of the
   vector's DM, and keeps a copy in memory so that a later load finds it (restart) */
PetscErrorCode MeshView_Cart_CGNS(void *mesh, PetscViewer v) { return (void)mesh, (void)v, PETSC_SUCCESS; }
PetscErrorCode MeshLoad_Cart_CGNS(void *mesh, PetscViewer v) { return (void)mesh, (void)v, PETSC_ERR_SUP; }
/* the reference's public constructor of its CGNS viewer (flucaviewer.h): here the stand-in; FILE_MODE_READ loads the dump */
PetscErrorCode MeshSetOutputSequenceNumber(void *mesh, PetscInt num, PetscReal val); /* meshbasic.c:198-212 */
PetscErrorCode MeshGetOutputSequenceNumber(void *mesh, PetscInt *num, PetscReal *val);
PetscErrorCode PetscViewerFlucaCGNSOpen(MPI_Comm comm, const char path[], PetscFileMode mode, PetscViewer *viewer)
{
  (void)comm;
  *viewer = (PetscViewer)calloc(1, sizeof(**viewer));
  ModelHeaderInit(*viewer, 31, "PetscViewer", "flucacgns", NULL);
  (*viewer)->path = strdup(path);
  (*viewer)->step = -1;
  if (mode == FILE_MODE_READ) {
    FILE *f = fopen(path, "rb");
    char  name[256];
    int   n;
    PetscCheck(f, 0, PETSC_ERR_LIB, "cannot read %s", path);
    free((*viewer)->path), (*viewer)->path = NULL; /* a reader never rewrites its file */
    while (fgets(name, sizeof(name), f)) {
      struct stored_vec *s = (struct stored_vec *)calloc(1, sizeof(*s));
      name[strcspn(name, "\n")] = 0;
      if (!strcmp(name, "@sequence")) { /* the output sequence number the writer's mesh carried (cartcgns.c:331, :714-724) */
        double t;
        free(s);
        char line[256];
        PetscCheck(fgets(line, sizeof(line), f) && sscanf(line, "%d %lf", &n, &t) == 2, 0, PETSC_ERR_LIB, "bad sequence record in %s", path);
        (*viewer)->step = n, (*viewer)->time = t;
        continue;
      }
      {
        char line[256]; /* not fscanf("%d\n"): its trailing white space would eat payload bytes that happen to be blanks */
        PetscCheck(fgets(line, sizeof(line), f) && sscanf(line, "%d", &n) == 1, 0, PETSC_ERR_LIB, "bad record in %s", path);
      }
      s->name = strdup(name), s->n = n, s->a = (double *)calloc((size_t)n, sizeof(double));
      PetscCheck(fread(s->a, sizeof(double), (size_t)n, f) == (size_t)n, 0, PETSC_ERR_LIB, "short record in %s", path);
      (void)fgetc(f);
      s->next = (*viewer)->store, (*viewer)->store = s;
    }
    fclose(f);
  }
  return PETSC_SUCCESS;
}
PetscErrorCode VecView_Cart_Local_CGNS(Vec v, PetscViewer w)
{
  struct stored_vec *s;
  const char        *name = v->hdr.name ? v->hdr.name : "";
  PetscObject        mesh = NULL;
  PetscCall(PetscObjectQuery((PetscObject)v, "Fluca_Mesh", &mesh)); /* cart.c:225: every mesh vector knows its mesh */
  if (mesh) PetscCall(MeshGetOutputSequenceNumber(mesh, &w->step, &w->time));
  for (s = w->store; s; s = s->next)
    if (!strcmp(s->name, name)) break;
  if (!s) {
    s       = (struct stored_vec *)calloc(1, sizeof(*s));
    s->name = strdup(name), s->n = v->n, s->a = (double *)calloc((size_t)v->n, sizeof(double));
    s->next = w->store, w->store = s;
  }
  PetscCheck(s->n == v->n, 0, PETSC_ERR_ARG_WRONG, "size of %s changed", name);
  memcpy(s->a, v->a, sizeof(double) * (size_t)v->n);
  if (w->path) { /* rewrite the whole file: the latest copy of every vector */
    FILE *f = fopen(w->path, "wb");
    PetscCheck(f, 0, PETSC_ERR_LIB, "cannot write %s", w->path);
    fprintf(f, "@sequence\n%d %.17g\n", (int)w->step, w->time);
    for (s = w->store; s; s = s->next) {
      fprintf(f, "%s\n%d\n", s->name, (int)s->n);
      fwrite(s->a, sizeof(double), (size_t)s->n, f);
      fputc('\n', f);
    }
    fclose(f);
  }
  return PETSC_SUCCESS;
}
PetscErrorCode VecLoad_Cart_CGNS(Vec v, PetscViewer w)
{
  struct stored_vec *s;
  for (s = w->store; s; s = s->next)
    if (v->hdr.name && !strcmp(s->name, v->hdr.name)) break;
  PetscCheck(s && s->n == v->n, 0, PETSC_ERR_ARG_WRONG, "no stored vector named %s", v->hdr.name ? v->hdr.name : "(unnamed)");
  memcpy(v->a, s->a, sizeof(double) * (size_t)v->n);
  ++v->hdr.state;
  { /* cartcgns.c:714-724: the first vector loaded sets the mesh's output sequence number from the file */
    PetscObject mesh = NULL;
    PetscInt    step;
    PetscReal   time;
    PetscCall(PetscObjectQuery((PetscObject)v, "Fluca_Mesh", &mesh));
    if (mesh) {
      PetscCall(MeshGetOutputSequenceNumber(mesh, &step, &time));
      if (step == -1 && time == 0.) PetscCall(MeshSetOutputSequenceNumber(mesh, w->step, w->time));
    }
  }
  return PETSC_SUCCESS;
}
PetscErrorCode FlucaVecLoad(Vec v, PetscViewer w) { return v->load_op ? v->load_op(v, w) : VecLoad_Cart_CGNS(v, w); }
const char *cg_get_error(void) { return "the CGNS library is not part of the model"; }

/* ------------------------------------------------------------------ Vec / IS additions */
PetscErrorCode VecNorm(Vec v, NormType t, PetscReal *r)
{
  double *a = (double *)malloc(sizeof(double) * (size_t)(v->n ? v->n : 1)), s = 0.;
  int     i;
  ModelVecGather(v, a);
  for (i = 0; i < v->n; ++i) s = t == NORM_2 ? s + a[i] * a[i] : (t == NORM_1 ? s + fabs(a[i]) : fmax(s, fabs(a[i])));
  *r = t == NORM_2 ? sqrt(s) : s;
  free(a);
  return PETSC_SUCCESS;
}
PetscErrorCode VecSetOperation(Vec v, VecOperation op, void (*f)(void))
{
  if (op == VECOP_VIEW) v->view_op = (PetscErrorCode(*)(Vec, PetscViewer))f;
  else if (op == VECOP_LOAD) v->load_op = (PetscErrorCode(*)(Vec, PetscViewer))f;
  return PETSC_SUCCESS;
}
PetscErrorCode VecGetDM(Vec v, DM *dm) { return *dm = v->dm, PETSC_SUCCESS; }
PetscErrorCode VecCreateNest(MPI_Comm c, PetscInt n, IS is[], Vec sub[], Vec *nest)
{
  (void)c, (void)is;
  PetscCheck(n == 3, 0, PETSC_ERR_SUP, "the model nests three fields");
  *nest = ModelVecCreateNest(3, sub);
  return PETSC_SUCCESS;
}
PetscErrorCode ISDestroy(IS *is)
{
  if (*is && (*is)->hdr.refct > 0 && --(*is)->hdr.refct == 0) ModelHeaderFree(*is), free(*is);
  *is = NULL;
  return PETSC_SUCCESS;
}

/* ------------------------------------------------------------------ DMStag creation, coordinates, DMComposite */
static PetscErrorCode dm_destroy_obj(PetscObject o)
{
  DM dm = (DM)o;
  return DMDestroy(&dm);
}
static DM stag_create(int dim, const int N[3], const int per[3], int d0, int d1, int d2, int d3, struct model_coords *coords)
{
  DM dm                 = ModelDMStagCreate(dim, N, per, d0, d1, d2, d3, coords);
  dm->hdr.destroy_model = dm_destroy_obj;
  return dm;
}
PetscErrorCode DMStagCreate2d(MPI_Comm c, DMBoundaryType bx, DMBoundaryType by, PetscInt M, PetscInt N, PetscInt m, PetscInt n, PetscInt d0, PetscInt d1, PetscInt d2, DMStagStencilType st, PetscInt sw, const PetscInt lx[], const PetscInt ly[], DM *dm)
{
  const int NN[3] = {M, N, 1}, per[3] = {bx == DM_BOUNDARY_PERIODIC, by == DM_BOUNDARY_PERIODIC, 0};
  (void)c, (void)lx, (void)ly, (void)st;
  PetscCheck((m == PETSC_DECIDE || m == 1) && (n == PETSC_DECIDE || n == 1), 0, PETSC_ERR_SUP, "the model has one rank");
  PetscCheck(sw == 1 && (bx == DM_BOUNDARY_NONE || bx == DM_BOUNDARY_PERIODIC) && (by == DM_BOUNDARY_NONE || by == DM_BOUNDARY_PERIODIC), 0, PETSC_ERR_SUP, "the model has stencil width 1 and NONE / PERIODIC boundaries");
  *dm = stag_create(2, NN, per, d0, d1, d2, 0, NULL);
  return PETSC_SUCCESS;
}
PetscErrorCode DMStagCreate3d(MPI_Comm c, DMBoundaryType bx, DMBoundaryType by, DMBoundaryType bz, PetscInt M, PetscInt N, PetscInt P, PetscInt m, PetscInt n, PetscInt p, PetscInt d0, PetscInt d1, PetscInt d2, PetscInt d3, DMStagStencilType st, PetscInt sw, const PetscInt lx[], const PetscInt ly[], const PetscInt lz[], DM *dm)
{
  const int NN[3] = {M, N, P}, per[3] = {bx == DM_BOUNDARY_PERIODIC, by == DM_BOUNDARY_PERIODIC, bz == DM_BOUNDARY_PERIODIC};
  (void)c, (void)lx, (void)ly, (void)lz, (void)st;
  PetscCheck((m == PETSC_DECIDE || m == 1) && (n == PETSC_DECIDE || n == 1) && (p == PETSC_DECIDE || p == 1), 0, PETSC_ERR_SUP, "the model has one rank");
  PetscCheck(sw == 1, 0, PETSC_ERR_SUP, "the model has stencil width 1");
  *dm = stag_create(3, NN, per, d0, d1, d2, d3, NULL);
  return PETSC_SUCCESS;
}
PetscErrorCode DMStagCreateCompatibleDMStag(DM dm, PetscInt d0, PetscInt d1, PetscInt d2, PetscInt d3, DM *out)
{
  *out = stag_create(dm->dim, dm->N, dm->per, d0, d1, d2, dm->dim == 3 ? d3 : 0, NULL);
  return PETSC_SUCCESS;
}
PetscErrorCode DMSetUp(DM dm) { return dm->setup = 1, PETSC_SUCCESS; }
PetscErrorCode DMDestroy(DM *pdm)
{
  DM dm = *pdm;
  if (!dm) return PETSC_SUCCESS;
  *pdm = NULL;
  if (--dm->hdr.refct > 0) return PETSC_SUCCESS;
  ModelDMDestroy(dm);
  return PETSC_SUCCESS;
}
PetscErrorCode DMSetMatrixPreallocateOnly(DM dm, PetscBool b) { return (void)dm, (void)b, PETSC_SUCCESS; }
PetscErrorCode DMStagSetRefinementFactor(DM dm, PetscInt a, PetscInt b, PetscInt c) { return (void)dm, (void)a, (void)b, (void)c, PETSC_SUCCESS; }
PetscErrorCode DMStagGetNumRanks(DM dm, PetscInt *x, PetscInt *y, PetscInt *z)
{
  if (x) *x = 1;
  if (y) *y = 1;
  if (z) *z = dm->dim == 3 ? 1 : MODEL_GARBAGE;
  return PETSC_SUCCESS;
}
PetscErrorCode DMStagGetOwnershipRanges(DM dm, const PetscInt *lx[], const PetscInt *ly[], const PetscInt *lz[])
{
  dm->ownership[0] = dm->N[0], dm->ownership[1] = dm->N[1], dm->ownership[2] = dm->N[2];
  if (lx) *lx = &dm->ownership[0];
  if (ly) *ly = &dm->ownership[1];
  if (lz) *lz = dm->dim == 3 ? &dm->ownership[2] : NULL;
  return PETSC_SUCCESS;
}
PetscErrorCode DMStagGetLocalSizes(DM dm, PetscInt *m, PetscInt *n, PetscInt *p)
{
  if (m) *m = dm->N[0];
  if (n) *n = dm->N[1];
  if (p) *p = dm->dim == 3 ? dm->N[2] : MODEL_GARBAGE;
  return PETSC_SUCCESS;
}
/* every local element gets its lower-vertex and centre coordinate, the partial element of a non-periodic direction and the ghost
   elements of a periodic one included: xmin + i h and xmin + (i + 1/2) h */
PetscErrorCode DMStagSetUniformCoordinatesProduct(DM dm, PetscReal x0, PetscReal x1, PetscReal y0, PetscReal y1, PetscReal z0, PetscReal z1)
{
  const double lo[3] = {x0, y0, z0}, hi[3] = {x1, y1, z1};
  int          d, li;
  if (!dm->coords) dm->coords = ModelCoordsCreate(dm->dim, dm->N, dm->per); /* an existing coordinate DM is updated in place: the DMs that share it see the new values */
  for (d = 0; d < dm->dim; ++d) {
    const double h = (hi[d] - lo[d]) / dm->N[d];
    for (li = 0; li < dm->coords->gn[d]; ++li) {
      const int g                  = li + dm->coords->gs[d];
      dm->coords->coord[d][2 * li] = lo[d] + g * h, dm->coords->coord[d][2 * li + 1] = lo[d] + (g + 0.5) * h;
    }
  }
  return PETSC_SUCCESS;
}
PetscErrorCode DMStagGetProductCoordinateArrays(DM dm, void *ax, void *ay, void *az) { return DMStagGetProductCoordinateArraysRead(dm, ax, ay, az); }
PetscErrorCode DMStagRestoreProductCoordinateArrays(DM dm, void *ax, void *ay, void *az) { return DMStagRestoreProductCoordinateArraysRead(dm, ax, ay, az); }
PetscErrorCode DMStagSetCoordinateDMType(DM dm, DMType t) { return (void)dm, (void)t, PETSC_SUCCESS; }
PetscErrorCode DMGetCoordinateDM(DM dm, DM *cdm) { return *cdm = dm, PETSC_SUCCESS; } /* the handle stands for the coordinates of dm */
PetscErrorCode DMSetCoordinateDM(DM dm, DM cdm)
{
  PetscCheck(cdm && cdm->coords, 0, PETSC_ERR_ARG_WRONGSTATE, "DMSetCoordinateDM: the source has no coordinates");
  if (dm->coords == cdm->coords) return PETSC_SUCCESS;
  ModelCoordsDestroy(dm->coords);
  dm->coords = cdm->coords, ++cdm->coords->refct;
  return PETSC_SUCCESS;
}
PetscErrorCode DMCompositeCreate(MPI_Comm c, DM *dm)
{
  (void)c;
  *dm = (DM)calloc(1, sizeof(**dm));
  ModelHeaderInit(*dm, 15, "DM", "composite", dm_destroy_obj);
  return PETSC_SUCCESS;
}
PetscErrorCode DMCompositeAddDM(DM c, DM sub)
{
  PetscCheck(c->ncomposite < 4, 0, PETSC_ERR_SUP, "composite of at most 4 DMs");
  c->composite[c->ncomposite++] = sub; /* borrowed: the mesh owns its DMs for the life of the NS */
  return PETSC_SUCCESS;
}
PetscErrorCode DMCompositeGetGlobalISs(DM c, IS *is[])
{
  int f;
  PetscCall(ModelMalloc(sizeof(IS) * (size_t)c->ncomposite, 1, is));
  for (f = 0; f < c->ncomposite; ++f) {
    (*is)[f] = (IS)calloc(1, sizeof(struct _p_IS));
    ModelHeaderInit((*is)[f], 16, "IS", "field", NULL);
    (*is)[f]->field = f, (*is)[f]->n = c->composite[f]->nglobal;
  }
  return PETSC_SUCCESS;
}

/* ------------------------------------------------------------------ SNES -> KSP -> PC */
static PetscFunctionList pc_list = NULL;
const char *const        snes_reasons_[] = {"DIVERGED_LINEAR_SOLVE", "", "", "CONVERGED_ITERATING", "", "", "CONVERGED_FNORM_RELATIVE"};
const char *const       *SNESConvergedReasons = snes_reasons_ + 3;
PetscErrorCode PCRegister(const char name[], PetscErrorCode (*create)(PC)) { return ModelFunctionListAdd(&pc_list, name, (void (*)(void))create); }
PetscErrorCode PCSetType(PC pc, PCType t)
{
  PetscErrorCode (*create)(PC) = NULL;
  PetscCall(ModelFunctionListFind(pc_list, t, (void (**)(void)) & create));
  PetscCheck(create, 0, PETSC_ERR_ARG_UNKNOWN_TYPE, "Unknown PC type %s (the model has the types the application registers)", t);
  if (pc->ops->destroy) PetscCall(pc->ops->destroy(pc));
  memset(pc->ops, 0, sizeof(pc->ops));
  PetscCall(PetscObjectChangeTypeName((PetscObject)pc, t));
  return create(pc);
}
PetscErrorCode KSPGetPC(KSP k, PC *pc)
{
  if (!k->pc) {
    k->pc = (PC)calloc(1, sizeof(*k->pc));
    ModelHeaderInit(k->pc, PC_CLASSID, "PC", NULL, NULL);
    if (k->hdr.prefix) k->pc->hdr.prefix = strdup(k->hdr.prefix);
  }
  *pc = k->pc;
  return PETSC_SUCCESS;
}
PetscErrorCode KSPSetTolerances(KSP k, PetscReal rtol, PetscReal abstol, PetscReal dtol, PetscInt maxits)
{
  (void)abstol, (void)dtol, (void)maxits;
  if (rtol > 0.) k->rtol = rtol;
  return PETSC_SUCCESS;
}
PetscErrorCode KSPSetNormType(KSP k, KSPNormType t) { return (void)k, (void)t, PETSC_SUCCESS; }
PetscErrorCode SNESCreate(MPI_Comm c, SNES *s)
{
  *s = (SNES)calloc(1, sizeof(**s));
  ModelHeaderInit(*s, 23, "SNES", "picard-model", NULL);
  (*s)->hdr.comm = c;
  return PETSC_SUCCESS;
}
PetscErrorCode SNESGetKSP(SNES s, KSP *k)
{
  if (!s->ksp) {
    PetscCall(KSPCreate(s->hdr.comm, &s->ksp));
    s->ksp->rtol = 1e-5;
    if (s->hdr.prefix) s->ksp->hdr.prefix = strdup(s->hdr.prefix);
  }
  *k = s->ksp;
  return PETSC_SUCCESS;
}
PetscErrorCode SNESDestroy(SNES *ps)
{
  SNES s = *ps;
  if (!s) return PETSC_SUCCESS;
  *ps = NULL;
  if (--s->hdr.refct > 0) return PETSC_SUCCESS;
  if (s->ksp) {
    if (s->ksp->pc) {
      if (s->ksp->pc->ops->destroy) PetscCall(s->ksp->pc->ops->destroy(s->ksp->pc));
      ModelHeaderFree(s->ksp->pc);
      free(s->ksp->pc);
    }
    PetscCall(KSPDestroy(&s->ksp));
  }
  ModelHeaderFree(s);
  free(s);
  return PETSC_SUCCESS;
}
PetscErrorCode SNESSetTolerances(SNES s, PetscReal a, PetscReal r, PetscReal st, PetscInt mi, PetscInt mf) { return (void)s, (void)a, (void)r, (void)st, (void)mi, (void)mf, PETSC_SUCCESS; }
PetscErrorCode SNESSetOptionsPrefix(SNES s, const char p[])
{
  free(s->hdr.prefix);
  s->hdr.prefix = p ? strdup(p) : NULL;
  return PETSC_SUCCESS;
}
PetscErrorCode SNESAppendOptionsPrefix(SNES s, const char p[])
{
  char buf[256];
  snprintf(buf, sizeof(buf), "%s%s", s->hdr.prefix ? s->hdr.prefix : "", p ? p : "");
  return SNESSetOptionsPrefix(s, buf);
}
/* -<prefix>ksp_type exact | preonly | gmres (default exact: the answer every convergent solver of the reference tends to),
   -<prefix>ksp_rtol for gmres; the options of PCABF through its own setfromoptions, with the prefix PETSc would prepend
   (abfpc.c:246-247 asks for -pc_abf_*_ainv_type, the user gives -ns_pc_abf_*_ainv_type) */
PetscErrorCode SNESSetFromOptions(SNES s)
{
  char      name[300], type[64] = "exact";
  PetscBool set;
  KSP       k;
  PC        pc;
  PetscCall(SNESGetKSP(s, &k));
  if (ModelKSPGetDefaultIterative()) strcpy(type, "gmres");
  snprintf(name, sizeof(name), "-%sksp_type", s->hdr.prefix ? s->hdr.prefix : "");
  PetscCall(ModelOptionsString(name, type, sizeof(type), &set));
  s->mode = !strcmp(type, "preonly") ? 1 : (!strcmp(type, "gmres") || !strcmp(type, "fgmres") ? 2 : 0);
  PetscCheck(s->mode || !strcmp(type, "exact"), 0, PETSC_ERR_SUP, "the model has -ksp_type exact, preonly and gmres, not %s", type);
  snprintf(name, sizeof(name), "-%sksp_rtol", s->hdr.prefix ? s->hdr.prefix : "");
  PetscCall(ModelOptionsReal(name, &k->rtol, NULL));
  PetscCall(KSPGetPC(k, &pc));
  if (pc->ops->setfromoptions) {
    ModelOptionsPrefixPush(s->hdr.prefix);
    PetscCall(pc->ops->setfromoptions(pc, NULL));
    ModelOptionsPrefixPop();
  }
  return PETSC_SUCCESS;
}
PetscErrorCode SNESSetPicard(SNES s, Vec r, PetscErrorCode (*b)(SNES, Vec, Vec, void *), Mat A, Mat P, PetscErrorCode (*J)(SNES, Vec, Mat, Mat, void *), void *ctx)
{
  s->r = r, s->bfunc = b, s->J = A, s->Jpre = P, s->jfunc = J, s->pctx = ctx;
  if (!s->func) s->func = SNESPicardComputeFunction, s->fctx = ctx;
  return PETSC_SUCCESS;
}
PetscErrorCode SNESSetFunction(SNES s, Vec r, PetscErrorCode (*f)(SNES, Vec, Vec, void *), void *ctx) { return s->r = r, s->func = f, s->fctx = ctx, PETSC_SUCCESS; }
PetscErrorCode SNESSetComputeInitialGuess(SNES s, PetscErrorCode (*g)(SNES, Vec, void *), void *ctx) { return s->guess = g, s->gctx = ctx, PETSC_SUCCESS; }
/* F(x) = A(x) x - b(x) */
PetscErrorCode SNESPicardComputeFunction(SNES s, Vec x, Vec f, void *ctx)
{
  Vec b;
  (void)ctx;
  PetscCall(VecDuplicate(f, &b));
  PetscCall(s->bfunc(s, x, b, s->pctx));
  PetscCall(s->jfunc(s, x, s->J, s->Jpre, s->pctx));
  PetscCall(MatMult(s->J, x, f));
  PetscCall(VecAXPY(f, -1., b));
  PetscCall(VecDestroy(&b));
  return PETSC_SUCCESS;
}
PetscErrorCode SNESMonitorCancel(SNES s) { return (void)s, PETSC_SUCCESS; }
PetscErrorCode SNESGetConvergedReason(SNES s, SNESConvergedReason *r) { return (void)s, *r = SNES_CONVERGED_FNORM_RELATIVE, PETSC_SUCCESS; }

static double nest_dot(Vec x, Vec y)
{
  double s = 0.;
  int    f, i;
  for (f = 0; f < 3; ++f)
    for (i = 0; i < x->sub[f]->n; ++i) s += x->sub[f]->a[i] * y->sub[f]->a[i];
  return s;
}
static PetscErrorCode linear_exact(SNES s, Vec b, Vec x)
{
  Mat       J     = s->J;
  const int nb[3] = {(int)b->sub[0]->n, (int)b->sub[1]->n, (int)b->sub[2]->n}, off[3] = {0, nb[0], nb[0] + nb[1]}, n = nb[0] + nb[1] + nb[2];
  const int border = J->nullspace ? 1 : 0, N = n + border;
  double   *a = (double *)calloc((size_t)N * N, sizeof(double)), *r = (double *)calloc((size_t)N, sizeof(double)), piv;
  int       bi, bj, i, q;
  for (bi = 0; bi < 3; ++bi)
    for (bj = 0; bj < 3; ++bj) {
      Mat B = J->blk[bi][bj];
      if (!B) continue;
      for (i = 0; i < B->m; ++i)
        for (q = 0; q < B->rn[i]; ++q) a[(size_t)(off[bi] + i) * N + off[bj] + B->rc[i][q]] += B->rv[i][q];
    }
  ModelVecGather(b, r);
  if (border) { /* the null vector of J (constant pressure): solution orthogonal to it, its component taken out of the residual */
    double *nv = (double *)calloc((size_t)n, sizeof(double));
    PetscCheck(J->nullspace->vec && !J->nullspace->has_cnst, 0, PETSC_ERR_SUP, "the coupled solve expects the one-vector null space of nsbasic.c:229-243");
    ModelVecGather(J->nullspace->vec, nv);
    for (i = 0; i < n; ++i) a[(size_t)i * N + n] = nv[i], a[(size_t)n * N + i] = nv[i];
    free(nv);
  }
  piv = ModelDenseSolve(N, a, r);
  free(a);
  if (!(piv > 1e-15)) {
    free(r);
    SETERRQ(0, PETSC_ERR_LIB, "coupled operator singular to working precision (pivot ratio %g)", piv);
  }
  ModelVecScatter(x, r);
  free(r);
  s->its = 1;
  return PETSC_SUCCESS;
}
/* right-preconditioned GMRES(30), zero guess; hist = residual norms (true residuals of x = M^-1 V y) */
static PetscErrorCode linear_gmres(SNES s, PC pc, Vec b, Vec x)
{
  enum { M = 30 };
  Vec          V[M + 1], z, w, r;
  double       H[M + 1][M], cs[M], sn[M], g[M + 1], y[M], rnorm, rnorm0;
  const double rtol = s->ksp->rtol;
  int          k, j, its = 0, done = 0;
  MatNullSpace ns = s->J->nullspace;
  PetscCall(VecDuplicate(b, &z));
  PetscCall(VecDuplicate(b, &w));
  PetscCall(VecDuplicate(b, &r));
  for (k = 0; k <= M; ++k) PetscCall(VecDuplicate(b, &V[k]));
  PetscCall(VecSet(x, 0.));
  PetscCall(VecCopy(b, r));
  rnorm0 = rnorm = sqrt(nest_dot(r, r));
  s->nhist = 0, s->hist[s->nhist++] = rnorm;
  while (!done && rnorm > rtol * rnorm0 && its < 10000) {
    PetscCall(VecCopy(r, V[0]));
    PetscCall(VecScale(V[0], 1. / rnorm));
    memset(g, 0, sizeof(g));
    g[0] = rnorm;
    for (k = 0; k < M; ++k) {
      PetscCall(pc->ops->apply(pc, V[k], z));
      PetscCall(MatMult(s->J, z, w));
      if (ns) PetscCall(MatNullSpaceRemove(ns, w));
      for (j = 0; j <= k; ++j) {
        H[j][k] = nest_dot(w, V[j]);
        PetscCall(VecAXPY(w, -H[j][k], V[j]));
      }
      H[k + 1][k] = sqrt(nest_dot(w, w));
      PetscCall(VecCopy(w, V[k + 1]));
      if (H[k + 1][k] > 0.) PetscCall(VecScale(V[k + 1], 1. / H[k + 1][k]));
      for (j = 0; j < k; ++j) {
        const double a = H[j][k], c = H[j + 1][k];
        H[j][k] = cs[j] * a + sn[j] * c, H[j + 1][k] = -sn[j] * a + cs[j] * c;
      }
      {
        const double a = H[k][k], c = H[k + 1][k], d = hypot(a, c);
        cs[k] = d > 0. ? a / d : 1., sn[k] = d > 0. ? c / d : 0.;
        H[k][k] = d, H[k + 1][k] = 0.;
        g[k + 1] = -sn[k] * g[k], g[k] = cs[k] * g[k];
      }
      ++its;
      rnorm = fabs(g[k + 1]);
      if (s->nhist < 256) s->hist[s->nhist++] = rnorm;
      if (rnorm <= rtol * rnorm0) {
        ++k;
        done = 1;
        break;
      }
    }
    for (j = k - 1; j >= 0; --j) {
      double sum = g[j];
      int    l;
      for (l = j + 1; l < k; ++l) sum -= H[j][l] * y[l];
      y[j] = sum / H[j][j];
    }
    PetscCall(VecSet(w, 0.));
    for (j = 0; j < k; ++j) PetscCall(VecAXPY(w, y[j], V[j]));
    PetscCall(pc->ops->apply(pc, w, z)); /* KSPGMRES builds the solution with one more application of the preconditioner per cycle */
    PetscCall(VecAXPY(x, 1., z));
    if (!done) {
      PetscCall(MatMult(s->J, x, w));
      PetscCall(VecWAXPY(r, -1., w, b));
      if (ns) PetscCall(MatNullSpaceRemove(ns, r));
      rnorm = sqrt(nest_dot(r, r));
    }
  }
  s->its = its;
  for (k = 0; k <= M; ++k) PetscCall(VecDestroy(&V[k]));
  PetscCall(VecDestroy(&z));
  PetscCall(VecDestroy(&w));
  PetscCall(VecDestroy(&r));
  return PETSC_SUCCESS;
}
/* the generic solve of the Picard form (SNESSetPicard + SNESSetFunction + SNESSetComputeInitialGuess): x0 = guess, F0 = F(x0),
   J dx = -F0 by the chosen linear solver, x = x0 + dx.  The problem is linear: one Picard iteration is the solution. */
PetscErrorCode ModelSNESSolvePicard(SNES s, Vec bunused, Vec x)
{
  Vec dx, rhs;
  PC  pc;
  (void)bunused;
  PetscCheck(s->func && s->jfunc && s->J, 0, PETSC_ERR_ARG_WRONGSTATE, "SNESSolve: the Picard callbacks are not set");
  if (s->guess) PetscCall(s->guess(s, x, s->gctx));
  else PetscCall(VecZeroEntries(x));
  PetscCall(s->func(s, x, s->r, s->fctx)); /* forms J as a side effect (SNESPicardComputeFunction) */
  PetscCall(VecDuplicate(x, &dx));
  PetscCall(VecDuplicate(x, &rhs));
  PetscCall(VecCopy(s->r, rhs));
  PetscCall(VecScale(rhs, -1.));
  s->nhist = 0;
  PetscCall(KSPGetPC(s->ksp, &pc));
  if (s->mode == 0) PetscCall(linear_exact(s, rhs, dx));
  else {
    pc->mat = s->J, pc->pmat = s->Jpre ? s->Jpre : s->J;
    PetscCheck(pc->ops->apply, 0, PETSC_ERR_ARG_WRONGSTATE, "the PC has no type");
    if (pc->ops->setup) PetscCall(pc->ops->setup(pc));
    if (s->mode == 1) {
      PetscCall(pc->ops->apply(pc, rhs, dx));
      s->its = 1;
    } else PetscCall(linear_gmres(s, pc, rhs, dx));
  }
  PetscCall(VecAXPY(x, 1., dx));
  PetscCall(VecDestroy(&dx));
  PetscCall(VecDestroy(&rhs));
  return PETSC_SUCCESS;
}

/* ------------------------------------------------------------------ what the reference keeps in its viewer package (needs CGNS) */
PetscErrorCode PetscViewerCreate_FlucaCGNS(PetscViewer v) { return (void)v, PETSC_ERR_SUP; }
/* -<prefix><name> [ascii]: view the object (MeshViewFromOptions, NSViewFromOptions); the model only acknowledges the option */
PetscErrorCode FlucaObjectViewFromOptions(PetscObject obj, PetscObject bobj, const char optionname[])
{
  (void)obj, (void)bobj;
  (void)opt_find(optionname);
  return PETSC_SUCCESS;
}
