// Minimal sequential stand-in for the slice of the PETSc 3.x C API that the reference's
// diffuclass.cpp calls (DMDA 2-D grid, shell matrix, KSP), written here so that the reference's own
// diffusionPETSc class can be compiled IN PLACE from /root/reference and run on one process as a parity
// pin for the oracle (oracle/Makefile -> oracle/_ref/libeq_fd_ref.so).  TEST INFRASTRUCTURE ONLY.
//
// What is real: the grid vectors, the row-pointer array views (DMDAVecGetArray), the shell-matrix
// dispatch to the reference's MyMatMult, global->local copies.  What is replaced: KSPSolve (PETSc's
// KSPFBCGSR [ext]) is an unpreconditioned BiCGStab on the shell matrix with zero initial guess, stopping
// at ||r|| <= rtol ||b|| (rtol: petsc_shim_rtol, PETSc's default 1e-5).  Parallel-only calls
// (AO / IS / VecScatter) and viewers are inert: the shim runs with a communicator of size one.
#pragma once
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <vector>

typedef int PetscErrorCode;
typedef int PetscInt;
typedef double PetscReal;
typedef double PetscScalar;
typedef enum { PETSC_FALSE = 0, PETSC_TRUE = 1 } PetscBool;
typedef int MPI_Comm_shim;

struct _p_DM { int nx, ny; };
typedef _p_DM *DM;
struct _p_Vec { DM dm; std::vector<double> data; std::vector<double *> rows; };
typedef _p_Vec *Vec;
struct _p_Mat;
typedef _p_Mat *Mat;
struct _p_Mat { void *ctx; PetscErrorCode (*mult)(Mat, Vec, Vec); };
struct _p_KSP;
typedef _p_KSP *KSP;
struct _p_KSP { DM dm; PetscErrorCode (*compute)(KSP, Mat, Mat, void *); void *user; Mat A; };
typedef void *PC;
typedef void *AO;
typedef void *IS;
typedef void *VecScatter;
typedef void *PetscViewer;
typedef int DMBoundaryType;
typedef int DMDAStencilType;
typedef int InsertMode;
typedef int ScatterMode;
typedef int PetscCopyMode;
typedef int MatOperation;

enum { DM_BOUNDARY_NONE = 0, DMDA_STENCIL_STAR = 0, INSERT_VALUES = 1, ADD_VALUES = 2, SCATTER_FORWARD = 0,
       PETSC_COPY_VALUES = 0, PETSC_DECIDE = -1, MATOP_MULT = 3, FILE_MODE_WRITE = 1, PETSC_VIEWER_VTK_VTR = 1 };
#define MATSHELL "shell"
#define PCNONE "none"
#define KSPFBCGSR "fbcgsr"
#define PETSC_COMM_SELF 0
#define PETSC_VIEWER_STDOUT_(comm) ((PetscViewer)0)
#define CHKERRQ(e) do { if (e) return (e); } while (0)
#define PetscFunctionBegin
#define PetscFunctionReturn(x) return (x)

// test hooks: tolerance of the stand-in Krylov solve, the last shell matrix that was set up, the right-hand
// side and iteration count of the last KSPSolve
extern double petsc_shim_rtol;
extern Mat petsc_shim_last_mat;
extern std::vector<double> petsc_shim_last_rhs;
extern int petsc_shim_last_its;

inline PetscErrorCode PetscInitialize(int *, char ***, const char *, const char *) { return 0; }
inline PetscErrorCode PetscFinalize() { return 0; }
inline PetscErrorCode PetscPrintf(int, const char *, ...) { return 0; }

inline Vec petsc_shim_new_vec(DM dm)
{
    Vec v = new _p_Vec;
    v->dm = dm;
    v->data.assign((size_t)dm->nx * dm->ny, 0.0);
    v->rows.resize(dm->ny);
    for (int j = 0; j < dm->ny; ++j) v->rows[j] = v->data.data() + (size_t)j * dm->nx;
    return v;
}

inline PetscErrorCode KSPCreate(int, KSP *k) { *k = new _p_KSP{nullptr, nullptr, nullptr, nullptr}; return 0; }
inline PetscErrorCode DMDACreate2d(int, DMBoundaryType, DMBoundaryType, DMDAStencilType, PetscInt M, PetscInt N,
                                   PetscInt, PetscInt, PetscInt, PetscInt, const PetscInt *, const PetscInt *, DM *dm)
{
    *dm = new _p_DM{M, N};
    return 0;
}
inline PetscErrorCode DMSetFromOptions(DM) { return 0; }
inline PetscErrorCode DMSetUp(DM) { return 0; }
inline PetscErrorCode DMDASetUniformCoordinates(DM, double, double, double, double, double, double) { return 0; }
inline PetscErrorCode DMView(DM, PetscViewer) { return 0; }
inline PetscErrorCode DMCreateGlobalVector(DM dm, Vec *v) { *v = petsc_shim_new_vec(dm); return 0; }
inline PetscErrorCode DMCreateLocalVector(DM dm, Vec *v) { *v = petsc_shim_new_vec(dm); return 0; }
inline PetscErrorCode VecGetOwnershipRange(Vec v, PetscInt *lo, PetscInt *hi)
{
    *lo = 0; *hi = (PetscInt)v->data.size();
    return 0;
}
inline PetscErrorCode KSPSetDM(KSP k, DM dm) { k->dm = dm; return 0; }
inline PetscErrorCode KSPSetComputeOperators(KSP k, PetscErrorCode (*f)(KSP, Mat, Mat, void *), void *user)
{
    k->compute = f; k->user = user;
    return 0;
}
inline PetscErrorCode KSPGetPC(KSP, PC *pc) { *pc = nullptr; return 0; }
inline PetscErrorCode PCSetType(PC, const char *) { return 0; }
inline PetscErrorCode KSPSetType(KSP, const char *) { return 0; }
inline PetscErrorCode KSPSetFromOptions(KSP) { return 0; }
inline PetscErrorCode KSPSetUp(KSP k)
{
    if (!k->A && k->compute) {
        k->A = new _p_Mat{nullptr, nullptr};
        PetscErrorCode e = k->compute(k, k->A, k->A, k->user);
        petsc_shim_last_mat = k->A;
        return e;
    }
    return 0;
}
inline PetscErrorCode DMDAGetAO(DM, AO *ao) { *ao = nullptr; return 0; }
inline PetscErrorCode DMDAGetCorners(DM dm, PetscInt *xs, PetscInt *ys, PetscInt *zs, PetscInt *xm, PetscInt *ym, PetscInt *zm)
{
    if (xs) *xs = 0; if (ys) *ys = 0; if (zs) *zs = 0;
    if (xm) *xm = dm->nx; if (ym) *ym = dm->ny; if (zm) *zm = 1;
    return 0;
}
// one process, DM_BOUNDARY_NONE: the ghosted patch is the grid itself
inline PetscErrorCode DMDAGetGhostCorners(DM dm, PetscInt *xs, PetscInt *ys, PetscInt *zs, PetscInt *xm, PetscInt *ym, PetscInt *zm)
{
    return DMDAGetCorners(dm, xs, ys, zs, xm, ym, zm);
}
inline PetscErrorCode DMDAVecGetArray(DM, Vec v, void *array) { *(double ***)array = v->rows.data(); return 0; }
inline PetscErrorCode DMDAVecRestoreArray(DM, Vec, void *) { return 0; }
inline PetscErrorCode DMDAVecGetArrayRead(DM, Vec v, void *array) { *(double ***)array = v->rows.data(); return 0; }
inline PetscErrorCode DMDAVecRestoreArrayRead(DM, Vec, void *) { return 0; }
inline PetscErrorCode DMGlobalToLocalBegin(DM, Vec g, InsertMode, Vec l) { l->data = g->data; return 0; }
inline PetscErrorCode DMGlobalToLocalEnd(DM, Vec, InsertMode, Vec) { return 0; }
inline PetscErrorCode VecGetDM(Vec v, DM *dm) { *dm = v->dm; return 0; }
inline PetscErrorCode MatSetSizes(Mat, PetscInt, PetscInt, PetscInt, PetscInt) { return 0; }
inline PetscErrorCode MatSetType(Mat, const char *) { return 0; }
inline PetscErrorCode MatSetUp(Mat) { return 0; }
inline PetscErrorCode MatShellSetContext(Mat A, void *ctx) { A->ctx = ctx; return 0; }
inline PetscErrorCode MatShellGetContext(Mat A, void *ctx) { *(void **)ctx = A->ctx; return 0; }
inline PetscErrorCode MatShellSetOperation(Mat A, MatOperation, void (*f)(void))
{
    A->mult = (PetscErrorCode(*)(Mat, Vec, Vec))f;
    return 0;
}
inline PetscErrorCode VecDestroy(Vec *v) { if (v && *v) { delete *v; *v = nullptr; } return 0; }
inline PetscErrorCode KSPDestroy(KSP *k) { if (k && *k) { delete (*k)->A; delete *k; *k = nullptr; } return 0; }
inline PetscErrorCode DMDestroy(DM *d) { if (d && *d) { delete *d; *d = nullptr; } return 0; }
// inert: parallel read path (never taken with one process) and viewers
inline PetscErrorCode AOApplicationToPetsc(AO, PetscInt, PetscInt *) { return 0; }
inline PetscErrorCode ISCreateGeneral(int, PetscInt, const PetscInt *, PetscCopyMode, IS *is) { *is = nullptr; return 0; }
inline PetscErrorCode ISDestroy(IS *) { return 0; }
inline PetscErrorCode VecCreateSeq(int, PetscInt, Vec *v) { *v = nullptr; return 0; }
inline PetscErrorCode VecScatterCreate(Vec, IS, Vec, IS, VecScatter *s) { *s = nullptr; return 0; }
inline PetscErrorCode VecScatterBegin(VecScatter, Vec, Vec, InsertMode, ScatterMode) { return 0; }
inline PetscErrorCode VecScatterEnd(VecScatter, Vec, Vec, InsertMode, ScatterMode) { return 0; }
inline PetscErrorCode VecScatterDestroy(VecScatter *) { return 0; }
inline PetscErrorCode VecGetArrayRead(Vec, const PetscScalar **a) { *a = nullptr; return 0; }
inline PetscErrorCode VecRestoreArrayRead(Vec, const PetscScalar **) { return 0; }
inline PetscErrorCode PetscViewerVTKOpen(int, const char *, int, PetscViewer *v) { *v = nullptr; return 0; }
inline PetscErrorCode PetscViewerPushFormat(PetscViewer, int) { return 0; }
inline PetscErrorCode PetscViewerDestroy(PetscViewer *) { return 0; }
inline PetscErrorCode VecView(Vec, PetscViewer) { return 0; }

// KSPSolve(ksp, b, x) with b == x allowed (the reference passes the same vector twice; PETSc then works on
// a copy of b): unpreconditioned BiCGStab on the shell matrix, zero initial guess.
PetscErrorCode KSPSolve(KSP k, Vec b, Vec x);
