// Builds oracle/_ref/libeq_fd_ref.so: the reference's OWN finite-difference solver class, diffusionPETSc,
// compiled in place from /root/reference/diffuclass.{h,cpp} (no copy of the sources enters this repository)
// on top of the one-process PETSc/MPI/boost interface shim in oracle/shim_petsc/ and the nlohmann/json.hpp
// the reference vendors.  TEST INFRASTRUCTURE ONLY: it pins the oracle's restatement of MyMatMult /
// ApplyBoundaryConditions / stepDiffusion (oracle/eq_oracle.c eqo_fd_*, oracle/oracle.py fd_*) and generates
// tests/golden/fd_ref.json.  Everything goes through the class's PUBLIC surface: initData (the wall
// coefficients upstream tells callers to write directly, diffuclass.cpp:121-124), the per-node boundary
// vectors, initDiffusion, solution_vector, stepDiffusion.
#include "diffuclass.h"   // /root/reference/diffuclass.h (-I$(REF))

#include <cstring>

// src/main.cpp:44 defines this static in the executable; the parity pin is not linked against main.cpp
eQ::data::parametersType eQ::data::parameters;

double petsc_shim_rtol = 1e-5;   // PETSc's default KSP rtol [ext]
Mat petsc_shim_last_mat = nullptr;
std::vector<double> petsc_shim_last_rhs;
int petsc_shim_last_its = 0;

static double dotv(const std::vector<double> &a, const std::vector<double> &b)
{
    double s = 0.0;
    for (size_t i = 0; i < a.size(); ++i) s += a[i] * b[i];
    return s;
}

PetscErrorCode KSPSolve(KSP k, Vec bin, Vec x)
{
    PetscErrorCode e = KSPSetUp(k);
    if (e) return e;
    Mat A = k->A;
    DM dm = bin->dm;
    const std::vector<double> b = bin->data;   // copy: b and x may be the same vector
    petsc_shim_last_rhs = b;
    const size_t n = b.size();
    Vec p = petsc_shim_new_vec(dm), v = petsc_shim_new_vec(dm), s = petsc_shim_new_vec(dm), t = petsc_shim_new_vec(dm);
    std::vector<double> r = b, rh = b, sol(n, 0.0);
    double bb = dotv(b, b);
    if (bb == 0.0) bb = 1.0;
    const double stop = petsc_shim_rtol * petsc_shim_rtol * bb;
    double rho = 1.0, alpha = 1.0, omega = 1.0, rr = dotv(r, r);
    int it = 0;
    while (rr > stop && it < 100000) {
        const double rho1 = dotv(rh, r);
        const double beta = (rho1 / rho) * (alpha / omega);
        for (size_t i = 0; i < n; ++i) p->data[i] = r[i] + beta * (p->data[i] - omega * v->data[i]);
        A->mult(A, p, v);
        alpha = rho1 / dotv(rh, v->data);
        for (size_t i = 0; i < n; ++i) s->data[i] = r[i] - alpha * v->data[i];
        A->mult(A, s, t);
        const double tt = dotv(t->data, t->data);
        omega = tt > 0.0 ? dotv(t->data, s->data) / tt : 0.0;
        for (size_t i = 0; i < n; ++i) {
            sol[i] += alpha * p->data[i] + omega * s->data[i];
            r[i] = s->data[i] - omega * t->data[i];
        }
        rr = dotv(r, r);
        rho = rho1;
        ++it;
        if (omega == 0.0) break;
    }
    petsc_shim_last_its = it;
    x->data = sol;
    VecDestroy(&p); VecDestroy(&v); VecDestroy(&s); VecDestroy(&t);
    return 0;
}

#define REF_API extern "C" __attribute__((visibility("default")))

struct FdRef {
    diffusionPETSc solver;
    eQ::diffusionSolver::params prm;
};

// Wall arrays ordered left, right, top, bottom (Dc u + Nc du/dn = BV).  Lengths are the integer microns
// upstream passes (PetscInt xLengthMicrons, diffuclass.h:17).
REF_API void *ref_fd_create(int widthMicrons, int heightMicrons, double nodesPerMicron, double dt, double D,
                            const double *Dc, const double *Nc, const double *BV)
{
    eQ::data::parameters["boundaryType"] = "SET_BY_CALLER";   // anything but DIRICHLET_0: initDiffusion leaves the walls alone
    FdRef *f = new FdRef();
    DiffusionData &d = f->solver.initData;
    d.leftDirichletCoefficient = Dc[0]; d.rightDirichletCoefficient = Dc[1];
    d.topDirichletCoefficient = Dc[2]; d.bottomDirichletCoefficient = Dc[3];
    d.leftNeumannCoefficient = Nc[0]; d.rightNeumannCoefficient = Nc[1];
    d.topNeumannCoefficient = Nc[2]; d.bottomNeumannCoefficient = Nc[3];
    d.leftBoundaryValue = BV[0]; d.rightBoundaryValue = BV[1];
    d.topBoundaryValue = BV[2]; d.bottomBoundaryValue = BV[3];
    f->prm = eQ::diffusionSolver::params();
    f->prm.argc = 0; f->prm.argv = nullptr; f->prm.comm = 0;
    f->prm.dt = dt; f->prm.D_HSL = D;
    f->prm.filePath = "/tmp/eq_fd_ref_";
    f->prm.trapWidthMicrons = widthMicrons; f->prm.trapHeightMicrons = heightMicrons;
    f->prm.nodesPerMicron = nodesPerMicron;
    f->solver.initDiffusion(f->prm);
    // ApplyBoundaryConditions indexes these per node (diffuclass.cpp:218,236); upstream never sizes them
    const size_t nx = (size_t)((int)(widthMicrons / (1.0 / nodesPerMicron)) + 1);
    f->solver.topBoundaryValue.assign(nx, BV[2]);
    f->solver.bottomBoundaryValue.assign(nx, BV[3]);
    return f;
}

REF_API long ref_fd_size(void *h) { return (long)((FdRef *)h)->solver.solution_vector.size(); }

// the DIRICHLET_0 wiring initDiffusion itself provides (diffuclass.cpp:68-86)
REF_API void *ref_fd_create_dirichlet0(int widthMicrons, int heightMicrons, double nodesPerMicron, double dt, double D)
{
    eQ::data::parameters["boundaryType"] = "DIRICHLET_0";
    FdRef *f = new FdRef();
    f->prm = eQ::diffusionSolver::params();
    f->prm.argc = 0; f->prm.argv = nullptr; f->prm.comm = 0;
    f->prm.dt = dt; f->prm.D_HSL = D;
    f->prm.filePath = "/tmp/eq_fd_ref_";
    f->prm.trapWidthMicrons = widthMicrons; f->prm.trapHeightMicrons = heightMicrons;
    f->prm.nodesPerMicron = nodesPerMicron;
    f->solver.initDiffusion(f->prm);
    const size_t nx = (size_t)((int)(widthMicrons / (1.0 / nodesPerMicron)) + 1);
    f->solver.topBoundaryValue.assign(nx, 0.0);
    f->solver.bottomBoundaryValue.assign(nx, 0.0);
    return f;
}

// diffusionPETSc::stepDiffusion (diffuclass.cpp:108-118) on u (in/out); rhs_out (optional) receives the
// right-hand side ApplyBoundaryConditions produced; returns the Krylov iteration count of the stand-in solve.
REF_API int ref_fd_step(void *h, double *u, double rtol, double *rhs_out)
{
    FdRef *f = (FdRef *)h;
    const size_t n = f->solver.solution_vector.size();
    petsc_shim_rtol = rtol;
    std::memcpy(f->solver.solution_vector.data(), u, sizeof(double) * n);
    f->solver.stepDiffusion();
    std::memcpy(u, f->solver.solution_vector.data(), sizeof(double) * n);
    if (rhs_out) std::memcpy(rhs_out, petsc_shim_last_rhs.data(), sizeof(double) * n);
    return petsc_shim_last_its;
}

// y = A x through the shell matrix the class registered: the reference's MyMatMult (diffuclass.cpp:637-872).
// Needs one ref_fd_step before (KSPSetUp creates the shell matrix on first use, as PETSc does).
REF_API int ref_fd_matmult(void *h, const double *x, double *y)
{
    FdRef *f = (FdRef *)h;
    Mat A = petsc_shim_last_mat;
    if (!A || !A->mult) return -1;
    _p_DM dm{(int)0, (int)0};
    const size_t n = f->solver.solution_vector.size();
    // grid shape from the class's public coordinate count: rows = n / nx, nx from the top boundary vector
    dm.nx = (int)f->solver.topBoundaryValue.size();
    dm.ny = (int)(n / (size_t)dm.nx);
    Vec X = petsc_shim_new_vec(&dm), Y = petsc_shim_new_vec(&dm);
    std::memcpy(X->data.data(), x, sizeof(double) * n);
    A->mult(A, X, Y);
    std::memcpy(y, Y->data.data(), sizeof(double) * n);
    VecDestroy(&X); VecDestroy(&Y);
    return 0;
}

REF_API void ref_fd_destroy(void *h)
{
    FdRef *f = (FdRef *)h;
    if (!f) return;
    f->solver.finalize();
    delete f;
}
