#include "gpuFD.h"

#include <cstdio>
#include <fstream>

gpuFD::~gpuFD() { finalize(); }

void gpuFD::finalize()
{
    if (h) { eqgpu_destroy(h); h = nullptr; }
}

// Wall coefficients -> the C ABI's (type, value, s): Nc == 0 Dirichlet value BV/Dc ("if N==0, then D must
// be 1", diffuclass.cpp:232); Dc == 0 homogeneous Neumann; else Robin rate r = D*Dc/Nc, s = BV/Dc.
void gpuFD::create()
{
    finalize();
    eqgpu_params p;
    eqgpu_default_params(&p);
    p.discretisation = EQGPU_DISC_FD;
    p.nW = int(gridNodesX);
    p.nH = int(gridNodesY);
    p.hx = p.hy = initData.h;
    p.dt = initData.dt;
    p.D = initData.diffusionConstant;
    p.device = device;
    const double Dc[4] = {initData.leftDirichletCoefficient, initData.rightDirichletCoefficient,
                          initData.topDirichletCoefficient, initData.bottomDirichletCoefficient};
    const double Nc[4] = {initData.leftNeumannCoefficient, initData.rightNeumannCoefficient,
                          initData.topNeumannCoefficient, initData.bottomNeumannCoefficient};
    const double BV[4] = {initData.leftBoundaryValue, initData.rightBoundaryValue, initData.topBoundaryValue,
                          initData.bottomBoundaryValue};
    for (int w = 0; w < 4; ++w) {
        if (Nc[w] == 0.0) {
            p.bc_type[w] = EQGPU_BC_DIRICHLET;
            p.bc_value[w] = BV[w];   // ApplyBoundaryConditions writes RHS = BV and ignores Dc ("if N==0, then D must be 1", diffuclass.cpp:232)
        } else if (Dc[w] == 0.0) {
            if (BV[w] != 0.0) throw std::runtime_error("gpuFD: a non-zero pure-Neumann flux is not supported");
            p.bc_type[w] = EQGPU_BC_NEUMANN;
            p.bc_value[w] = 0.0;
        } else {
            if (w == EQGPU_TOP || w == EQGPU_BOTTOM) throw std::runtime_error("gpuFD: Robin rows on the left/right walls only");
            p.bc_type[w] = EQGPU_BC_ROBIN;
            p.bc_value[w] = initData.diffusionConstant * Dc[w] / Nc[w];
            p.robin_s[w] = BV[w] / Dc[w];
        }
    }
    int rc = eqgpu_create(&p, &h);
    if (rc != EQGPU_OK) throw std::runtime_error(std::string("gpuFD: eqgpu_create failed: ") + eqgpu_last_error(nullptr));
}

// diffuclass.cpp:19-106
void gpuFD::initDiffusion(eQ::diffusionSolver::params &initParams)
{
    myParams = initParams;
    initData.diffusionConstant = initParams.D_HSL;
    initData.xLengthMicrons = initParams.trapWidthMicrons;
    initData.yLengthMicrons = initParams.trapHeightMicrons;
    initData.dt = initParams.dt;
    initData.h = 1.0 / initParams.nodesPerMicron;
    initData.directoryName = initParams.filePath + "petsc";
    initData.objectName = "grid";
    if (boundaryType == "DIRICHLET_0") {   // :68-86, the only wiring upstream ships
        initData.homogeneousDirichlet = true;
        initData.topDirichletCoefficient = initData.bottomDirichletCoefficient = 1;
        initData.leftDirichletCoefficient = initData.rightDirichletCoefficient = 1;
        initData.topNeumannCoefficient = initData.bottomNeumannCoefficient = 0;
        initData.leftNeumannCoefficient = initData.rightNeumannCoefficient = 0;
        initData.topBoundaryValue = initData.bottomBoundaryValue = 0;
        initData.leftBoundaryValue = initData.rightBoundaryValue = 0;
    }
    // InitializeDiffusion (:357-361): PetscInt arithmetic on the integer lengths
    gridNodesX = size_t((long)(initData.xLengthMicrons) / initData.h) + 1;
    gridNodesY = size_t((long)(initData.yLengthMicrons) / initData.h) + 1;
    initData.fourierNumber = (initData.diffusionConstant * initData.dt) / (initData.h * initData.h);
    create();
    solution_vector.assign(gridNodesX * gridNodesY, 0.0);   // :92-106
}

void gpuFD::applyBoundaryCoefficients()
{
    create();
}

// diffuclass.cpp:108-118: WriteGridValues -> TimeStep -> ReadGridValues; the natural order of
// allXCoordinates / allYCoordinates (i + j*gridNodesX, :96-99) is the device order, so the two
// application-ordering scatters are plain copies.
void gpuFD::stepDiffusion()
{
    int rc = eqgpu_step_host(h, solution_vector.data());
    if (rc != EQGPU_OK) throw std::runtime_error(std::string("gpuFD: eqgpu_step_host failed: ") + eqgpu_last_error(h));
    eqgpu_stats st;
    if (eqgpu_get_stats(h, &st) == EQGPU_OK) totalBoundaryFlux = st.total_boundary_flux;
}

// Upstream's loop (diffuclass.cpp:135-166, "TODO: verify accuracy") accumulates into an uninitialised
// double and strides rows by gridNodesX-1, so it has no defined value to reproduce; this returns the
// boundary functional D*dt*(-oint grad u . n ds) that fenicsInterface reports (src/fHSL.cpp:156-160).
eQ::data::parametersType gpuFD::getBoundaryFlux(void)
{
    eQ::data::parametersType j;
    j["totalFlux"] = totalBoundaryFlux;
    return j;
}

int gpuFD::lastIterations() const
{
    eqgpu_stats st;
    if (eqgpu_get_stats(h, &st) != EQGPU_OK) return -1;
    return st.iterations;
}

// RecordData (diffuclass.cpp:560-583) writes a PETSc binary viewer file per step; without PETSc the same
// snapshot goes out as legacy-VTK structured points, like gpuHSL's.
void gpuFD::writeDiffusionFiles(double timestamp)
{
    if (myParams.filePath.empty()) return;
    char name[512];
    snprintf(name, sizeof name, "%s_%s_%012.4f.vtk", initData.directoryName.c_str(), initData.objectName.c_str(), timestamp);
    std::ofstream f(name);
    if (!f) return;
    f.precision(17);   // round-trip doubles
    f << "# vtk DataFile Version 3.0\nHSL t=" << timestamp << "\nASCII\nDATASET STRUCTURED_POINTS\n";
    f << "DIMENSIONS " << gridNodesX << " " << gridNodesY << " 1\nORIGIN 0 0 0\nSPACING " << initData.h << " "
      << initData.h << " 1\n";
    f << "POINT_DATA " << gridNodesX * gridNodesY << "\nSCALARS u double 1\nLOOKUP_TABLE default\n";
    for (double v : solution_vector) f << v << "\n";
}
