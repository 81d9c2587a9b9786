// gpuFD -- drop-in for diffusionPETSc (diffuclass.h:34-97, diffuclass.cpp), the reference's
// finite-difference eQ::diffusionSolver, on a B200.
//
// Same public surface: solution_vector, initData (the DiffusionData coefficients a caller edits to
// "set boundaries explicitly by writing the data structures", diffuclass.cpp:123-124), initDiffusion /
// stepDiffusion / setBoundaryValues / getBoundaryFlux / getDiffusionConstant / writeDiffusionFiles /
// finalize.  The numerical work goes through the C ABI (include/eqgpu.h) with EQGPU_DISC_FD: the same
// 5-point ghost-node system MyMatMult applies (diffuclass.cpp:786-862), solved by multigrid-PCG to
// rtol 1e-12 instead of unpreconditioned FBCGSR to PETSc's default 1e-5.
#pragma once
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/eqgpu.h"
#include "eq_compat.h"

class gpuFD : public eQ::diffusionSolver {
public:
    // the boundary slice of DiffusionData (diffuclass.h:16-32): Dc*u + Nc*du/dn = BV per wall
    struct DiffusionData {
        double xLengthMicrons = 0, yLengthMicrons = 0, diffusionConstant = 0, h = 0, dt = 0, fourierNumber = 0;
        double topDirichletCoefficient = 1, bottomDirichletCoefficient = 1, leftDirichletCoefficient = 1,
               rightDirichletCoefficient = 1;
        double topNeumannCoefficient = 0, bottomNeumannCoefficient = 0, leftNeumannCoefficient = 0,
               rightNeumannCoefficient = 0;
        double topBoundaryValue = 0, bottomBoundaryValue = 0, leftBoundaryValue = 0, rightBoundaryValue = 0;
        bool homogeneousDirichlet = true;
        std::string directoryName, objectName;
    };

    gpuFD() = default;
    ~gpuFD();   // eQ::diffusionSolver has no virtual destructor (src/eQ.h:302-330)

    std::string boundaryType = "DIRICHLET_0";   // eQ::data::parameters["boundaryType"] (diffuclass.cpp:68)
    int device = 0;
    std::vector<double> solution_vector;        // diffuclass.h:71
    DiffusionData initData, *gridData = &initData;
    size_t gridNodesX = 0, gridNodesY = 0;

    void initDiffusion(eQ::diffusionSolver::params &) override;  // diffuclass.cpp:19-106
    void stepDiffusion() override;                               // diffuclass.cpp:108-118
    // upstream returns at once ("over-ride for now", diffuclass.cpp:121-124): kept a no-op
    void setBoundaryValues(const eQ::data::parametersType &) {}
    eQ::data::parametersType getBoundaryFlux(void) override;     // diffuclass.cpp:135-166 (see gpuFD.cpp)
    double getDiffusionConstant(void) { return initData.diffusionConstant; }
    void writeDiffusionFiles(double timestamp) override;
    void finalize(void) override;
    // Re-reads initData's wall coefficients (a caller that edits them after initDiffusion calls this;
    // upstream MyMatMult reads them on every product).
    void applyBoundaryCoefficients();

    eqgpu_solver *handle() { return h; }
    int lastIterations() const;

private:
    eqgpu_solver *h = nullptr;
    eQ::diffusionSolver::params myParams;
    double totalBoundaryFlux = 0.0;
    void create();
};
