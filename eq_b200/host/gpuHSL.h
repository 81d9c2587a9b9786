// gpuHSL -- drop-in for fenicsInterface (src/fHSL.h:333-449) on a B200.
//
// It implements eQ's own solver interface, eQ::diffusionSolver
// (src/eQ.h:302-330), and additionally exposes the de-facto members that
// src/simulation.cpp reaches into on the Fenics class:
//   solution_vector, D11/D22/D12, totalBoundaryFlux, setBoundaryValues(),
//   topChannelData/bottomChannelData, nodesH/nodesW, and a `shell` carrying
//   mesh_coords / dof_from_vertex (identity) / num_vertices().
// All numerical work goes through the C ABI in include/eqgpu.h.
//
// The reference reads its configuration from the global eQ::data::parameters
// JSON (src/fHSL.cpp:45-47,110,117,336-341,446-570); here the same keys arrive
// as an explicit gpuHSL::config filled by the caller (INTEGRATION.md shows the
// six lines that copy them from the JSON inside Simulation::create_HSLgrid).
#pragma once
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/eqgpu.h"
#include "eq_compat.h"

class gpuHSL : public eQ::diffusionSolver {
public:
    // the eQ::data::parameters keys fenicsInterface reads
    struct config {
        std::string boundaryType = "DIRICHLET_0";  // "DIRICHLET_0" | "DIRICHLET_UPDATE" | "MICROFLUIDIC_TRAP" | ...
        std::string trapType = "NOWALLED";         // "NOWALLED" | "THREEWALLED" | "TWOWALLED" | "ONEWALLED" | "H_TRAP"
        // parameters["boundaries"][wall][1] = {a, b, v} (src/eQ.h:399-419), order left,right,top,bottom
        double boundaries[4][3] = {{0, 1, 0}, {0, 1, 0}, {0, 1, 0}, {0, 1, 0}};
        double lengthScaling = 5.0;                 // src/main.cpp:511
        double simulationFlowRate = 120.0;          // src/main.cpp:364
        double simulationChannelLengthLeft = 20.0;  // src/main.cpp:333-336
        double simulationChannelLengthRight = 20.0;
        int channelSolverNumberIterations = 48;     // src/main.cpp:426-429
        int device = 0;                             // CUDA device of this layer
        // D11/D22/D12 are host vectors the controller may overwrite at any time (src/simulation.cpp:503-505); finding out
        // costs a pass over 3N doubles.  -1: look every step up to 2^18 nodes, above that only after
        // notifyTensorChanged(); 0: only after notifyTensorChanged(); 1: look every step whatever the size
        int tensorScan = -1;
        double rtol = 1e-12;
        // starting guess of the iterative solve (eqgpu_set_warm_start): -1 = the library's default for the mesh size,
        // 0..7 as in include/eqgpu.h (7 = image ring, opt-in)
        int warmStart = -1;
        // true: a step whose PCG stops at max_iters above rtol is reported on stderr and the run continues with the best
        // iterate (what the reference does with its solver's diagnostics); false: stepDiffusion throws
        bool continueOnNoConvergence = false;
        // One large mesh over several GPUs (the reference's precedent is the DMDA split of
        // diffuclass.cpp:364-370): every MPI rank that owns a GPU builds its gpuHSL with the same slabWorld, its own
        // slabRank and the 128-byte id made by gpuHSL::makeSlabId on rank 0 and broadcast by the controller (MPI_Bcast).
        // solution_vector then stays the WHOLE field on every rank: stepDiffusion reads this rank's window of it and
        // writes back the rows slabRows() names; the controller assembles the rest as it does for its own data.
        int slabRank = 0, slabWorld = 1;
        unsigned char slabId[128] = {};
    };
    // rank 0, before constructing the solvers: the id every rank of a slab group must share (eqgpu_nccl_unique_id)
    static void makeSlabId(unsigned char out[128]);
    // the rows [g0, g1) of the mesh this rank owns (the whole mesh without slabs)
    void slabRows(int &g0, int &g1);

    // what simulation.cpp reads through `diffusionSolver->shell->...` (src/simulation.cpp:298-308)
    struct meshShell {
        struct meshInfo {
            size_t n = 0;
            size_t num_vertices() const { return n; }
        };
        std::shared_ptr<meshInfo> mesh = std::make_shared<meshInfo>();
        std::vector<double> mesh_coords;   // 2N doubles, vertex order, x then y
        std::vector<int> dof_from_vertex;  // identity
    };

    gpuHSL() = default;
    explicit gpuHSL(const config &c) : cfg(c) {}
    // the data-recording constructor (src/fHSL.cpp:27-34): the controller uses the same class, with no solver behind
    // it, to write the ABM's data grids over a mesh at nodesPerMicronData (src/simulation.cpp:556-580)
    explicit gpuHSL(const eQ::diffusionSolver::params &recorderParams);
    ~gpuHSL();   // eQ::diffusionSolver has no virtual destructor (src/eQ.h:302-330)

    config cfg;
    eQ::diffusionSolver::params myParams;
    std::shared_ptr<meshShell> shell = std::make_shared<meshShell>();

    // eQ::diffusionSolver
    void initDiffusion(eQ::diffusionSolver::params &) override;  // src/fHSL.cpp:37-53
    void stepDiffusion() override;                               // src/fHSL.cpp:98-161
    eQ::data::parametersType getBoundaryFlux(void) override;     // {"totalFlux": totalBoundaryFlux}
    void writeDiffusionFiles(double timestamp) override;         // src/fHSL.cpp:630-636 (VTK ImageData instead of PVD)
    void finalize(void) override;
    void writeDataFiles(double timestamp);                       // src/fHSL.cpp:637-654 (legacy VTK instead of PVD)

    // fenicsInterface's extra surface
    void setBoundaryValues(const double);    // src/fHSL.cpp:601-604
    void setRobinBoundaryConditions();       // src/fHSL.cpp:331-364
    // what initDiffusion hands eqgpu_create: node counts, spacings, per-wall types/values, channel wiring -- decoded
    // from myParams and cfg exactly as fenicsClassInit / createHSL decode them (src/fHSL.cpp:242-283,436-574)
    void decodeParameters(eqgpu_params &p);
    size_t nodesH = 0, nodesW = 0;
    std::vector<double> solution_vector;     // N doubles; Simulation Isend/Irecv's into it
    std::vector<double> topChannelData, bottomChannelData;
    std::shared_ptr<std::vector<double>> D11, D22, D12;
    double totalBoundaryFlux = 0.0;
    double wellScaling = 0.0;

    // device-resident path (controller-side fusion): eQabm::updateCells' lambdas on the GPU
    void uploadCells(const double *records, size_t ncells);  // EQGPU_CELL_STRIDE doubles each
    void readHSL(double *out);                               // src/abm/eQabm.cpp:326-337, all cells
    void writeHSL(const double *amount_nM);                  // src/abm/eQabm.cpp:338-359, all cells
    // setDiffusionTensor for all uploaded cells (src/abm/eQabm.cpp:246-248,306-325,407) straight into the
    // solver's tensor; Dx, Dy = parameters["AnisotropicDiffusion_Axial" / "_Transverse"].  fetch = also copy
    // the three grids into D11/D22/D12 (what the controller would have sent, src/simulation.cpp:503-505)
    void setDiffusionTensorFromCells(double Dx, double Dy, bool fetch = false);
    void stepDiffusionResident();                            // stepDiffusion without the host round trip
    void fetchSolution();                                    // device field -> solution_vector

    void notifyTensorChanged() { tensorPending = true; }     // D11/D22/D12 were rewritten (see config::tensorScan)
    eqgpu_solver *handle() { return h; }
    int lastIterations() const;

private:
    eqgpu_solver *h = nullptr;
    bool isDataRecordingNode = false;
    bool tensorFromCells = false;
    double leftRate = 0.0, rightRate = 0.0, channelFlowVelocity = 0.0;
    bool tensorDirty = false, tensorPending = false;
    int64_t unconvergedSeen = 0;
    void check(int rc, const char *what);
    void pushTensorIfChanged();
    void reportUnconverged();
};

// The boundary-well model Simulation keeps next to a DIRICHLET_UPDATE layer (src/simulation.cpp:581-627):
// the HSL that leaves the trap accumulates in the flow channels' volume, decays with the flow, and comes
// back as the Dirichlet value of every wall.  Same members and arithmetic as Simulation's; `solver` is any
// class with totalBoundaryFlux and setBoundaryValues(double) (gpuHSL, fenicsInterface).
struct boundaryWell {
    double boundaryWellConcentration = 0.0, boundaryDecayRate = 0.0, wellScaling = 0.0, dt = 0.0;
    long boundaryUnderFlow = 0;
    bool dirichletUpdate = true;   // "DIRICHLET_UPDATE" == parameters["boundaryType"]
    // initBoundaryWell + setBoundaryRate (src/simulation.cpp:607-627)
    void init(double dt_, double lengthScaling, double simulationTrapWidthMicrons, double simulationTrapHeightMicrons,
              double simulationFlowRate)
    {
        dt = dt_;
        boundaryWellConcentration = 0.0;
        boundaryUnderFlow = 0;
        wellScaling = 2.0 * 10.0 * (15.0 / lengthScaling) * (simulationTrapWidthMicrons + simulationTrapHeightMicrons);
        boundaryDecayRate = simulationFlowRate / simulationTrapWidthMicrons;
    }
    // computeBoundaryWell (src/simulation.cpp:581-605)
    template <class Solver>
    void compute(Solver &solver)
    {
        const double flux = solver.totalBoundaryFlux;
        boundaryWellConcentration += flux / wellScaling;
        boundaryWellConcentration -= dt * boundaryDecayRate * boundaryWellConcentration;
        if (boundaryWellConcentration < 0.0) {
            boundaryWellConcentration = 0.0;
            boundaryUnderFlow++;
        }
        if (dirichletUpdate) solver.setBoundaryValues(boundaryWellConcentration);
    }
};
