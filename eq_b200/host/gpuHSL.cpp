#include "gpuHSL.h"

#include <cmath>
#include <cstdio>
#include <fstream>
#include <iostream>

void gpuHSL::check(int rc, const char *what)
{
    if (rc == EQGPU_OK) return;
    // fenicsInterface has no error convention (failures print to stdout and go on);
    // a GPU solver that silently went on would hand back garbage, so this one throws.
    throw std::runtime_error(std::string("gpuHSL: ") + what + " failed: " + eqgpu_last_error(h));
}

gpuHSL::~gpuHSL() { finalize(); }

// src/fHSL.cpp:27-34 + the data-recording branch of fenicsClassInit (:223-234,281-283): a vertex grid over the trap
// at the data resolution -- note the reference rounds the trap size up BEFORE scaling here, unlike the solver branch
gpuHSL::gpuHSL(const eQ::diffusionSolver::params &recorderParams) : myParams(recorderParams), isDataRecordingNode(true)
{
    nodesH = unsigned(ceil(myParams.trapHeightMicrons) * myParams.nodesPerMicron) + 1;
    nodesW = unsigned(ceil(myParams.trapWidthMicrons) * myParams.nodesPerMicron) + 1;
    shell->mesh->n = nodesH * nodesW;
}

// src/fHSL.cpp:637-654: every registered data grid is sampled at the mesh vertices (scalarDataExpression /
// vectorDataExpression, src/Expressions.h:45-73 -> eQ::data::tensor::eval, nearest node) and written out with the
// time stamp; one legacy-VTK file per grid and time instead of dolfin::File's PVD series.
void gpuHSL::writeDataFiles(double timestamp)
{
    if (!myParams.dataFiles) return;
    const double hx = myParams.trapWidthMicrons / double(nodesW - 1), hy = myParams.trapHeightMicrons / double(nodesH - 1);
    for (auto &file : *myParams.dataFiles) {
        if (!file.data) continue;
        char name[768];
        snprintf(name, sizeof name, "%s%s_%012.4f.vtk", myParams.filePath.c_str(), file.fileName.c_str(), timestamp);
        std::ofstream f(name);
        if (!f) continue;
        f.precision(17);
        const bool vec = (eQ::data::tensor::rank::VECTOR == file.data->getRank());
        f << "# vtk DataFile Version 3.0\n" << file.fileName << " t=" << timestamp << "\nASCII\nDATASET STRUCTURED_POINTS\n";
        f << "DIMENSIONS " << nodesW << " " << nodesH << " 1\nORIGIN 0 0 0\nSPACING " << hx << " " << hy << " 1\n";
        f << "POINT_DATA " << nodesW * nodesH << "\n";
        if (vec) f << "VECTORS data double\n";
        else f << "SCALARS data double 1\nLOOKUP_TABLE default\n";
        for (size_t i = 0; i < nodesH; ++i)
            for (size_t j = 0; j < nodesW; ++j) {
                const double x = double(j) * hx, y = double(i) * hy;
                if (vec) {
                    const auto v = file.data->evalVector(x, y);
                    f << v.first << " " << v.second << " 0\n";
                } else
                    f << file.data->eval(x, y) << "\n";
            }
    }
}

void gpuHSL::finalize()
{
    if (h) { eqgpu_destroy(h); h = nullptr; }
}

// src/fHSL.cpp:331-364
void gpuHSL::setRobinBoundaryConditions()
{
    channelFlowVelocity = cfg.simulationFlowRate;
    const double D = myParams.D_HSL;
    const double lvdl = (cfg.simulationChannelLengthLeft * channelFlowVelocity) / D;
    const double lvdr = (cfg.simulationChannelLengthRight * channelFlowVelocity) / D;
    if (channelFlowVelocity > 1.0e-6) {
        leftRate = channelFlowVelocity * (1.0 / (1.0 - exp(-lvdl)));
        rightRate = channelFlowVelocity * (1.0 / (exp(lvdr) - 1.0));
    } else {
        leftRate = D / cfg.simulationChannelLengthLeft;
        rightRate = D / cfg.simulationChannelLengthRight;
    }
}

// src/fHSL.cpp:37-53 + fenicsClassInit (:195-328) + createHSL's boundary decode (:436-574), host arithmetic only:
// fills the C-ABI parameter block from myParams and cfg (no device call, so the decoding is testable without a GPU
// against what the reference class itself binds into its forms, tests/test_host_decode.py)
void gpuHSL::decodeParameters(eqgpu_params &p)
{
    // cell counts, then +1 for node counts (src/fHSL.cpp:242-243,281-283)
    nodesH = unsigned(ceil(myParams.trapHeightMicrons * myParams.nodesPerMicron)) + 1;
    nodesW = unsigned(ceil(myParams.trapWidthMicrons * myParams.nodesPerMicron)) + 1;
    const double h_sim = 1.0 / myParams.nodesPerMicron;
    wellScaling = 10.0 * (25.0 / cfg.lengthScaling) * h_sim;  // src/fHSL.cpp:47
    setRobinBoundaryConditions();

    eqgpu_default_params(&p);
    p.nW = int(nodesW);
    p.nH = int(nodesH);
    // DOLFIN spreads the vertices evenly over [0,W]x[0,H] (src/fHSL.cpp:164-172)
    p.hx = myParams.trapWidthMicrons / double(nodesW - 1);
    p.hy = myParams.trapHeightMicrons / double(nodesH - 1);
    p.dt = myParams.dt;
    p.D = myParams.D_HSL;
    p.device = cfg.device;
    p.rtol = cfg.rtol;
    p.channels = 0;
    for (int w = 0; w < 4; ++w) { p.bc_type[w] = EQGPU_BC_NEUMANN; p.bc_value[w] = 0.0; }
    if (cfg.boundaryType == "MICROFLUIDIC_TRAP") {
        if (cfg.trapType == "H_TRAP") { leftRate = channelFlowVelocity; rightRate = channelFlowVelocity; }  // :448-453
        const double rates[2] = {leftRate, rightRate};
        for (int w = 0; w < 4; ++w) {
            const double *d = cfg.boundaries[w];
            if (d[0] == 0.0) {  // Dirichlet (:471,490,508,525)
                if ((w == EQGPU_TOP || w == EQGPU_BOTTOM) && d[2] == -1.0) p.bc_type[w] = EQGPU_BC_DIRICHLET_CHANNEL;
                else { p.bc_type[w] = EQGPU_BC_DIRICHLET; p.bc_value[w] = d[2]; }
            } else if (d[1] == 0.0) {
                p.bc_type[w] = EQGPU_BC_NEUMANN;  // homogeneous only (:476-481)
            } else if (w == EQGPU_LEFT || w == EQGPU_RIGHT) {
                p.bc_type[w] = EQGPU_BC_ROBIN;  // rate computed above (:483-486)
                p.bc_value[w] = rates[w];
            }
        }
        p.channels = (cfg.trapType != "H_TRAP") ? 1 : 0;  // src/fHSL.cpp:110
        p.channel_iters = cfg.channelSolverNumberIterations;
        p.channel_v = channelFlowVelocity;
        p.channel_r[0] = leftRate;
        p.channel_r[1] = rightRate;
        p.well_scaling = wellScaling;
    } else if (cfg.boundaryType == "DIRICHLET_UPDATE") {  // :545-555
        for (int w = 0; w < 4; ++w) p.bc_type[w] = EQGPU_BC_DIRICHLET;
    } else if (cfg.boundaryType == "DIRICHLET_0") {  // :557-569 with the SubDomains of src/fHSL.h:456-504
        bool dir[4] = {false, false, false, false};  // left,right,top,bottom
        if (cfg.trapType == "NOWALLED") dir[0] = dir[1] = dir[2] = dir[3] = true;
        if (cfg.trapType == "THREEWALLED") dir[3] = true;
        if (cfg.trapType == "TWOWALLED") dir[2] = dir[3] = true;
        if (cfg.trapType == "ONEWALLED") dir[0] = dir[1] = dir[3] = true;
        for (int w = 0; w < 4; ++w) if (dir[w]) p.bc_type[w] = EQGPU_BC_DIRICHLET;
    } else if (cfg.boundaryType == "NEUMANN_3WALLED_TEST") {  // :570-571 uses dbc_oneWall
        p.bc_type[EQGPU_LEFT] = p.bc_type[EQGPU_RIGHT] = p.bc_type[EQGPU_BOTTOM] = EQGPU_BC_DIRICHLET;
    } else {  // :572-573
        for (int w = 0; w < 4; ++w) p.bc_type[w] = EQGPU_BC_DIRICHLET;
    }
}

void gpuHSL::initDiffusion(eQ::diffusionSolver::params &initParams)
{
    myParams = initParams;
    eqgpu_params p;
    decodeParameters(p);
    int rc = cfg.slabWorld > 1 ? eqgpu_create_slab(&p, cfg.slabRank, cfg.slabWorld, cfg.slabId, &h) : eqgpu_create(&p, &h);
    if (rc != EQGPU_OK) throw std::runtime_error(std::string("gpuHSL: eqgpu_create failed: ") + eqgpu_last_error(nullptr));
    if (cfg.warmStart >= 0) check(eqgpu_set_warm_start(h, cfg.warmStart), "eqgpu_set_warm_start");
    if (cfg.continueOnNoConvergence) check(eqgpu_set_nonconvergence_policy(h, 1), "eqgpu_set_nonconvergence_policy");

    const size_t N = nodesH * nodesW;
    solution_vector.assign(N, 0.0);  // src/fHSL.cpp:583-584
    topChannelData.assign(nodesW, 0.0);
    bottomChannelData.assign(nodesW, 0.0);
    D11 = std::make_shared<std::vector<double>>(N, 1.0);  // src/fHSL.cpp:313-323
    D22 = std::make_shared<std::vector<double>>(N, 1.0);
    D12 = std::make_shared<std::vector<double>>(N, 0.0);
    // vertex coordinates and the (identity) vertex->dof map that Simulation::create_HSLgrid
    // exchanges and turns into its lookup table (src/simulation.cpp:298-313,363-387)
    shell->mesh->n = N;
    shell->mesh_coords.resize(2 * N);
    shell->dof_from_vertex.resize(N);
    for (size_t i = 0; i < nodesH; ++i)
        for (size_t j = 0; j < nodesW; ++j) {
            const size_t v = i * nodesW + j;
            shell->mesh_coords[2 * v] = double(j) * p.hx;
            shell->mesh_coords[2 * v + 1] = double(i) * p.hy;
            shell->dof_from_vertex[v] = int(v);
        }
}

void gpuHSL::makeSlabId(unsigned char out[128])
{
    if (eqgpu_nccl_unique_id(out) != EQGPU_OK)
        throw std::runtime_error("gpuHSL::makeSlabId: eqgpu_nccl_unique_id failed (is libnccl.so.2 loadable?)");
}

void gpuHSL::slabRows(int &g0, int &g1)
{
    int32_t a = 0, b = 0;
    check(eqgpu_slab_rows(h, &a, &b), "eqgpu_slab_rows");
    g0 = a; g1 = b;
}

void gpuHSL::pushTensorIfChanged()
{
    // The controller sends D11/D22/D12 every step (src/simulation.cpp:503-505) but as shipped they
    // stay 1,1,0 (SURVEY.md finding 5): only a non-trivial tensor switches the variable-tensor operator on.
    if (tensorFromCells) return;   // the tensor was rasterised on the device (setDiffusionTensorFromCells)
    const size_t N = solution_vector.size();
    // Looking for a change means reading 3N doubles on the host: nothing at the shipped 201 x 41 nodes, ten times the
    // GPU's whole step at 2048^2.  Large meshes therefore wait to be told (notifyTensorChanged) unless cfg says scan.
    const bool scan = cfg.tensorScan > 0 || (cfg.tensorScan < 0 && N <= (size_t(1) << 18));
    if (!scan && !tensorPending) return;
    tensorPending = false;
    bool iso = true;
    for (size_t k = 0; k < N && iso; ++k)
        iso = ((*D11)[k] == 1.0 && (*D22)[k] == 1.0 && (*D12)[k] == 0.0);
    if (!iso) {
        check(eqgpu_set_tensor(h, D11->data(), D22->data(), D12->data()), "eqgpu_set_tensor");
        tensorDirty = true;
    } else if (tensorDirty) {
        check(eqgpu_set_tensor(h, nullptr, nullptr, nullptr), "eqgpu_set_tensor");
        tensorDirty = false;
    }
}

// src/fHSL.cpp:98-161: host vector in, solved field out
// (continueOnNoConvergence) a step that did not reach rtol: say so once per step and go on, as the reference does
void gpuHSL::reportUnconverged()
{
    int64_t n = 0;
    if (eqgpu_unconverged_steps(h, &n) == EQGPU_OK && n > unconvergedSeen) {
        unconvergedSeen = n;
        std::cerr << "gpuHSL: " << eqgpu_last_error(h) << " -- continuing with the best iterate (" << n
                  << " such steps so far)" << std::endl;
    }
}

void gpuHSL::stepDiffusion()
{
    pushTensorIfChanged();
    check(eqgpu_step_host(h, solution_vector.data()), "eqgpu_step_host");
    reportUnconverged();
    eqgpu_stats st;
    check(eqgpu_get_stats(h, &st), "eqgpu_get_stats");
    totalBoundaryFlux = st.total_boundary_flux;
    check(eqgpu_get_channels(h, topChannelData.data(), bottomChannelData.data()), "eqgpu_get_channels");
}

void gpuHSL::stepDiffusionResident()
{
    check(eqgpu_step(h), "eqgpu_step");
    reportUnconverged();
    eqgpu_stats st;
    check(eqgpu_get_stats(h, &st), "eqgpu_get_stats");
    totalBoundaryFlux = st.total_boundary_flux;
}

void gpuHSL::fetchSolution() { check(eqgpu_get_field(h, solution_vector.data()), "eqgpu_get_field"); }

int gpuHSL::lastIterations() const
{
    eqgpu_stats st;
    if (eqgpu_get_stats(h, &st) != EQGPU_OK) return -1;
    return st.iterations;
}

// The base-class version upstream has no return statement although Simulation calls it
// (src/eQ.h:327, src/simulation.cpp:545 -- SURVEY.md appendix B.1); this one returns the value.
eQ::data::parametersType gpuHSL::getBoundaryFlux(void)
{
    eQ::data::parametersType j;
    j["totalFlux"] = totalBoundaryFlux;
    return j;
}

void gpuHSL::setBoundaryValues(const double v) { check(eqgpu_set_boundary_value(h, v), "eqgpu_set_boundary_value"); }

void gpuHSL::uploadCells(const double *records, size_t n)
{
    check(eqgpu_cells_upload(h, records, (int64_t)n, myParams.nodesPerMicron), "eqgpu_cells_upload");
}
void gpuHSL::setDiffusionTensorFromCells(double Dx, double Dy, bool fetch)
{
    check(eqgpu_cells_tensor(h, Dx, Dy), "eqgpu_cells_tensor");
    tensorFromCells = true;
    if (fetch) check(eqgpu_get_tensor(h, D11->data(), D22->data(), D12->data()), "eqgpu_get_tensor");
}
void gpuHSL::readHSL(double *out) { check(eqgpu_cells_gather(h, out), "eqgpu_cells_gather"); }
void gpuHSL::writeHSL(const double *amount) { check(eqgpu_cells_scatter(h, amount), "eqgpu_cells_scatter"); }

// src/fHSL.cpp:630-636 writes compressed PVD through dolfin::File; without DOLFIN the same
// snapshot goes out as legacy-VTK structured points (ParaView reads both).
void gpuHSL::writeDiffusionFiles(double timestamp)
{
    if (myParams.filePath.empty()) return;
    char name[512];
    snprintf(name, sizeof name, "%s_%012.4f.vtk", myParams.filePath.c_str(), timestamp);
    std::ofstream f(name);
    if (!f) return;
    f.precision(17);   // round-trip doubles
    const double hx = myParams.trapWidthMicrons / double(nodesW - 1), hy = myParams.trapHeightMicrons / double(nodesH - 1);
    f << "# vtk DataFile Version 3.0\nHSL t=" << timestamp << "\nASCII\nDATASET STRUCTURED_POINTS\n";
    f << "DIMENSIONS " << nodesW << " " << nodesH << " 1\nORIGIN 0 0 0\nSPACING " << hx << " " << hy << " 1\n";
    f << "POINT_DATA " << nodesW * nodesH << "\nSCALARS u double 1\nLOOKUP_TABLE default\n";
    for (double v : solution_vector) f << v << "\n";
}
