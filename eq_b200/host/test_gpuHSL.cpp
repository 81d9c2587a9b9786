// Driver used by tests/test_host_class.py: exercises gpuHSL exactly the way
// Simulation does (src/simulation.cpp:207,244,466-476,491-505) and dumps the
// fields for comparison with the oracle.
//   test_gpuHSL <case> <in.bin> <out.bin>
// in.bin : doubles [W, H, npm, dt, D, nsteps, ncells, <ncells*16 records>, <N deposits added before each step>]
// Cases "decode" and "golden" take the whole configuration from the file instead (tests/test_host_decode.py,
// tests/test_zz_fenics_golden_gpu.py): after the 7 header doubles (ncells = 0) come
//   [boundaryType code, trapType code, boundaries[4][3], lengthScaling, simulationFlowRate, channel length left, right,
//    channelSolverNumberIterations, hasTensor], then 3N tensor doubles if hasTensor, then per step [wall value or NaN, N field].
// "decode" makes no device call: it writes the parameter block gpuHSL::decodeParameters hands eqgpu_create.
// "golden" writes [nW, nH] and per step [N field, nW top channel, nW bottom channel, totalBoundaryFlux].
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <string>
#include <vector>

#include "gpuFD.h"
#include "gpuHSL.h"

static std::vector<double> read_all(const char *path)
{
    FILE *f = fopen(path, "rb");
    if (!f) { perror(path); exit(2); }
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    std::vector<double> v(n / sizeof(double));
    if (fread(v.data(), sizeof(double), v.size(), f) != v.size()) { perror("read"); exit(2); }
    fclose(f);
    return v;
}

int main(int argc, char **argv)
{
    if (argc < 4) { fprintf(stderr, "usage: %s case in.bin out.bin\n", argv[0]); return 2; }
    const std::string kase = argv[1];
    std::vector<double> in = read_all(argv[2]);
    const double W = in[0], H = in[1], npm = in[2], dt = in[3], D = in[4];
    const int nsteps = int(in[5]);
    const size_t ncells = size_t(in[6]);
    const double *cells = in.data() + 7;
    const double *deposit = cells + ncells * EQGPU_CELL_STRIDE;

    eQ::diffusionSolver::params p{};
    p.dt = dt; p.D_HSL = D; p.trapHeightMicrons = H; p.trapWidthMicrons = W; p.nodesPerMicron = npm;
    p.uniqueID = 0; p.comm = 0;

    if (kase == "fd" || kase == "fd_robin") {
        // diffusionPETSc driven through eQ::diffusionSolver the way src/simulation.cpp:211,244,475 would
        // (the PETSC_SIMULATION branch): initDiffusion, then deposits arrive and stepDiffusion runs
        try {
            std::shared_ptr<gpuFD> solver = std::make_shared<gpuFD>();
            solver->initDiffusion(p);
            if (kase == "fd_robin") {   // "set boundaries explicitly by writing the data structures" (diffuclass.cpp:123)
                solver->initData.leftNeumannCoefficient = 1.0; solver->initData.leftDirichletCoefficient = 0.1;
                solver->initData.leftBoundaryValue = 0.03;
                solver->initData.rightNeumannCoefficient = 1.0; solver->initData.rightDirichletCoefficient = 0.02;
                solver->initData.rightBoundaryValue = 0.0;
                solver->initData.topBoundaryValue = 2.0; solver->initData.bottomBoundaryValue = 0.5;
                solver->applyBoundaryCoefficients();
            }
            const size_t N = solver->solution_vector.size();
            std::vector<double> flux;
            for (int s = 0; s < nsteps; ++s) {
                for (size_t k = 0; k < N; ++k) solver->solution_vector[k] += deposit[k];
                solver->stepDiffusion();
                flux.push_back(solver->getBoundaryFlux()["totalFlux"]);
            }
            FILE *f = fopen(argv[3], "wb");
            double hdr[3] = {double(solver->gridNodesX), double(solver->gridNodesY), double(solver->lastIterations())};
            fwrite(hdr, sizeof(double), 3, f);
            fwrite(solver->solution_vector.data(), sizeof(double), N, f);
            std::vector<double> zero(2 * solver->gridNodesX, 0.0);
            fwrite(zero.data(), sizeof(double), zero.size(), f);
            fwrite(flux.data(), sizeof(double), flux.size(), f);
            fclose(f);
            solver->finalize();
        } catch (const std::exception &e) {
            fprintf(stderr, "%s\n", e.what());
            return 1;
        }
        return 0;
    }

    if (kase == "recorder") {
        // the controller's data-recording use of the class (src/simulation.cpp:556-580): no solver, no device call.
        // One scalar and one vector grid with recognisable values; argv[3] is the file-path prefix.
        const size_t n = size_t(npm), y = size_t(H), x = size_t(W);
        auto sc = std::make_shared<eQ::data::tensor>(n, y, x, eQ::data::tensor::rank::SCALAR);
        auto ve = std::make_shared<eQ::data::tensor>(n, y, x, eQ::data::tensor::rank::VECTOR);
        for (size_t i = 0; i < sc->nh; ++i)
            for (size_t j = 0; j < sc->nw; ++j) {
                sc->grid[0][i * sc->nw + j] = 1000.0 * double(i) + double(j) + 1.0 / 3.0;
                ve->grid[0][i * ve->nw + j] = double(i);
                ve->grid[1][i * ve->nw + j] = -double(j);
            }
        p.dataFiles = std::make_shared<eQ::data::files_t>();
        p.dataFiles->push_back(eQ::data::record{0, 0, "scalarGrid", sc});
        p.dataFiles->push_back(eQ::data::record{1, 1, "vectorGrid", ve});
        p.filePath = argv[3];
        gpuHSL recorder(p);
        recorder.writeDataFiles(in[5]);
        printf("%zu %zu\n", recorder.nodesW, recorder.nodesH);
        return 0;
    }
    if (kase == "decode" || kase == "golden") {
        static const char *btypes[] = {"DIRICHLET_0", "DIRICHLET_UPDATE", "MICROFLUIDIC_TRAP", "NEUMANN_3WALLED_TEST", "SOMETHING_ELSE"};
        static const char *ttypes[] = {"NOWALLED", "THREEWALLED", "TWOWALLED", "ONEWALLED", "H_TRAP"};
        const double *c = in.data() + 7;
        gpuHSL::config cfg;
        cfg.boundaryType = btypes[int(c[0])];
        cfg.trapType = ttypes[int(c[1])];
        for (int w = 0; w < 4; ++w) for (int k = 0; k < 3; ++k) cfg.boundaries[w][k] = c[2 + 3 * w + k];
        cfg.lengthScaling = c[14]; cfg.simulationFlowRate = c[15];
        cfg.simulationChannelLengthLeft = c[16]; cfg.simulationChannelLengthRight = c[17];
        cfg.channelSolverNumberIterations = int(c[18]);
        const bool hasTensor = c[19] != 0.0;
        const double *rest = c + 20;
        std::shared_ptr<gpuHSL> solver = std::make_shared<gpuHSL>(cfg);
        FILE *f = nullptr;
        try {
            if (kase == "decode") {
                solver->myParams = p;
                eqgpu_params q;
                solver->decodeParameters(q);
                const double out[] = {double(q.nW), double(q.nH), q.hx, q.hy, q.dt, q.D,
                                      double(q.bc_type[0]), double(q.bc_type[1]), double(q.bc_type[2]), double(q.bc_type[3]),
                                      q.bc_value[0], q.bc_value[1], q.bc_value[2], q.bc_value[3],
                                      double(q.channels), double(q.channel_iters), q.channel_v, q.channel_r[0], q.channel_r[1],
                                      q.well_scaling, q.robin_s[0], q.robin_s[1]};
                f = fopen(argv[3], "wb");
                fwrite(out, sizeof(double), sizeof out / sizeof out[0], f);
                fclose(f);
                return 0;
            }
            solver->initDiffusion(p);
            const size_t N = solver->solution_vector.size();
            if (hasTensor) {   // what the controller's MPI transfer would have written (src/simulation.cpp:503-505)
                std::copy(rest, rest + N, solver->D11->begin());
                std::copy(rest + N, rest + 2 * N, solver->D22->begin());
                std::copy(rest + 2 * N, rest + 3 * N, solver->D12->begin());
                rest += 3 * N;
            }
            f = fopen(argv[3], "wb");
            double hdr[2] = {double(solver->nodesW), double(solver->nodesH)};
            fwrite(hdr, sizeof(double), 2, f);
            for (int s = 0; s < nsteps; ++s) {
                if (rest[0] == rest[0]) solver->setBoundaryValues(rest[0]);   // not NaN: src/simulation.cpp:602
                std::copy(rest + 1, rest + 1 + N, solver->solution_vector.begin());
                rest += 1 + N;
                solver->stepDiffusion();
                fwrite(solver->solution_vector.data(), sizeof(double), N, f);
                fwrite(solver->topChannelData.data(), sizeof(double), solver->nodesW, f);
                fwrite(solver->bottomChannelData.data(), sizeof(double), solver->nodesW, f);
                fwrite(&solver->totalBoundaryFlux, sizeof(double), 1, f);
            }
            fclose(f);
            solver->finalize();
        } catch (const std::exception &e) {
            fprintf(stderr, "%s\n", e.what());
            return 1;
        }
        return 0;
    }

    gpuHSL::config cfg;
    bool well = false;
    if (kase == "well") { cfg.boundaryType = "DIRICHLET_UPDATE"; cfg.trapType = "NOWALLED"; well = true; }
    else if (kase == "default") { cfg.boundaryType = "DIRICHLET_0"; cfg.trapType = "NOWALLED"; }
    else if (kase == "threewall") { cfg.boundaryType = "DIRICHLET_0"; cfg.trapType = "THREEWALLED"; }
    else if (kase == "htrap") {  // MICROFLUIDIC_TRAP + H_TRAP: Robin left/right = flow rate, top/bottom Neumann
        cfg.boundaryType = "MICROFLUIDIC_TRAP"; cfg.trapType = "H_TRAP";
        const double rob[3] = {1.0, 1.0, 0.0}, neu[3] = {1.0, 0.0, 0.0};
        memcpy(cfg.boundaries[0], rob, sizeof rob); memcpy(cfg.boundaries[1], rob, sizeof rob);
        memcpy(cfg.boundaries[2], neu, sizeof neu); memcpy(cfg.boundaries[3], neu, sizeof neu);
    } else if (kase == "channels") {  // MICROFLUIDIC_TRAP, default trap: Robin left/right, channel Dirichlet top/bottom
        cfg.boundaryType = "MICROFLUIDIC_TRAP"; cfg.trapType = "NOWALLED";
        const double rob[3] = {1.0, 1.0, 0.0}, chan[3] = {0.0, 1.0, -1.0};
        memcpy(cfg.boundaries[0], rob, sizeof rob); memcpy(cfg.boundaries[1], rob, sizeof rob);
        memcpy(cfg.boundaries[2], chan, sizeof chan); memcpy(cfg.boundaries[3], chan, sizeof chan);
    } else { fprintf(stderr, "unknown case\n"); return 2; }

    // EQ_TEST_SLAB=<rank>,<world>,<id file>: this process is one rank of a row-slab group (gpuHSL::config::slabRank ...;
    // INTEGRATION 4d).  Rank 0 makes the id and leaves it in the file -- the stand-in for the controller's MPI_Bcast.
    if (const char *sl = getenv("EQ_TEST_SLAB")) {
        int rank = 0, world = 1;
        char path[512] = {0};
        if (sscanf(sl, "%d,%d,%511s", &rank, &world, path) != 3) { fprintf(stderr, "bad EQ_TEST_SLAB\n"); return 2; }
        cfg.slabRank = rank; cfg.slabWorld = world; cfg.device = rank;
        if (rank == 0) {
            try { gpuHSL::makeSlabId(cfg.slabId); } catch (const std::exception &e) { fprintf(stderr, "%s\n", e.what()); return 1; }
            const std::string tmp = std::string(path) + ".tmp";
            FILE *f = fopen(tmp.c_str(), "wb");
            if (!f || fwrite(cfg.slabId, 1, 128, f) != 128) { perror(tmp.c_str()); return 2; }
            fclose(f);
            rename(tmp.c_str(), path);
        } else {
            FILE *f = nullptr;
            for (int tries = 0; tries < 600 && !(f = fopen(path, "rb")); ++tries) { struct timespec ts = {0, 100000000}; nanosleep(&ts, nullptr); }
            if (!f || fread(cfg.slabId, 1, 128, f) != 128) { fprintf(stderr, "no slab id in %s\n", path); return 2; }
            fclose(f);
        }
    }
    std::shared_ptr<gpuHSL> solver = std::make_shared<gpuHSL>(cfg);  // simulation.cpp:207
    // EQ_TEST_SNAPSHOT=<prefix>: the controller's field snapshot (src/main.cpp:148-162 -> writeDiffusionFiles,
    // src/fHSL.cpp:630-636) after the last step, for the read-back test
    const char *snap = getenv("EQ_TEST_SNAPSHOT");
    if (snap) p.filePath = snap;
    try {
        solver->initDiffusion(p);                                    // simulation.cpp:244
        const size_t N = solver->solution_vector.size();
        if (solver->shell->mesh->num_vertices() != N) return 3;      // simulation.cpp:298-301
        std::vector<double> flux;
        boundaryWell bw;
        if (well) bw.init(dt, cfg.lengthScaling, W, H, cfg.simulationFlowRate);   // simulation.cpp:607-627
        for (int s = 0; s < nsteps; ++s) {
            for (size_t k = 0; k < N; ++k) solver->solution_vector[k] += deposit[k];  // controller's writeHSL result arrives
            solver->stepDiffusion();                                 // simulation.cpp:475
            if (well) {                                              // simulation.cpp:581-605 (after every HSL step)
                bw.compute(*solver);
                flux.push_back(bw.boundaryWellConcentration);
            } else
                flux.push_back(solver->getBoundaryFlux()["totalFlux"]);
        }
        if (snap) solver->writeDiffusionFiles(double(nsteps) * dt);
        FILE *f = fopen(argv[3], "wb");
        double hdr[3] = {double(solver->nodesW), double(solver->nodesH), double(solver->lastIterations())};
        fwrite(hdr, sizeof(double), 3, f);
        fwrite(solver->solution_vector.data(), sizeof(double), N, f);
        fwrite(solver->topChannelData.data(), sizeof(double), solver->nodesW, f);
        fwrite(solver->bottomChannelData.data(), sizeof(double), solver->nodesW, f);
        fwrite(flux.data(), sizeof(double), flux.size(), f);
        // fused path: cells on the GPU
        if (ncells) {
            std::vector<double> g(ncells), amt(ncells, 100.0);
            solver->uploadCells(cells, ncells);
            solver->readHSL(g.data());
            solver->writeHSL(amt.data());
            solver->stepDiffusionResident();
            solver->fetchSolution();
            fwrite(g.data(), sizeof(double), ncells, f);
            fwrite(solver->solution_vector.data(), sizeof(double), N, f);
        }
        fclose(f);
        solver->finalize();
    } catch (const std::exception &e) {
        fprintf(stderr, "%s\n", e.what());
        return 1;
    }
    return 0;
}
