// Stand-alone mirror of the slice of eQ's src/eQ.h that the HSL solver plugs
// into, so gpuHSL builds and is tested outside the eQ tree (no mpi.h, no
// nlohmann/json here).  Inside eQ, compile with -DEQ_B200_IN_EQ_TREE and the
// real "eQ.h" is used instead; the names below are the reference's own
// (src/eQ.h:302-330) so that gpuHSL is source-compatible with both.
#pragma once
#ifdef EQ_B200_IN_EQ_TREE
#include "eQ.h"
#else
#include <cmath>
#include <map>
#include <memory>
#include <string>
#include <vector>

typedef int MPI_Comm;  // opaque here; Simulation passes the layer communicator (src/simulation.cpp:657-666)

namespace eQ {
struct data {
    // src/eQ.h:136-180: what a data-recording grid exposes to the writer (nearest-node lookup at the
    // data resolution); only the members gpuHSL::writeDataFiles reads
    class tensor {
    public:
        enum rank { SCALAR = 0, VECTOR = 1, NUM_RANKS };
        tensor(size_t n, size_t y, size_t x, size_t thisRank) : nh(y * n + 1), nw(x * n + 1), nodesPerMicron(double(n)), rank_(thisRank)
        {
            for (size_t i = 0; i <= rank_; ++i) grid.push_back(std::vector<double>(nh * nw, 0.0));
        }
        size_t getRank() { return rank_; }
        size_t index(double x, double y) const { return size_t(round(y * nodesPerMicron)) * nw + size_t(round(x * nodesPerMicron)); }
        double eval(double x, double y) { return grid[0][index(x, y)]; }
        std::pair<double, double> evalVector(double x, double y) { return std::make_pair(grid[0][index(x, y)], grid[1][index(x, y)]); }
        std::vector<std::vector<double>> grid;   // row-major nh x nw per component
        size_t nh, nw;
    private:
        double nodesPerMicron;
        size_t rank_;
    };
    struct record {   // src/eQ.h:95-101
        int type;
        size_t index;
        std::string fileName;
        std::shared_ptr<tensor> data;
    };
    using files_t = std::vector<record>;
    // the reference returns nlohmann::json from getBoundaryFlux(); {"totalFlux": v} is all it carries
    using parametersType = std::map<std::string, double>;
};

// src/eQ.h:302-330
class diffusionSolver {
public:
    struct params {
        int argc;
        char **argv;
        size_t uniqueID;
        MPI_Comm comm;
        double dt;
        double D_HSL;
        std::string filePath;
        std::string filePathTopChannel;
        std::string filePathBottomChannel;
        std::shared_ptr<eQ::data::files_t> dataFiles;
        double trapHeightMicrons;
        double trapWidthMicrons;
        double nodesPerMicron;
        double trapChannelVelocity;
    };
    // (no virtual destructor upstream either: Simulation holds the solver by its concrete type)
    virtual void initDiffusion(eQ::diffusionSolver::params &) = 0;
    virtual void stepDiffusion() {}
    virtual eQ::data::parametersType getBoundaryFlux(void) { return {}; }
    virtual void writeDiffusionFiles(double timestamp) = 0;
    virtual void finalize(void) {}
};
}  // namespace eQ
#endif
