// Stand-alone mirror of the slice of eQ's src/eQ.h that the HSL solver plugs
// into, so gpuHSL builds and is tested outside the eQ tree (no mpi.h, no
// nlohmann/json here).  Inside eQ, compile with -DEQ_B200_IN_EQ_TREE and the
// real "eQ.h" is used instead; the names below are the reference's own
// (src/eQ.h:302-330) so that gpuHSL is source-compatible with both.
#pragma once
#ifdef EQ_B200_IN_EQ_TREE
#include "eQ.h"
#else
#include <map>
#include <memory>
#include <string>
#include <vector>

typedef int MPI_Comm;  // opaque here; Simulation passes the layer communicator (src/simulation.cpp:657-666)

namespace eQ {
struct data {
    struct record {};
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
    virtual ~diffusionSolver() = default;
    virtual void initDiffusion(eQ::diffusionSolver::params &) = 0;
    virtual void stepDiffusion() {}
    virtual eQ::data::parametersType getBoundaryFlux(void) { return {}; }
    virtual void writeDiffusionFiles(double timestamp) = 0;
    virtual void finalize(void) {}
};
}  // namespace eQ
#endif
