// Builds oracle/_ref/libeq_fenics_ref.so: the reference's OWN class `fenicsInterface` (src/fHSL.{h,cpp}, with
// src/Expressions.h and the FFC-generated form headers fenics/*.h, qtensor/*.h) compiled in place from /root/reference
// (no copy of the sources enters this repository) on the one-process DOLFIN interface shim of oracle/shim_dolfin/
// (DOLFIN 2019.1.0 is not vendored by the reference; the shim restates mesh numbering, assembly, DirichletBC::apply
// and a banded LU), the UFC base classes of oracle/shim/, the Chipmunk and MPI stand-ins of shim_cpm/, shim_petsc/.
// TEST INFRASTRUCTURE ONLY: it pins the oracle's restatement of the P1 step (oracle.step: assemble, Dirichlet rows,
// solve, computeBoundaryFlux, channel sub-steps, flux functional; the boundary decoding of createHSL and the Robin
// rates of setRobinBoundaryConditions) and generates tests/golden/fenics_ref.json.
//
// What runs unmodified: fenicsInterface::initDiffusion / fenicsClassInit / createMesh / createHSL /
// setRobinBoundaryConditions / stepDiffusion / computeBoundaryFlux / setBoundaryValues, fenicsBaseClass and
// fenicsChannel (form wiring, dof lookup tables), the SubDomain classes of src/fHSL.h, AnisotropicDiffusionTensor
// and updatingDirchletBoundary of src/Expressions.h, and every tabulate_tensor of hslD / AdvectionDiffusion /
// boundary.
#include "fHSL.h"   // /root/reference/src/fHSL.h (-I$(REF)/src)

#include <cstring>
#include <sstream>

// src/main.cpp:44 defines this static in the executable; the parity pin is not linked against main.cpp
eQ::data::parametersType eQ::data::parameters;
namespace dolfin { Parameters parameters; }
int PETSC_COMM_WORLD = 0;

#define REF_API extern "C" __attribute__((visibility("default")))

struct FenicsRef {
    std::shared_ptr<fenicsInterface> f;
    std::string err;
};

static thread_local std::string g_err;
REF_API const char *ref_fenics_last_error() { return g_err.c_str(); }

// params_json: the keys of eQ::data::parameters the path reads (SURVEY 8b "Globals read by the solver"), merged
// into the global; the numeric members of eQ::diffusionSolver::params are passed explicitly.
REF_API void *ref_fenics_create(const char *params_json, double dt, double D, double widthMicrons, double heightMicrons,
                                double npm, double channelVelocity)
{
    try {
        std::streambuf *keep = std::cout.rdbuf();
        std::ostringstream sink;
        std::cout.rdbuf(sink.rdbuf());          // the reference narrates its set-up on stdout
        auto j = eQ::data::parametersType::parse(params_json);
        eQ::data::parameters = j;
        eQ::diffusionSolver::params p;
        p.argc = 0; p.argv = nullptr; p.uniqueID = 0; p.comm = 0;
        p.dt = dt; p.D_HSL = D;
        p.filePath = "/dev/null"; p.filePathTopChannel = "/dev/null"; p.filePathBottomChannel = "/dev/null";
        p.dataFiles = std::make_shared<eQ::data::files_t>();
        p.trapHeightMicrons = heightMicrons; p.trapWidthMicrons = widthMicrons;
        p.nodesPerMicron = npm; p.trapChannelVelocity = channelVelocity;
        FenicsRef *r = new FenicsRef();
        r->f = std::make_shared<fenicsInterface>();
        r->f->initDiffusion(p);
        std::cout.rdbuf(keep);
        return r;
    } catch (const std::exception &e) {
        g_err = e.what();
        return nullptr;
    }
}

REF_API void ref_fenics_destroy(void *h) { delete (FenicsRef *)h; }

REF_API void ref_fenics_sizes(void *h, long *nW, long *nH, long *nChannel)
{
    fenicsInterface &f = *((FenicsRef *)h)->f;
    *nW = (long)f.nodesW; *nH = (long)f.nodesH; *nChannel = (long)f.solution_vectorTopChannel.size();
}

// what Simulation::create_HSLgrid copies out (src/simulation.cpp:298-308): 2N coordinates, N dofs
REF_API void ref_fenics_mesh(void *h, double *coords, int *dof_from_vertex)
{
    fenicsInterface &f = *((FenicsRef *)h)->f;
    std::memcpy(coords, f.shell->mesh_coords.data(), f.shell->mesh_coords.size() * sizeof(double));
    std::memcpy(dof_from_vertex, f.shell->dof_from_vertex.data(), f.shell->dof_from_vertex.size() * sizeof(int));
}

// the (iy, jx) -> dof table of fenicsShell::createGridCoordinatesToDofMapping (src/fHSL.h:89-114), row-major
REF_API void ref_fenics_lookup(void *h, long *table)
{
    fenicsInterface &f = *((FenicsRef *)h)->f;
    for (size_t i = 0; i < f.nodesH; ++i)
        for (size_t j = 0; j < f.nodesW; ++j) {
            eQ::nodeType jj = j;
            std::pair<eQ::nodeType, eQ::nodeType &> pt{i, jj};
            table[i * f.nodesW + j] = (long)f.shell->dofLookupTable->operator[](pt);
        }
}

REF_API void ref_fenics_set_field(void *h, const double *u)
{
    fenicsInterface &f = *((FenicsRef *)h)->f;
    std::memcpy(f.solution_vector.data(), u, f.solution_vector.size() * sizeof(double));
}
REF_API void ref_fenics_get_field(void *h, double *u)
{
    fenicsInterface &f = *((FenicsRef *)h)->f;
    std::memcpy(u, f.solution_vector.data(), f.solution_vector.size() * sizeof(double));
}
REF_API void ref_fenics_set_tensor(void *h, const double *d11, const double *d22, const double *d12)
{
    fenicsInterface &f = *((FenicsRef *)h)->f;
    const size_t n = f.solution_vector.size();
    std::memcpy(f.D11->data(), d11, n * sizeof(double));
    std::memcpy(f.D22->data(), d22, n * sizeof(double));
    std::memcpy(f.D12->data(), d12, n * sizeof(double));
}
REF_API void ref_fenics_set_channels(void *h, const double *top, const double *bottom)
{
    fenicsInterface &f = *((FenicsRef *)h)->f;
    const size_t n = f.solution_vectorTopChannel.size();
    std::memcpy(f.solution_vectorTopChannel.data(), top, n * sizeof(double));
    std::memcpy(f.solution_vectorBottomChannel.data(), bottom, n * sizeof(double));
    // the trap's DirichletBC reads the channel Functions (src/fHSL.cpp:511-515), which stepDiffusion refreshes only
    // after the trap solve; seed them too so a non-zero start is seen by the first step
    f.topChannel->u->vector()->set_local(f.solution_vectorTopChannel);
    f.bottomChannel->u->vector()->set_local(f.solution_vectorBottomChannel);
}
REF_API void ref_fenics_get_channels(void *h, double *top, double *bottom, double *fluxTop, double *fluxBottom)
{
    fenicsInterface &f = *((FenicsRef *)h)->f;
    std::memcpy(top, f.topChannelData.data(), f.topChannelData.size() * sizeof(double));
    std::memcpy(bottom, f.bottomChannelData.data(), f.bottomChannelData.size() * sizeof(double));
    if (fluxTop) std::memcpy(fluxTop, f.fluxTopChannel.data(), f.fluxTopChannel.size() * sizeof(double));
    if (fluxBottom) std::memcpy(fluxBottom, f.fluxBottomChannel.data(), f.fluxBottomChannel.size() * sizeof(double));
}
REF_API void ref_fenics_set_boundary_value(void *h, double v) { ((FenicsRef *)h)->f->setBoundaryValues(v); }

REF_API int ref_fenics_step(void *h)
{
    try {
        ((FenicsRef *)h)->f->stepDiffusion();
        return 0;
    } catch (const std::exception &e) {
        g_err = e.what();
        return 1;
    }
}
REF_API double ref_fenics_total_boundary_flux(void *h) { return ((FenicsRef *)h)->f->totalBoundaryFlux; }

// the Robin rates the reference bound into the trap forms (after createHSL's decoding) and into the channels
REF_API void ref_fenics_robin(void *h, double *trap_left, double *trap_right, double *chan_left, double *chan_right, double *well)
{
    fenicsInterface &f = *((FenicsRef *)h)->f;
    *trap_left = double(*f.shell->data.r_left);
    *trap_right = double(*f.shell->data.r_right);
    // the channels were wired before createHSL re-decided r_left/r_right: read what their forms hold
    auto cl = std::dynamic_pointer_cast<const dolfin::Constant>(f.topChannel->a->coefficients()[f.topChannel->a->coefficient_number("r1")]);
    auto cr = std::dynamic_pointer_cast<const dolfin::Constant>(f.topChannel->a->coefficients()[f.topChannel->a->coefficient_number("r2")]);
    *chan_left = cl ? double(*cl) : 0.0;
    *chan_right = cr ? double(*cr) : 0.0;
    *well = f.wellScaling;
}
