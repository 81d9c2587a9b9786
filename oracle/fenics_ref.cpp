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
    std::streambuf *keep = std::cout.rdbuf();
    std::ostringstream sink;
    try {
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
        std::cout.rdbuf(keep);
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

// ---------------------------------------------------------------------------------------------------------------------
// The reference's three trap forms (fenics/hsl.ufl, hslRobin.ufl, hslD.ufl -- the north star names all three; the
// shipped fenicsInterface instantiates hslD only, src/fHSL.cpp:23) assembled through their own generated wrappers and
// kernels on RectangleMesh(nx, ny, "right"), with ARBITRARY constants: source f, Robin rates and external
// concentrations s (both zero as shipped, so the stepDiffusion golden cases never exercise them), tensor fields.
// which: 0 hsl (D, dt, f), 1 hslRobin (ds(0) = left wall, ds(1) = right wall), 2 hslD (ds(1) = left, ds(2) = right).
// Out: the matrix as triplets (row-major by row, ascending column; returns nnz) and the load vector.
// ---------------------------------------------------------------------------------------------------------------------
namespace {
class FieldExpression : public dolfin::Expression {   // nearest-vertex lookup, like AnisotropicDiffusionTensor
public:
    FieldExpression(const double *v, size_t nxp, double hx, double hy, double dflt) : v(v), nxp(nxp), hx(hx), hy(hy), dflt(dflt) {}
    void eval(dolfin::Array<double> &values, const dolfin::Array<double> &x) const override
    {
        values[0] = v ? v[size_t(round(x[1] / hy)) * nxp + size_t(round(x[0] / hx))] : dflt;
    }
    const double *v; size_t nxp; double hx, hy, dflt;
};
template <class A, class L>
long assemble_pair(A &a, L &l, size_t n, long *rows, long *cols, double *vals, double *b)
{
    dolfin::SparseSystem sa, sl;
    sa.A.assign(n, {}); sa.b.assign(n, 0.0);
    sl.b.assign(n, 0.0);
    dolfin::assemble_form(a, &sa, true, nullptr);
    dolfin::assemble_form(l, &sl, false, nullptr);
    long k = 0;
    for (size_t i = 0; i < n; ++i)
        for (const auto &e : sa.A[i]) { rows[k] = (long)i; cols[k] = (long)e.first; vals[k] = e.second; ++k; }
    for (size_t i = 0; i < n; ++i) b[i] = sl.b[i];
    return k;
}
}

REF_API long ref_form_assemble(int which, int nx, int ny, double W, double H, double D, double dt, double f,
                               double rA, double sA, double rB, double sB, const double *d11, const double *d22,
                               const double *d12, const double *u0, long *rows, long *cols, double *vals, double *b)
{
    try {
        auto mesh = std::make_shared<dolfin::RectangleMesh>(0, dolfin::Point(0.0, 0.0), dolfin::Point(W, H), nx, ny, "right");
        const size_t n = mesh->num_vertices();
        auto cD = std::make_shared<dolfin::Constant>(D), cdt = std::make_shared<dolfin::Constant>(dt),
             cf = std::make_shared<dolfin::Constant>(f), crA = std::make_shared<dolfin::Constant>(rA),
             csA = std::make_shared<dolfin::Constant>(sA), crB = std::make_shared<dolfin::Constant>(rB),
             csB = std::make_shared<dolfin::Constant>(sB);
        auto left = std::make_shared<DirichletBoundary_TrapEdge>(H, W, DirichletBoundary_TrapEdge::LEFT);
        auto right = std::make_shared<DirichletBoundary_TrapEdge>(H, W, DirichletBoundary_TrapEdge::RIGHT);
        if (which == 0) {
            auto V = std::make_shared<hsl::FunctionSpace>(mesh);
            auto fu0 = std::make_shared<dolfin::Function>(V);
            fu0->vector()->set_local(std::vector<double>(u0, u0 + n));
            hsl::Form_a a(V, V); hsl::Form_L l(V);
            a.D = cD; a.dt = cdt; l.u0 = fu0; l.dt = cdt; l.f = cf;
            return assemble_pair(a, l, n, rows, cols, vals, b);
        }
        if (which == 1) {
            auto V = std::make_shared<hslRobin::FunctionSpace>(mesh);
            auto fu0 = std::make_shared<dolfin::Function>(V);
            fu0->vector()->set_local(std::vector<double>(u0, u0 + n));
            auto mf = std::make_shared<dolfin::MeshFunction<size_t>>(mesh, mesh->topology().dim() - 1, 99);
            left->mark(*mf, 0); right->mark(*mf, 1);
            hslRobin::Form_a a(V, V); hslRobin::Form_L l(V);
            a.D = cD; a.dt = cdt; a.r0 = crA; a.r1 = crB;
            l.u0 = fu0; l.dt = cdt; l.f = cf; l.r0 = crA; l.s0 = csA; l.r1 = crB; l.s1 = csB;
            a.ds = mf; l.ds = mf;
            return assemble_pair(a, l, n, rows, cols, vals, b);
        }
        auto V = std::make_shared<hslD::FunctionSpace>(mesh);
        auto fu0 = std::make_shared<dolfin::Function>(V);
        fu0->vector()->set_local(std::vector<double>(u0, u0 + n));
        auto mf = std::make_shared<dolfin::MeshFunction<size_t>>(mesh, mesh->topology().dim() - 1, 0);
        left->mark(*mf, 1); right->mark(*mf, 2);
        const double hx = W / nx, hy = H / ny;
        auto e11 = std::make_shared<FieldExpression>(d11, size_t(nx) + 1, hx, hy, 1.0);
        auto e22 = std::make_shared<FieldExpression>(d22, size_t(nx) + 1, hx, hy, 1.0);
        auto e12 = std::make_shared<FieldExpression>(d12, size_t(nx) + 1, hx, hy, 0.0);
        hslD::Form_a a(V, V); hslD::Form_L l(V);
        a.D = cD; a.D11 = e11; a.D22 = e22; a.D12 = e12; a.dt = cdt; a.r1 = crA; a.r2 = crB;
        l.u0 = fu0; l.dt = cdt; l.f = cf; l.r1 = crA; l.s1 = csA; l.r2 = crB; l.s2 = csB;
        a.ds = mf; l.ds = mf;
        return assemble_pair(a, l, n, rows, cols, vals, b);
    } catch (const std::exception &e) {
        g_err = e.what();
        return -1;
    }
}
