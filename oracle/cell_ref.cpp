// Builds oracle/_ref/libeq_cell_ref.so: the reference's OWN agent-based-model classes -- eQabm (with
// updateCells' findInteriorPoints / readHSL / writeHSL / setDiffusionTensor lambdas), Ecoli (updatePoleCenters),
// cpmEcoli (rod geometry, ratchet, pointIsInCell), Strain -- compiled in place from /root/reference/src (no copy
// of the sources enters this repository) on the Chipmunk 7.0.1 interface shim in oracle/shim_cpm/ (Chipmunk is
// not vendored by the reference; the shim restates the rigid-body transform arithmetic of its published source
// and makes space/shapes/constraints inert), the MPI stand-in of oracle/shim_petsc/ and the nlohmann/json.hpp the
// reference vendors.  TEST INFRASTRUCTURE ONLY: it pins the oracle's restatement of the cell <-> mesh coupling
// (oracle/eq_oracle.c eqo_point_in_cell / eqo_raster_cell / eqo_update_cells_sequential / eqo_cells_tensor /
// eqo_make_cell) and generates tests/golden/cells_ref.json.
#include "abm/eQabm.h"   // /root/reference/src/abm/eQabm.h (-I$(REF)/src)

#include <cstring>
#include "../eq_b200/host/cellRecords.h"

// src/main.cpp:44 defines this static in the executable; the parity pin is not linked against main.cpp
eQ::data::parametersType eQ::data::parameters;

#define REF_API extern "C" __attribute__((visibility("default")))

// Stand-in for the gene circuit (Strain::computeProteins is out of scope): deposits a0 + a1 * (sampled HSL), and
// remembers what readHSL handed it.
class PinStrain : public Strain {
public:
    PinStrain(const Strain::Params &p) : Strain(p) {}
    double a0 = 0.0, a1 = 0.0, seen = 0.0;
    std::shared_ptr<Strain> clone() const override { return std::make_shared<PinStrain>(*this); }
    std::vector<double> computeProteins(const std::vector<double> &eHSL, const std::vector<double> &, const double) override
    {
        seen = eHSL.empty() ? 0.0 : eHSL[0];
        return std::vector<double>{a0 + a1 * seen};
    }
};

struct CellRef {
    std::shared_ptr<eQabm> abm;
    size_t nH = 0, nW = 0;
    double npm = 0;
};

// One HSL layer, identity dof lookup (what gpuHSL hands Simulation::create_HSLgrid).  Integer micron sizes, as
// upstream (src/simulation.cpp:170-172).
REF_API void *ref_abm_create(int widthMicrons, int heightMicrons, double npm, double Dx, double Dy)
{
    auto &P = eQ::data::parameters;
    P["dt"] = 0.1;
    P["simulationTrapWidthMicrons"] = widthMicrons;
    P["simulationTrapHeightMicrons"] = heightMicrons;
    P["nodesPerMicronSignaling"] = npm;
    P["nodesPerMicronData"] = 1.0;
    P["D_HSL"] = std::vector<double>{1200.0};
    P["membraneDiffusionRates"] = std::vector<double>{3.0};
    P["AnisotropicDiffusion_Axial"] = Dx;
    P["AnisotropicDiffusion_Transverse"] = Dy;
    P["trapType"] = "NOWALLED";
    P["defaultAspectRatioFactor"] = 1.0;
    eQabm::Params ap;
    ap.fileIO = nullptr;
    ap.dataFiles = std::make_shared<eQ::data::files_t>();
    ap.zeroOne = nullptr;
    CellRef *r = new CellRef();
    r->abm = std::make_shared<eQabm>(ap);
    r->npm = npm;
    r->nH = size_t(heightMicrons * npm) + 1;
    r->nW = size_t(widthMicrons * npm) + 1;
    r->abm->createDataVectors(r->nH, r->nW);                      // src/simulation.cpp:371
    for (size_t i = 0; i < r->nH; ++i)
        for (size_t j = 0; j < r->nW; ++j)
            r->abm->writeLookupTable(0, eQ::nodePoint{i, j}, eQ::nodeType(i * r->nW + j));   // :384, identity dofs
    return r;
}

// A fresh cell the way eQabm::initCells builds one (src/abm/eQabm.cpp:100-133), at a chosen place.  Cells are
// visited newest first afterwards (forward_list push_front, src/abm/eQabm.h:87-93).
REF_API void ref_abm_add_cell(void *h, double x, double y, double angle, double length, double a0, double a1)
{
    CellRef *r = (CellRef *)h;
    Ecoli::Params cp;
    r->abm->assignDefaultParameters(cp);
    cp.x = x; cp.y = y; cp.angle = angle; cp.length = length;
    Strain::Params sp;
    sp.whichType = eQ::Cell::strainType::ACTIVATOR;
    sp.dt = 0.1; sp.nodesPerMicronScale = r->npm; sp.numHSL = 1; sp.promoterDelayTimeMins = 7.5; sp.baseData = nullptr;
    auto st = std::make_shared<PinStrain>(sp);
    st->a0 = a0; st->a1 = a1;
    cp.strain = st;
    auto cell = std::make_shared<Ecoli>(cp);
    r->abm->cellList << cell;
}

static std::shared_ptr<Ecoli> nth_cell(CellRef *r, long k)
{
    std::shared_ptr<Ecoli> c;
    r->abm->cellList.beginIteration();
    long i = 0;
    while (++(r->abm->cellList >> c)) {
        if (i == k) return c;
        ++i;
    }
    return nullptr;
}

REF_API long ref_abm_count(void *h) { return (long)((CellRef *)h)->abm->cellList.cellCount(); }

// Moves the two body halves of cell k (list order) as a physics step would have, then runs the reference's own
// post-step update `calls` times (cpmEcoli::updateModel with its ratchet, Ecoli::updatePoleCenters).
REF_API void ref_abm_move_cell(void *h, long k, double ax, double ay, double aangle, double bx, double by, double bangle,
                               int calls)
{
    auto c = nth_cell((CellRef *)h, k);
    cpBodySetPosition(c->cpmCell->bodyA, cpv(ax, ay));
    cpBodySetAngle(c->cpmCell->bodyA, aangle);
    cpBodySetPosition(c->cpmCell->bodyB, cpv(bx, by));
    cpBodySetAngle(c->cpmCell->bodyB, bangle);
    for (int i = 0; i < calls; ++i) {
        c->cpmCell->updateModel();     // src/abm/cpmEcoli.cpp:377-450
        c->updatePoleCenters();        // src/abm/Ecoli.cpp:36-63
    }
}

// The 16-double record of include/eqgpu.h, read from the reference's own objects, for cell k in list order.
REF_API void ref_abm_record(void *h, long k, double *rec)
{
    // the PRODUCT's record builder (eq_b200/host/cellRecords.h), instantiated on the reference's own Ecoli
    auto c = nth_cell((CellRef *)h, k);
    eqgpu::cellRecord(*c, rec);
}

REF_API int ref_abm_point_in_cell(void *h, long k, double x, double y)
{
    return nth_cell((CellRef *)h, k)->cpmCell->pointIsInCell(std::make_pair(x, y)) ? 1 : 0;
}

// eQabm::updateCells (src/abm/eQabm.cpp:234-425) on the field u (in/out, natural order): every cell samples,
// deposits a0 + a1*sample and writes its tensor, in list order.  gathered[k] = what readHSL returned for cell k;
// d11/d22/d12 = the grids after the pass.
REF_API void ref_abm_update_cells(void *h, double *u, double *gathered, double *d11, double *d22, double *d12)
{
    CellRef *r = (CellRef *)h;
    const size_t N = r->nH * r->nW;
    std::memcpy(r->abm->hslSolutionVector[0]->data(), u, sizeof(double) * N);
    r->abm->updateCellModels();
    std::memcpy(u, r->abm->hslSolutionVector[0]->data(), sizeof(double) * N);
    std::shared_ptr<Ecoli> c;
    r->abm->cellList.beginIteration();
    long k = 0;
    while (++(r->abm->cellList >> c)) gathered[k++] = std::static_pointer_cast<PinStrain>(c->strain)->seen;
    for (size_t i = 0; i < r->nH; ++i)
        for (size_t j = 0; j < r->nW; ++j) {
            const eQ::nodePoint pt{i, j};
            d11[i * r->nW + j] = r->abm->D11grid->operator[](pt);
            d22[i * r->nW + j] = r->abm->D22grid->operator[](pt);
            d12[i * r->nW + j] = r->abm->D12grid->operator[](pt);
        }
}

REF_API void ref_abm_destroy(void *h) { delete (CellRef *)h; }
