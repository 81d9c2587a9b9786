// Stand-alone instantiation of cellRecords.h on plain structs that carry the member names of the reference's
// Ecoli / cpmEcoli / cpBody (src/abm/Ecoli.h:39-61, src/abm/cpmEcoli.h:73-102): proves the header needs no reference
// or Chipmunk include, and prints the record for tests/test_colony.py to compare.
#include <cstdio>
#include <memory>
#include <utility>
#include "cellRecords.h"

struct cpVect { double x, y; };
struct cpTransform { double a, b, c, d, tx, ty; };
struct cpBody { cpVect p; cpTransform transform; };
static cpVect cpBodyGetPosition(const cpBody *b) { return cpVect{b->transform.tx, b->transform.ty}; }
static cpVect cpBodyGetRotation(const cpBody *b) { return cpVect{b->transform.a, b->transform.b}; }

struct MockCpm {
    cpBody *bodyA;
    cpVect vertsA[4];
    double radius, offset, length, angle;
    cpVect center;
};
struct MockCell {
    std::shared_ptr<MockCpm> cpmCell;
    std::pair<double, double> polePositionA, polePositionB;
    double getLengthMicrons() { return cpmCell->length; }
    double getCenter_x() { return cpmCell->center.x; }
    double getCenter_y() { return cpmCell->center.y; }
    double getAngle() { return cpmCell->angle; }
};

int main()
{
    cpBody body{{3.0, 4.0}, {0.6, 0.8, -0.8, 0.6, 3.0, 4.0}};
    auto m = std::make_shared<MockCpm>();
    m->bodyA = &body;
    m->vertsA[1] = cpVect{1.35, 0.5};
    m->radius = 0.5; m->offset = 1.05; m->length = 3.4; m->angle = 0.9272952180016122; m->center = cpVect{3.09, 4.12};
    MockCell cells[2];
    for (auto &c : cells) { c.cpmCell = m; c.polePositionA = {4.0, 5.0}; c.polePositionB = {1.5, 2.5}; }
    std::vector<double> out;
    const std::size_t n = eqgpu::cellRecords([&](auto f) { for (auto &c : cells) f(c); }, out);
    if (n != 2 || out.size() != 2 * eqgpu::kCellStride) return 1;
    for (int k = 0; k < eqgpu::kCellStride; ++k) std::printf("%.17g\n", out[eqgpu::kCellStride + k]);
    return 0;
}
