// cellRecords.h -- host-side builder of the 16-double cell record the device kernels read
// (include/eqgpu.h EQGPU_CELL_STRIDE).  This is what eQabm::updateCells calls once per cell after the
// Chipmunk step in the fused mode: it copies, without recomputing anything, the
// numbers the reference's own lambdas read --
//   findInteriorPoints   src/abm/eQabm.cpp:268-305   polePositionA/B, center
//   pointIsInCell        src/abm/cpmEcoli.cpp:313-327 bodyA position + rotation, offset, vertsA[1].x, radius
//   updatePoleCenters    src/abm/Ecoli.cpp:36-63      (already applied: the poles are read, not rebuilt)
//   writeHSL             src/abm/eQabm.cpp:338-359    length
//   setDiffusionTensor   src/abm/eQabm.cpp:306-325    cos/sin of cpmCell->angle
// Header-only templates: the cell type is the reference's `Ecoli` (src/abm/Ecoli.h) inside the eQ tree;
// nothing here includes a reference header, so the file also builds stand-alone (tests instantiate it on a
// plain struct with the same member names).  No device call is made here.
#pragma once
#include <cmath>
#include <cstddef>
#include <vector>

namespace eqgpu {

constexpr int kCellStride = 16;   // == EQGPU_CELL_STRIDE

// One record.  `Cell` needs: cpmCell (pointer-like to an object with bodyA, offset, radius, vertsA[]),
// polePositionA / polePositionB (std::pair<double,double>), getCenter_x(), getCenter_y(), getLengthMicrons(),
// getAngle() (non-const virtuals upstream, hence the non-const reference).  bodyA is a cpBody*: position and
// rotation go through Chipmunk's public accessors cpBodyGetPosition / cpBodyGetRotation (found by argument-
// dependent lookup at instantiation, so this header includes no Chipmunk header).  The rotation is read from
// the body, never recomputed from the angle: cpBodySetAngle stores (cos, sin) once and pointIsInCell
// (cpBodyWorldToLocal) multiplies with exactly those doubles.
template <class Cell>
inline void cellRecord(Cell &c, double *rec)
{
    const auto &m = *c.cpmCell;
    const auto pos = cpBodyGetPosition(m.bodyA);
    const auto rot = cpBodyGetRotation(m.bodyA);
    rec[0] = pos.x;
    rec[1] = pos.y;
    rec[2] = rot.x;
    rec[3] = rot.y;
    rec[4] = m.offset;            // (L0 - W)/2, src/abm/cpmEcoli.cpp:121
    rec[5] = m.vertsA[1].x;       // the ratcheted newOffset, src/abm/cpmEcoli.cpp:407-415
    rec[6] = m.radius;            // W/2, src/abm/cpmEcoli.cpp:122
    rec[7] = c.polePositionA.first;
    rec[8] = c.polePositionA.second;
    rec[9] = c.polePositionB.first;
    rec[10] = c.polePositionB.second;
    rec[11] = c.getCenter_x();
    rec[12] = c.getCenter_y();
    rec[13] = c.getLengthMicrons();
    const double angle = c.getAngle();   // mean of the two body angles, src/abm/Ecoli.h:42
    rec[14] = std::cos(angle);
    rec[15] = std::sin(angle);
}

// All cells of a container in its own iteration order (the order readHSL/writeHSL results are indexed by).
// `forEach(f)` must call f(Cell&) once per cell; returns the number of records written.
template <class ForEach>
inline std::size_t cellRecords(ForEach &&forEach, std::vector<double> &out)
{
    std::size_t n = 0;
    forEach([&](auto &cell) {
        if (out.size() < (n + 1) * kCellStride) out.resize((n + 1) * kCellStride);
        cellRecord(cell, out.data() + n * kCellStride);
        ++n;
    });
    out.resize(n * kCellStride);
    return n;
}

}  // namespace eqgpu
