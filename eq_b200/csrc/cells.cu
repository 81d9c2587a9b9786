// Cell <-> mesh coupling: the lambdas of eQabm::updateCells
// (src/abm/eQabm.cpp:268-359) -- findInteriorPoints, readHSL, writeHSL -- with
// one warp per rod.  The point-in-rod predicate restates cpmEcoli::pointIsInCell
// (src/abm/cpmEcoli.cpp:313-327) on top of Chipmunk 7.0.1's cpBodyWorldToLocal
// and cpvcross [ext]; all of its arithmetic uses explicit round-to-nearest
// mul/add/sub intrinsics so that nvcc cannot contract a*b+c into an FMA: the
// node set must equal the CPU's bit for bit.
#include "eqgpu_internal.cuh"

#define CELLS_PER_BLOCK 8  // warps per block

struct CellBox { int i1, i2, j1, j2; };

__device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }

// src/eQ.h:119-130: size_t(round(x*n)), C round() = half away from zero
__device__ __forceinline__ long long ij_round(double x, double n) { return (long long)round(mul(x, n)); }

__device__ __forceinline__ bool point_in_cell(const double *c, double px, double py)
{
    const double posx = c[0], posy = c[1], rx = c[2], ry = c[3];
    const double off = c[4], noff = c[5], rad = c[6];
    // body->transform = (a=rot.x, b=rot.y, c=-rot.y, d=rot.x, tx=p.x, ty=p.y);
    // cpTransformRigidInverse, then cpTransformPoint (left-to-right sums)
    const double tc = -ry;
    const double ia = rx, ic = -tc, itx = sub(mul(tc, posy), mul(posx, rx));
    const double ib = -ry, id = rx, ity = sub(mul(posx, ry), mul(rx, posy));
    const double lx = add(add(mul(ia, px), mul(ic, py)), itx);
    const double ly = add(add(mul(ib, px), mul(id, py)), ity);
    // vertsA (cpmEcoli.cpp:127-130,177-180,413-415) and edges (:71-76)
    const double v0x = -off, v0y = rad, v1x = noff, v1y = rad;
    const double v2x = noff, v2y = -rad, v3x = -off, v3y = -rad;
    const double e0x = sub(v1x, v0x), e0y = sub(v1y, v0y);
    const double e1x = sub(v2x, v1x), e1y = sub(v2y, v1y);
    const double e2x = sub(v3x, v2x), e2y = sub(v3y, v2y);
    const double e3x = sub(v0x, v3x), e3y = sub(v0y, v3y);
    const double p0x = sub(lx, v1x), p0y = sub(ly, v1y);
    const double p1x = sub(lx, v2x), p1y = sub(ly, v2y);
    const double p2x = sub(lx, v3x), p2y = sub(ly, v3y);
    const double p3x = sub(lx, v0x), p3y = sub(ly, v0y);
    return (sub(mul(e0x, p0y), mul(e0y, p0x)) < 0.0) && (sub(mul(e1x, p1y), mul(e1y, p1x)) < 0.0) &&
           (sub(mul(e2x, p2y), mul(e2y, p2x)) < 0.0) && (sub(mul(e3x, p3y), mul(e3y, p3x)) < 0.0);
}

// search box of findInteriorPoints (src/abm/eQabm.cpp:274-287)
__device__ __forceinline__ CellBox cell_box(const double *c, double npm, int nH, int nW, int nte)
{
    long long ai = ij_round(c[8], npm), aj = ij_round(c[7], npm);
    long long bi = ij_round(c[10], npm), bj = ij_round(c[9], npm);
    long long i1 = min(ai, bi), i2 = max(ai, bi), j1 = min(aj, bj), j2 = max(aj, bj);
    i1 = (i1 >= nte) ? (i1 - nte) : 0;
    j1 = (j1 >= nte) ? (j1 - nte) : 0;
    i2 = ((i2 + nte) >= (long long)(nH - 1)) ? (nH - 1) : (i2 + nte);
    j2 = ((j2 + nte) >= (long long)(nW - 1)) ? (nW - 1) : (j2 + nte);
    CellBox b; b.i1 = (int)i1; b.i2 = (int)i2; b.j1 = (int)j1; b.j2 = (int)j2;
    return b;
}

// Walks the search box 32 nodes at a time in the reference's row-major order.
// visit(node, inside_mask, lane_is_inside) is called once per chunk by the
// whole warp; returns the number of interior points.
template <class F>
__device__ __forceinline__ int raster_walk(const double *c, double npm, int nH, int nW, int nte, F visit)
{
    const int lane = threadIdx.x & 31;
    const CellBox b = cell_box(c, npm, nH, nW, nte);
    const int bw = b.j2 - b.j1 + 1, bh = b.i2 - b.i1 + 1;
    const int total = (b.i2 >= b.i1 && b.j2 >= b.j1) ? bw * bh : 0;
    int count = 0;
    for (int base = 0; base < total; base += 32) {
        const int t = base + lane;
        bool in = false;
        long long node = -1;
        if (t < total) {
            const int pi = b.i1 + t / bw, pj = b.j1 + t % bw;
            // src/eQ.h:115-118 xy_from_ij
            const double x = __ddiv_rn((double)pj, npm), y = __ddiv_rn((double)pi, npm);
            in = point_in_cell(c, x, y);
            node = (long long)pi * nW + pj;
        }
        const unsigned m = __ballot_sync(0xffffffffu, in);
        visit(node, m, in, count);
        count += __popc(m);
    }
    return count;
}

// fallback point when no node is inside (src/abm/eQabm.cpp:299-303)
__device__ __forceinline__ long long centre_node(const double *c, double npm, int nW)
{
    return ij_round(c[12], npm) * (long long)nW + ij_round(c[11], npm);
}

__global__ void __launch_bounds__(32 * CELLS_PER_BLOCK)
k_cells_raster(const double *__restrict__ cells, long long ncells, double npm, int nH, int nW, int nte,
               int *__restrict__ counts, long long *__restrict__ nodes, int cap)
{
    const long long k = (long long)blockIdx.x * CELLS_PER_BLOCK + (threadIdx.x >> 5);
    if (k >= ncells) return;
    const int lane = threadIdx.x & 31;
    const double *c = cells + k * EQGPU_CELL_STRIDE;
    long long *out = nodes ? nodes + k * cap : nullptr;
    int n = raster_walk(c, npm, nH, nW, nte, [&](long long node, unsigned m, bool in, int count) {
        if (in && out) {
            const int pos = count + __popc(m & ((1u << lane) - 1u));
            if (pos < cap) out[pos] = node;
        }
    });
    if (n == 0) {
        if (lane == 0 && out && cap > 0) out[0] = centre_node(c, npm, nW);
        n = 1;
    }
    if (lane == 0) counts[k] = n;
}

// readHSL (src/abm/eQabm.cpp:326-337): mean over the cell's points, summed in
// the reference's order (every lane carries the same running sum).
__global__ void __launch_bounds__(32 * CELLS_PER_BLOCK)
k_cells_gather(const double *__restrict__ cells, long long ncells, double npm, int nH, int nW, int nte,
               const double *__restrict__ u, double *__restrict__ out, int *__restrict__ counts, int row0, int g0,
               int g1)
{
    const long long k = (long long)blockIdx.x * CELLS_PER_BLOCK + (threadIdx.x >> 5);
    if (k >= ncells) return;
    const int lane = threadIdx.x & 31;
    const double *c = cells + k * EQGPU_CELL_STRIDE;
    double HSL = 0.0;
    // row-slab mode: only nodes in owned rows [g0, g1) are read (the partial means are summed over ranks)
    auto owned = [&](long long node) { const int gi = (int)(node / nW); return gi >= g0 && gi < g1; };
    int n = raster_walk(c, npm, nH, nW, nte, [&](long long node, unsigned m, bool in, int) {
        const double v = (in && owned(node)) ? __ldg(u + node - (long long)row0 * nW) : 0.0;
        while (m) {
            const int src = __ffs(m) - 1;
            HSL = add(HSL, __shfl_sync(0xffffffffu, v, src));
            m &= m - 1;
        }
    });
    if (n == 0) {
        const long long cn = centre_node(c, npm, nW);
        HSL = owned(cn) ? __ldg(u + cn - (long long)row0 * nW) : 0.0;
        n = 1;
    }
    if (lane == 0) {
        out[k] = __ddiv_rn(HSL, (double)n);
        counts[k] = n;   // the point count writeHSL divides by: the scatter that follows needs no raster pass
    }
}

// src/eQcell.h:40-93 (poleRadius = 1/2)
__device__ __forceinline__ double cell_volume(double L)
{
    const double poleRadius = 1.0 / 2.0;
    const double cyl = 3.14159265358979323846 * (poleRadius * poleRadius);
    const double poleVolume = 4.0 / 3.0 * 3.14159265358979323846 * (poleRadius * poleRadius * poleRadius);
    return add(mul(sub(L, 1.0), cyl), poleVolume);
}

// writeHSL (src/abm/eQabm.cpp:338-359): nM -> molecules -> per-node increment,
// same amount on every interior point.  Rods of one colony do not overlap, so
// each node receives at most a handful of (usually one) atomic adds.
__global__ void __launch_bounds__(32 * CELLS_PER_BLOCK)
k_cells_scatter(const double *__restrict__ cells, long long ncells, double npm, int nH, int nW, int nte,
                const double *__restrict__ amount, const int *__restrict__ counts,
                double *__restrict__ u, int row0, int g0, int g1)
{
    const long long k = (long long)blockIdx.x * CELLS_PER_BLOCK + (threadIdx.x >> 5);
    if (k >= ncells) return;
    const int lane = threadIdx.x & 31;
    const double *c = cells + k * EQGPU_CELL_STRIDE;
    const int n = counts[k];
    const double L = c[13];
    const double vol = cell_volume(L);
    const double nanoMolarPerMoleculePerCubicMicron = 1.0 / 0.602;
    const double numberHSL = mul(__ddiv_rn(amount[k], nanoMolarPerMoleculePerCubicMicron), vol);
    const double extra = sub(1.0, __ddiv_rn(vol, mul(mul(L, 1.0), 1.0)));
    const double perSquareMicron = __ddiv_rn(numberHSL, extra);
    const double onePoint = mul(mul(perSquareMicron, npm), npm);
    const double dHSL = __ddiv_rn(onePoint, (double)n);
    auto owned = [&](long long node) { const int gi = (int)(node / nW); return gi >= g0 && gi < g1; };
    int found = raster_walk(c, npm, nH, nW, nte, [&](long long node, unsigned, bool in, int) {
        if (in && owned(node)) atomicAdd(u + node - (long long)row0 * nW, dHSL);
    });
    if (found == 0 && lane == 0) {
        const long long cn = centre_node(c, npm, nW);
        if (owned(cn)) atomicAdd(u + cn - (long long)row0 * nW, dHSL);
    }
}

// ---------------------------------------------------------------------------
// Binned scatter (dense colonies): rods are bucketed by the 64x64-node tile that holds the
// origin of their search box; one CTA per non-empty tile rasterises its rods into a
// shared-memory window (tile + BIN_EXT nodes, which covers every search box that starts
// in the tile) with shared-memory atomics, then flushes the touched nodes to the field
// with one global atomic each.  Overlapping rods thus collide in shared memory instead
// of at the L2 atomic units.  Rods whose box is larger than BIN_EXT fall back to direct
// global atomics.
// ---------------------------------------------------------------------------
#define BIN_T 64
#define BIN_EXT 32
#define BIN_W (BIN_T + BIN_EXT)

__global__ void k_bin_count(const double *__restrict__ cells, long long ncells, double npm, int nH, int nW, int nte,
                            int ntx, int *__restrict__ tile_of, int *__restrict__ tile_count)
{
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= ncells) return;
    const CellBox b = cell_box(cells + k * EQGPU_CELL_STRIDE, npm, nH, nW, nte);
    int t = -1;  // -1: direct path (box too large for the window, or degenerate)
    if (b.i2 >= b.i1 && b.j2 >= b.j1 && b.i2 - (b.i1 / BIN_T) * BIN_T < BIN_W && b.j2 - (b.j1 / BIN_T) * BIN_T < BIN_W) {
        t = (b.i1 / BIN_T) * ntx + b.j1 / BIN_T;
        atomicAdd(tile_count + t, 1);
    }
    tile_of[k] = t;
}

// exclusive scan of the tile counts (single block) -> tile_start[0..ntiles]; resets the fill cursors
__global__ void k_bin_scan(int ntiles, const int *__restrict__ tile_count, int *__restrict__ tile_start,
                           int *__restrict__ tile_fill)
{
    __shared__ int carry;
    __shared__ int wsum[32];
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int base = 0; base < ntiles; base += blockDim.x) {
        const int t = base + threadIdx.x;
        const int v = t < ntiles ? tile_count[t] : 0;
        int inc = v;
        for (int o = 1; o < 32; o <<= 1) { const int n = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += n; }
        if (lane == 31) wsum[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            int w = lane < (int)(blockDim.x >> 5) ? wsum[lane] : 0;
            for (int o = 1; o < 32; o <<= 1) { const int n = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += n; }
            wsum[lane] = w;
        }
        __syncthreads();
        const int excl = carry + (warp > 0 ? wsum[warp - 1] : 0) + inc - v;
        if (t < ntiles) { tile_start[t] = excl; tile_fill[t] = 0; }
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) tile_start[ntiles] = carry;
}

__global__ void k_bin_fill(long long ncells, const int *__restrict__ tile_of, const int *__restrict__ tile_start,
                           int *__restrict__ tile_fill, int *__restrict__ list)
{
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= ncells) return;
    const int t = tile_of[k];
    if (t < 0) return;
    list[tile_start[t] + atomicAdd(tile_fill + t, 1)] = (int)k;
}

__device__ __forceinline__ double deposit_of(const double *c, double amount, double npm, int n)
{
    const double L = c[13];
    const double vol = cell_volume(L);
    const double nanoMolarPerMoleculePerCubicMicron = 1.0 / 0.602;
    const double numberHSL = mul(__ddiv_rn(amount, nanoMolarPerMoleculePerCubicMicron), vol);
    const double extra = sub(1.0, __ddiv_rn(vol, mul(mul(L, 1.0), 1.0)));
    const double perSquareMicron = __ddiv_rn(numberHSL, extra);
    const double onePoint = mul(mul(perSquareMicron, npm), npm);
    return __ddiv_rn(onePoint, (double)n);
}

__global__ void __launch_bounds__(32 * CELLS_PER_BLOCK)
k_cells_scatter_binned(const double *__restrict__ cells, double npm, int nH, int nW, int nte, int ntx,
                       const double *__restrict__ amount, const int *__restrict__ counts,
                       const int *__restrict__ tile_start, const int *__restrict__ list, double *__restrict__ u,
                       int row0, int g0, int g1)
{
    extern __shared__ double acc[];  // BIN_W x BIN_W window
    const int t = blockIdx.x, n0 = tile_start[t], n1 = tile_start[t + 1];
    if (n1 == n0) return;
    const int oi = (t / ntx) * BIN_T, oj = (t % ntx) * BIN_T;
    for (int q = threadIdx.x; q < BIN_W * BIN_W; q += blockDim.x) acc[q] = 0.0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int e = n0 + warp; e < n1; e += CELLS_PER_BLOCK) {
        const int k = list[e];
        const double *c = cells + (long long)k * EQGPU_CELL_STRIDE;
        const double dHSL = deposit_of(c, amount[k], npm, counts[k]);
        int found = raster_walk(c, npm, nH, nW, nte, [&](long long node, unsigned, bool in, int) {
            if (in) {
                const int gi = (int)(node / nW), gj = (int)(node - (long long)gi * nW);
                atomicAdd(acc + (gi - oi) * BIN_W + (gj - oj), dHSL);
            }
        });
        if (found == 0 && lane == 0) {
            const long long cn = centre_node(c, npm, nW);
            const int gi = (int)(cn / nW), gj = (int)(cn - (long long)gi * nW);
            if (gi - oi >= 0 && gi - oi < BIN_W && gj - oj >= 0 && gj - oj < BIN_W)
                atomicAdd(acc + (gi - oi) * BIN_W + (gj - oj), dHSL);
            else if (gi >= g0 && gi < g1)
                atomicAdd(u + cn - (long long)row0 * nW, dHSL);
        }
    }
    __syncthreads();
    for (int q = threadIdx.x; q < BIN_W * BIN_W; q += blockDim.x) {
        const double v = acc[q];
        if (v != 0.0) {
            const int gi = oi + q / BIN_W, gj = oj + q % BIN_W;
            if (gi >= g0 && gi < g1 && gj < nW) atomicAdd(u + (size_t)(gi - row0) * nW + gj, v);
        }
    }
}

// rods that did not fit a window: direct global atomics
__global__ void __launch_bounds__(32 * CELLS_PER_BLOCK)
k_cells_scatter_rest(const double *__restrict__ cells, long long ncells, double npm, int nH, int nW, int nte,
                     const double *__restrict__ amount, const int *__restrict__ counts,
                     const int *__restrict__ tile_of, double *__restrict__ u, int row0, int g0, int g1)
{
    const long long k = (long long)blockIdx.x * CELLS_PER_BLOCK + (threadIdx.x >> 5);
    if (k >= ncells || tile_of[k] >= 0) return;
    const int lane = threadIdx.x & 31;
    const double *c = cells + k * EQGPU_CELL_STRIDE;
    const double dHSL = deposit_of(c, amount[k], npm, counts[k]);
    auto owned = [&](long long node) { const int gi = (int)(node / nW); return gi >= g0 && gi < g1; };
    int found = raster_walk(c, npm, nH, nW, nte, [&](long long node, unsigned, bool in, int) {
        if (in && owned(node)) atomicAdd(u + node - (long long)row0 * nW, dHSL);
    });
    if (found == 0 && lane == 0) {
        const long long cn = centre_node(c, npm, nW);
        if (owned(cn)) atomicAdd(u + cn - (long long)row0 * nW, dHSL);
    }
}

// ---------------------------------------------------------------------------
// setDiffusionTensor (src/abm/eQabm.cpp:246-248,306-325,407): the D11/D22/D12 grids are reset to 1,1,0
// and every cell writes the rotated tensor on its interior points; cells are visited in list order, so on
// a node shared by two rods the LATER record wins.  Two passes reproduce that without ordering the rods:
// k_tensor_owner leaves the highest record index per node (atomicMax), k_tensor_write lets exactly that
// rod store.  cos/sin of cpmCell->angle come in record slots 14,15 from the host's libm (as upstream), and
// the three products are explicit round-to-nearest mul/add, so the grids equal the CPU's bit for bit.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_tensor_reset(size_t n, double *__restrict__ d11, double *__restrict__ d22, double *__restrict__ d12,
               int *__restrict__ owner)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < n; g += stride) {
        d11[g] = 1.0; d22[g] = 1.0; d12[g] = 0.0; owner[g] = -1;
    }
}

__global__ void __launch_bounds__(32 * CELLS_PER_BLOCK)
k_tensor_owner(const double *__restrict__ cells, long long ncells, double npm, int nH, int nW, int nte,
               int *__restrict__ owner)
{
    const long long k = (long long)blockIdx.x * CELLS_PER_BLOCK + (threadIdx.x >> 5);
    if (k >= ncells) return;
    const int lane = threadIdx.x & 31;
    const double *c = cells + k * EQGPU_CELL_STRIDE;
    const long long N = (long long)nH * nW;
    int found = raster_walk(c, npm, nH, nW, nte, [&](long long node, unsigned, bool in, int) {
        if (in) atomicMax(owner + node, (int)k);
    });
    if (found == 0 && lane == 0) {
        const long long cn = centre_node(c, npm, nW);
        if (cn >= 0 && cn < N) atomicMax(owner + cn, (int)k);   // gridFunction::isValidIndex (src/eQ.h:66-69)
    }
}

__global__ void __launch_bounds__(32 * CELLS_PER_BLOCK)
k_tensor_write(const double *__restrict__ cells, long long ncells, double npm, int nH, int nW, int nte,
               double Dx, double Dy, const int *__restrict__ owner, double *__restrict__ d11,
               double *__restrict__ d22, double *__restrict__ d12)
{
    const long long k = (long long)blockIdx.x * CELLS_PER_BLOCK + (threadIdx.x >> 5);
    if (k >= ncells) return;
    const int lane = threadIdx.x & 31;
    const double *c = cells + k * EQGPU_CELL_STRIDE;
    const long long N = (long long)nH * nW;
    const double ct = c[14], st = c[15];
    const double cos2t = mul(ct, ct), sin2t = mul(st, st), sincost = mul(st, ct);
    const double v11 = add(mul(Dx, cos2t), mul(Dy, sin2t));
    const double v22 = add(mul(Dx, sin2t), mul(Dy, cos2t));
    const double v12 = mul(sub(Dx, Dy), sincost);
    int found = raster_walk(c, npm, nH, nW, nte, [&](long long node, unsigned, bool in, int) {
        if (in && owner[node] == (int)k) { d11[node] = v11; d22[node] = v22; d12[node] = v12; }
    });
    if (found == 0 && lane == 0) {
        const long long cn = centre_node(c, npm, nW);
        if (cn >= 0 && cn < N && owner[cn] == (int)k) { d11[cn] = v11; d22[cn] = v22; d12[cn] = v12; }
    }
}

static int nte_of(double npm) { return (int)llround(npm * 1.0 / 2.0); }  // src/abm/eQabm.cpp:75

int cells_raster(eqgpu_solver *s, int32_t *d_counts, long long *d_nodes, int cap)
{
    if (s->ncells == 0) return 0;
    const int blocks = (int)((s->ncells + CELLS_PER_BLOCK - 1) / CELLS_PER_BLOCK);
    k_cells_raster<<<blocks, 32 * CELLS_PER_BLOCK, 0, s->stream>>>(
        s->cells, s->ncells, s->npm, s->p.nH, s->p.nW, nte_of(s->npm), d_counts, d_nodes, cap);
    s->launches++;
    EQ_CUDA(cudaGetLastError());
    return 0;
}

int cells_gather(eqgpu_solver *s, double *d_out)
{
    if (s->ncells == 0) return 0;
    const int blocks = (int)((s->ncells + CELLS_PER_BLOCK - 1) / CELLS_PER_BLOCK);
    k_cells_gather<<<blocks, 32 * CELLS_PER_BLOCK, 0, s->stream>>>(
        s->cells, s->ncells, s->npm, s->p.nH, s->p.nW, nte_of(s->npm), s->u, d_out, s->cell_counts,
        s->levels[0].dev.row0, s->levels[0].g0, s->levels[0].g1);
    s->launches++;
    s->counts_valid = true;   // until the next eqgpu_cells_upload
    EQ_CUDA(cudaGetLastError());
    if (s->slab) {
        int rc = slab_allreduce(s, d_out, d_out, (int)s->ncells);
        if (rc) return rc;
    }
    return 0;
}

static int cells_scatter_binned(eqgpu_solver *s, const double *d_amount)
{
    const int nW = s->p.nW, nH = s->p.nH, nte = nte_of(s->npm);
    const int ntx = (nW + BIN_T - 1) / BIN_T, nty = (nH + BIN_T - 1) / BIN_T, ntiles = ntx * nty;
    const long long n = s->ncells;
    if (s->bin_cap_cells < n || s->bin_cap_tiles < ntiles) {
        cudaFree(s->bin_ints);
        s->bin_ints = nullptr;
        const size_t ints = (size_t)2 * (n + 64) + (size_t)3 * (ntiles + 1);
        EQ_CUDA(cudaMalloc(&s->bin_ints, sizeof(int) * ints));
        s->bin_cap_cells = n + 64;
        s->bin_cap_tiles = ntiles;
    }
    int *tile_of = s->bin_ints, *list = tile_of + s->bin_cap_cells, *tile_count = list + s->bin_cap_cells;
    int *tile_start = tile_count + (ntiles + 1), *tile_fill = tile_start + (ntiles + 1);
    const int blocks = (int)((n + CELLS_PER_BLOCK - 1) / CELLS_PER_BLOCK), tb = (int)((n + 255) / 256);
    const Level &l0 = s->levels[0];
    const size_t smem = sizeof(double) * BIN_W * BIN_W;
    // per device, not per process: set on every call (a host-side table write)
    EQ_CUDA(cudaFuncSetAttribute(k_cells_scatter_binned, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    EQ_CUDA(cudaMemsetAsync(tile_count, 0, sizeof(int) * (ntiles + 1), s->stream));
    k_cells_raster<<<blocks, 32 * CELLS_PER_BLOCK, 0, s->stream>>>(s->cells, n, s->npm, nH, nW, nte, s->cell_counts,
                                                                   nullptr, 0);
    k_bin_count<<<tb, 256, 0, s->stream>>>(s->cells, n, s->npm, nH, nW, nte, ntx, tile_of, tile_count);
    k_bin_scan<<<1, 1024, 0, s->stream>>>(ntiles, tile_count, tile_start, tile_fill);
    k_bin_fill<<<tb, 256, 0, s->stream>>>(n, tile_of, tile_start, tile_fill, list);
    k_cells_scatter_binned<<<ntiles, 32 * CELLS_PER_BLOCK, smem, s->stream>>>(
        s->cells, s->npm, nH, nW, nte, ntx, d_amount, s->cell_counts, tile_start, list, s->u, l0.dev.row0, l0.g0, l0.g1);
    k_cells_scatter_rest<<<blocks, 32 * CELLS_PER_BLOCK, 0, s->stream>>>(
        s->cells, n, s->npm, nH, nW, nte, d_amount, s->cell_counts, tile_of, s->u, l0.dev.row0, l0.g0, l0.g1);
    s->launches += 6;
    EQ_CUDA(cudaGetLastError());
    return 0;
}

int cells_scatter(eqgpu_solver *s, const double *d_amount)
{
    if (s->ncells == 0) return 0;
    if (s->scatter_mode == 1) return cells_scatter_binned(s, d_amount);
    const int blocks = (int)((s->ncells + CELLS_PER_BLOCK - 1) / CELLS_PER_BLOCK);
    // point counts first (the per-node amount divides by them), unless a gather of the same cell set left them
    if (!s->counts_valid) {
        k_cells_raster<<<blocks, 32 * CELLS_PER_BLOCK, 0, s->stream>>>(
            s->cells, s->ncells, s->npm, s->p.nH, s->p.nW, nte_of(s->npm), s->cell_counts, nullptr, 0);
        s->launches++;
        s->counts_valid = true;
    }
    k_cells_scatter<<<blocks, 32 * CELLS_PER_BLOCK, 0, s->stream>>>(
        s->cells, s->ncells, s->npm, s->p.nH, s->p.nW, nte_of(s->npm), d_amount, s->cell_counts, s->u,
        s->levels[0].dev.row0, s->levels[0].g0, s->levels[0].g1);
    s->launches++;
    EQ_CUDA(cudaGetLastError());
    return 0;
}

// D11/D22/D12 from the uploaded rods into the solver's tensor fields (allocated on first use).
int cells_tensor(eqgpu_solver *s, double Dx, double Dy)
{
    const size_t n = s->N;
    if (!s->d11) {
        EQ_CUDA(cudaMalloc(&s->d11, sizeof(double) * n));
        EQ_CUDA(cudaMalloc(&s->d22, sizeof(double) * n));
        EQ_CUDA(cudaMalloc(&s->d12, sizeof(double) * n));
    }
    if (!s->tensor_owner) EQ_CUDA(cudaMalloc(&s->tensor_owner, sizeof(int) * n));
    k_tensor_reset<<<4 * s->num_sms, 256, 0, s->stream>>>(n, s->d11, s->d22, s->d12, s->tensor_owner);
    s->launches++;
    if (s->ncells > 0) {
        const int blocks = (int)((s->ncells + CELLS_PER_BLOCK - 1) / CELLS_PER_BLOCK);
        k_tensor_owner<<<blocks, 32 * CELLS_PER_BLOCK, 0, s->stream>>>(s->cells, s->ncells, s->npm, s->p.nH, s->p.nW,
                                                                       nte_of(s->npm), s->tensor_owner);
        k_tensor_write<<<blocks, 32 * CELLS_PER_BLOCK, 0, s->stream>>>(s->cells, s->ncells, s->npm, s->p.nH, s->p.nW,
                                                                       nte_of(s->npm), Dx, Dy, s->tensor_owner, s->d11,
                                                                       s->d22, s->d12);
        s->launches += 2;
    }
    EQ_CUDA(cudaGetLastError());
    return 0;
}
