// Register-tile multigrid smoothers (isotropic operator, one GPU): the kernels that carry the bytes of a V-cycle.
//
// What they replace, and why.  The shared-memory tile kernels of mg_tile.inc keep a 64 x 64 tile of the iterate in shared
// memory and pay 2 LDS.64 + 1 STS.64 and a share of a __syncthreads for every node update of every sweep; ncu shows them
// waiting on the shared-memory pipe (stalls mio + short_scoreboard 63 %) at 0.3 of the HBM roofline, with the load, sweep
// and store phases of a CTA in sequence.  Here
//   * a WARP owns 8 rows x 64 columns of the tile in REGISTERS (lane l holds columns 2l, 2l+1 of its 8 rows of the iterate
//     and of the right-hand side); north / south neighbours are the lane's own registers, west / east (and the south-west /
//     north-east ones of the "right"-diagonal mesh) arrive as ONE shuffled partial sum per side -- the neighbour lane forms
//     cEW * x(r, c) + cD * x(r -+ 1, c) of the column it owns and shuffles it over -- so a node update costs two SHFL.32 and
//     no shared memory; only the two boundary rows of a warp's band pass through shared memory once per sweep (128-bit);
//   * the inputs of a tile (right-hand side; incoming iterate) arrive by TMA: one cp.async.bulk.tensor.2d of a 64 x 64 box
//     per vector onto an mbarrier, issued for the NEXT tile as soon as the warps have copied the current one into registers,
//     so HBM latency hides behind the sweeps of the current tile; CTAs are persistent (two per SM) and draw tiles from a
//     counter; results leave by 128-bit stores straight from registers;
//   * only tiles of regular nodes are handled (constant coefficients as immediate operands); the perimeter tiles of a level
//     (walls, Dirichlet rows, the narrower last cell of a coarse grid: ~10 % of the tiles at 2048^2) stay with the general
//     shared-memory kernels, launched on a tile list beside this kernel (solver.cu: launch_pre / launch_post).
// Halo: H nodes per side, H even (128-bit alignment of the owned columns) and >= the stencil passes of the kernel.
//
// Same mathematics as mg_tile.inc (Chebyshev-weighted Jacobi sweeps, full-weighting restriction = P^T, P1 prolongation), so
// the two kernel families can share a level tile by tile; values differ by rounding only (different summation order).
#include "eqgpu_internal.cuh"
#include <algorithm>
#include <cstdlib>

namespace RT {

constexpr int TS = 64;            // tile edge, nodes
constexpr int NWARP = 8, RW = 8;  // warps per CTA, rows per warp
constexpr int NT = 32 * NWARP;
constexpr int PS = 36;            // row stride of the coarse patch (34 x 34 used)
constexpr int PATCH = 34;

struct Geom {
    int gx, gy;                  // tile grid of the whole level (shared with the tile-list kernel)
    int bx0, by0, nbx, nby;      // rectangle of tiles handled here
};

// ---- mbarrier / TMA --------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, unsigned long long *bar, int x, int y)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y)
        : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

// ---- the band of one warp -------------------------------------------------------------------------------------------
// xa / xb: columns 2l / 2l+1, rows 0 .. RW+1 of the window (0 and RW+1: the neighbouring warps' boundary rows)
struct Coef { double cC, cEW, cNS, cD; };

// exch[buf][warp][0 = bottom row, 1 = top row][64]
__device__ __forceinline__ void band_publish(double *exch, int buf, int w, int lane, const double (&xa)[RW + 2],
                                             const double (&xb)[RW + 2])
{
    double *q = exch + ((buf * NWARP + w) * 2) * TS + 2 * lane;
    *reinterpret_cast<double2 *>(q) = make_double2(xa[1], xb[1]);
    *reinterpret_cast<double2 *>(q + TS) = make_double2(xa[RW], xb[RW]);
}
__device__ __forceinline__ void band_halo(const double *exch, int buf, int w, int lane, double (&xa)[RW + 2],
                                          double (&xb)[RW + 2])
{
    double2 lo = make_double2(0.0, 0.0), hi = make_double2(0.0, 0.0);
    if (w > 0) lo = *reinterpret_cast<const double2 *>(exch + ((buf * NWARP + w - 1) * 2 + 1) * TS + 2 * lane);
    if (w < NWARP - 1) hi = *reinterpret_cast<const double2 *>(exch + ((buf * NWARP + w + 1) * 2) * TS + 2 * lane);
    xa[0] = lo.x; xb[0] = lo.y; xa[RW + 1] = hi.x; xb[RW + 1] = hi.y;
}

// One pass over the band.  RESID false: x <- x + wd (b - A x) in place;  true: the residual b - A x of the band's rows goes
// to rows (row0 .. row0+RW-1) of the 64-column shared array `rt`, x is left alone.
// Lane 0's west partial and lane 31's east partial are their own values (the shuffle has no source): the outermost columns
// of the tile belong to the ring that goes stale, one node per pass, like the outermost rows.
template <bool RESID>
__device__ __forceinline__ void band_pass(const Coef &c, double wd, const double (&ba)[RW], const double (&bb)[RW],
                                          double (&xa)[RW + 2], double (&xb)[RW + 2], double *rt, int row0, int lane)
{
    double oa = xa[0], ob = xb[0];
#pragma unroll
    for (int k = 1; k <= RW; ++k) {
        const double ca = xa[k], cb = xb[k], na = xa[k + 1], nb = xb[k + 1];
        const double ma = c.cEW * ca, mb = c.cEW * cb;
        const double pown = fma(c.cD, ob, mb);    // west + south-west share of (k, 2l+2), formed by the owner of column 2l+1
        const double qown = fma(c.cD, na, ma);    // east + north-east share of (k, 2l-1), formed by the owner of column 2l
        const double ea = fma(c.cD, nb, mb);      // east + north-east of my column a
        const double wb = fma(c.cD, oa, ma);      // west + south-west of my column b
        const double pl = __shfl_up_sync(0xffffffffu, pown, 1);
        const double qr = __shfl_down_sync(0xffffffffu, qown, 1);
        const double axa = fma(c.cC, ca, fma(c.cNS, na + oa, ea)) + pl;
        const double axb = fma(c.cC, cb, fma(c.cNS, nb + ob, wb)) + qr;
        const double ra = ba[k - 1] - axa, rb = bb[k - 1] - axb;
        if (RESID) {
            *reinterpret_cast<double2 *>(rt + (row0 + k - 1) * TS + 2 * lane) = make_double2(ra, rb);
        } else {
            xa[k] = fma(wd, ra, ca);
            xb[k] = fma(wd, rb, cb);
        }
        oa = ca; ob = cb;
    }
}

// tile index -> origin
template <int H>
__device__ __forceinline__ void tile_origin(const LevelDev &F, const Geom &G, int t, int &bx, int &by, int &ox, int &oy)
{
    constexpr int TO = TS - 2 * H;
    by = G.by0 + t / G.nbx;
    bx = G.bx0 + t - (t / G.nbx) * G.nbx;
    ox = bx * TO - H;
    oy = F.tbase + by * TO - H;
}

// =====================================================================================================================
// pre-smoothing: NU sweeps from a zero guess, residual, restriction.        reads b        writes x, b_coarse
// =====================================================================================================================
template <int NU, int H>
__global__ void __launch_bounds__(NT, 2)
k_pre_rt(const LevelDev F, const LevelDev Cc, const __grid_constant__ CUtensorMap map_b, double *__restrict__ x,
         double *__restrict__ bc, const SmoothW sw, const Geom G, unsigned *sched, const CGScalars *sc)
{
    static_assert(H % 2 == 0 && H >= NU + 1, "even halo of at least NU + 1 nodes");
    constexpr int TO = TS - 2 * H, CT = TO / 2;
    extern __shared__ __align__(128) unsigned char smraw[];
    double *bbuf = reinterpret_cast<double *>(smraw);   // TMA box: 64 x 64
    double *rt = bbuf + TS * TS;                        // residual tile
    double *exch = rt + TS * TS;                        // 2 x NWARP x 2 x 64
    unsigned long long *bar = reinterpret_cast<unsigned long long *>(exch + 2 * NWARP * 2 * TS);
    int *ids = reinterpret_cast<int *>(bar + 1);
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int ntiles = G.nbx * G.nby;
    if (tid == 0) { mbar_init(bar, 1); fence_barrier_init(); }
    pdl_wait();
    if (sc->done) return;
    if (tid == 0) {
        const int t0 = (int)atomicAdd(sched, 2u);
        ids[0] = t0; ids[1] = t0 + 1;
        if (t0 < ntiles) {
            int bx, by, ox, oy;
            tile_origin<H>(F, G, t0, bx, by, ox, oy);
            mbar_expect_tx(bar, TS * TS * 8);
            tma_load_2d(bbuf, &map_b, bar, ox, oy);
        }
    }
    __syncthreads();
    int cur = ids[0], nxt = ids[1];
    unsigned phase = 0;
    int it = 0, buf = 0;
    const Coef c{F.cC, F.cEW, F.cNS, F.cD};
    const double icC = F.icC;
    while (cur < ntiles) {
        int bx, by, ox, oy;
        tile_origin<H>(F, G, cur, bx, by, ox, oy);
        mbar_wait(bar, phase);
        phase ^= 1u;
        double ba[RW], bb[RW], xa[RW + 2], xb[RW + 2];
#pragma unroll
        for (int k = 0; k < RW; ++k) {
            const double2 v = *reinterpret_cast<const double2 *>(bbuf + (w * RW + k) * TS + 2 * lane);
            ba[k] = v.x; bb[k] = v.y;
        }
        __syncthreads();   // everyone holds its rows: the box may be refilled
        unsigned t_nn = 0;
        if (tid == 0) {
            t_nn = atomicAdd(sched, 1u);   // the tile after the next one; consumed at the end of this tile
            if (nxt < ntiles) {
                int bx2, by2, ox2, oy2;
                tile_origin<H>(F, G, nxt, bx2, by2, ox2, oy2);
                fence_proxy_async();
                mbar_expect_tx(bar, TS * TS * 8);
                tma_load_2d(bbuf, &map_b, bar, ox2, oy2);
            }
        }
        // sweep 1 from zero: x = w0 D^-1 b
        {
            const double wd = sw.w[0] * icC;
#pragma unroll
            for (int k = 0; k < RW; ++k) { xa[k + 1] = wd * ba[k]; xb[k + 1] = wd * bb[k]; }
        }
#pragma unroll
        for (int s = 1; s < NU; ++s) {
            band_publish(exch, buf, w, lane, xa, xb);
            __syncthreads();
            band_halo(exch, buf, w, lane, xa, xb);
            buf ^= 1;
            band_pass<false>(c, sw.w[s] * icC, ba, bb, xa, xb, nullptr, 0, lane);
        }
        band_publish(exch, buf, w, lane, xa, xb);
        __syncthreads();
        band_halo(exch, buf, w, lane, xa, xb);
        buf ^= 1;
        band_pass<true>(c, 0.0, ba, bb, xa, xb, rt, w * RW, lane);
        if (tid == 0) ids[it & 1] = (int)t_nn;
        // pre-smoothed iterate of the owned region
        if (lane >= H / 2 && lane < 32 - H / 2) {
#pragma unroll
            for (int k = 0; k < RW; ++k) {
                const int ly = w * RW + k;
                if (ly >= H && ly < TS - H)
                    *reinterpret_cast<double2 *>(x + (size_t)(oy + ly) * F.nx + ox + 2 * lane) = make_double2(xa[k + 1], xb[k + 1]);
            }
        }
        __syncthreads();   // residual tile complete (and ids[] visible)
        // restriction (P^T, full weighting on the six mesh neighbours): coarse nodes = even fine nodes of the owned region
        {
            const int I0 = (oy + H) >> 1, J0 = (ox + H) >> 1;
            for (int q = tid; q < CT * CT; q += NT) {
                const int cy = q / CT, cx = q - cy * CT;
                const int cc = (H + 2 * cy) * TS + H + 2 * cx;
                const double h = rt[cc + 1] + rt[cc - 1] + rt[cc + TS] + rt[cc - TS] + rt[cc + TS + 1] + rt[cc - TS - 1];
                bc[(size_t)(I0 + cy) * Cc.nx + J0 + cx] = rt[cc] + 0.5 * h;
            }
        }
        const int nn = ids[it & 1];
        cur = nxt; nxt = nn;
        ++it;
        // (the next tile's first write to rt comes after two more __syncthreads: no barrier needed here)
    }
    if (tid == 0) {
        const unsigned prev = atomicAdd(sched + 1, 1u);
        if (prev == gridDim.x - 1) { sched[0] = 0u; sched[1] = 0u; }
    }
}

// =====================================================================================================================
// post-smoothing: x = xin + P x_coarse, NU sweeps (+ x.b).        reads b, xin, x_coarse        writes x
// =====================================================================================================================
__device__ __forceinline__ void patch_fetch(const LevelDev &Cc, const double *__restrict__ xc, int ox, int oy, int tid,
                                            double (&pv)[5])
{
    const int J0 = ox >> 1, I0 = oy >> 1;
#pragma unroll
    for (int q = 0; q < 5; ++q) {
        const int e = tid + q * NT;
        const int ci = e / PATCH, cj = e - ci * PATCH;
        pv[q] = e < PATCH * PATCH ? __ldg(xc + (size_t)min(I0 + ci, Cc.ny - 1) * Cc.nx + min(J0 + cj, Cc.nx - 1)) : 0.0;
    }
}
__device__ __forceinline__ void patch_store(double *patch, int tid, const double (&pv)[5])
{
#pragma unroll
    for (int q = 0; q < 5; ++q) {
        const int e = tid + q * NT;
        const int ci = e / PATCH, cj = e - ci * PATCH;
        if (e < PATCH * PATCH) patch[ci * PS + cj] = pv[q];
    }
}

template <int NU, int H, bool DOT>
__global__ void __launch_bounds__(NT, 2)
k_post_rt(const LevelDev F, const LevelDev Cc, const __grid_constant__ CUtensorMap map_b,
          const __grid_constant__ CUtensorMap map_x, double *__restrict__ x, const double *__restrict__ xc, const SmoothW sw,
          const Geom G, unsigned *sched, CGScalars *sc, double *partials, unsigned *counter, double *out_dot)
{
    static_assert(H % 2 == 0 && H >= NU, "even halo of at least NU nodes");
    extern __shared__ __align__(128) unsigned char smraw[];
    double *bbuf = reinterpret_cast<double *>(smraw);   // TMA boxes: 64 x 64 each
    double *xbuf = bbuf + TS * TS;
    double *exch = xbuf + TS * TS;                      // 2 x NWARP x 2 x 64
    double *patch = exch + 2 * NWARP * 2 * TS;          // PATCH rows of stride PS
    double *red = patch + PATCH * PS;                   // 32 doubles
    unsigned long long *bar = reinterpret_cast<unsigned long long *>(red + 32);
    int *ids = reinterpret_cast<int *>(bar + 1);
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int ntiles = G.nbx * G.nby;
    if (tid == 0) { mbar_init(bar, 1); fence_barrier_init(); }
    pdl_wait();
    if (sc->done) return;
    if (tid == 0) {
        const int t0 = (int)atomicAdd(sched, 2u);
        ids[0] = t0; ids[1] = t0 + 1;
        if (t0 < ntiles) {
            int bx, by, ox, oy;
            tile_origin<H>(F, G, t0, bx, by, ox, oy);
            mbar_expect_tx(bar, 2 * TS * TS * 8);
            tma_load_2d(bbuf, &map_b, bar, ox, oy);
            tma_load_2d(xbuf, &map_x, bar, ox, oy);
        }
    }
    __syncthreads();
    int cur = ids[0], nxt = ids[1];
    if (cur < ntiles) {   // the first tile's coarse patch
        int bx, by, ox, oy;
        tile_origin<H>(F, G, cur, bx, by, ox, oy);
        double pv[5];
        patch_fetch(Cc, xc, ox, oy, tid, pv);
        patch_store(patch, tid, pv);
    }
    unsigned phase = 0, mine = 0;
    int it = 0, buf = 0;
    const Coef c{F.cC, F.cEW, F.cNS, F.cD};
    const double icC = F.icC;
    while (cur < ntiles) {
        int bx, by, ox, oy;
        tile_origin<H>(F, G, cur, bx, by, ox, oy);
        mbar_wait(bar, phase);
        phase ^= 1u;
        double ba[RW], bb[RW], xa[RW + 2], xb[RW + 2];
#pragma unroll
        for (int k = 0; k < RW; ++k) {
            const double2 v = *reinterpret_cast<const double2 *>(bbuf + (w * RW + k) * TS + 2 * lane);
            const double2 u = *reinterpret_cast<const double2 *>(xbuf + (w * RW + k) * TS + 2 * lane);
            ba[k] = v.x; bb[k] = v.y; xa[k + 1] = u.x; xb[k + 1] = u.y;
        }
        __syncthreads();   // boxes consumed, and the patch written at the end of the previous tile (or above) is visible
        unsigned t_nn = 0;
        if (tid == 0) {
            t_nn = atomicAdd(sched, 1u);
            if (nxt < ntiles) {
                int bx2, by2, ox2, oy2;
                tile_origin<H>(F, G, nxt, bx2, by2, ox2, oy2);
                fence_proxy_async();
                mbar_expect_tx(bar, 2 * TS * TS * 8);
                tma_load_2d(bbuf, &map_b, bar, ox2, oy2);
                tma_load_2d(xbuf, &map_x, bar, ox2, oy2);
            }
        }
        // prolongation: tile rows / columns are even <=> coincident with a coarse node (the tile origin is even)
        {
            double p0[5], p1[5];
#pragma unroll
            for (int j = 0; j < 5; ++j) {
                p0[j] = patch[(w * (RW / 2) + j) * PS + lane];
                p1[j] = patch[(w * (RW / 2) + j) * PS + lane + 1];
            }
#pragma unroll
            for (int k = 0; k < RW; ++k) {
                const int j = k >> 1;
                if (k & 1) {   // midpoint row
                    xa[k + 1] += 0.5 * (p0[j] + p0[j + 1]);
                    xb[k + 1] += 0.5 * (p0[j] + p1[j + 1]);
                } else {
                    xa[k + 1] += 0.5 * (p0[j] + p0[j]);
                    xb[k + 1] += 0.5 * (p0[j] + p1[j]);
                }
            }
        }
        // the next tile's patch: loads in flight during the sweeps, stored once every warp is past its prolongation
        double pv[5];
        const bool more = nxt < ntiles;
        if (more) {
            int bx2, by2, ox2, oy2;
            tile_origin<H>(F, G, nxt, bx2, by2, ox2, oy2);
            patch_fetch(Cc, xc, ox2, oy2, tid, pv);
        }
#pragma unroll
        for (int s = 0; s < NU; ++s) {
            band_publish(exch, buf, w, lane, xa, xb);
            __syncthreads();
            band_halo(exch, buf, w, lane, xa, xb);
            buf ^= 1;
            band_pass<false>(c, sw.w[s] * icC, ba, bb, xa, xb, nullptr, 0, lane);
        }
        if (more) patch_store(patch, tid, pv);   // (after >= 1 __syncthreads since the prolongation read the old patch)
        if (tid == 0) ids[it & 1] = (int)t_nn;
        double acc = 0.0;
        if (lane >= H / 2 && lane < 32 - H / 2) {
#pragma unroll
            for (int k = 0; k < RW; ++k) {
                const int ly = w * RW + k;
                if (ly >= H && ly < TS - H) {
                    *reinterpret_cast<double2 *>(x + (size_t)(oy + ly) * F.nx + ox + 2 * lane) = make_double2(xa[k + 1], xb[k + 1]);
                    if (DOT) acc += xa[k + 1] * ba[k] + xb[k + 1] * bb[k];
                }
            }
        }
        if (DOT) {
            const double t = cta_sum(acc, red);   // contains the __syncthreads that publishes ids[] and the patch
            if (tid == 0) partials[by * G.gx + bx] = t;
            ++mine;
        } else {
            __syncthreads();
        }
        const int nn = ids[it & 1];
        cur = nxt; nxt = nn;
        ++it;
    }
    if (DOT && mine) tiles_arrive(partials, counter, mine, (unsigned)(G.gx * G.gy), out_dot);
    if (tid == 0) {
        const unsigned prev = atomicAdd(sched + 1, 1u);
        if (prev == gridDim.x - 1) { sched[0] = 0u; sched[1] = 0u; }
    }
}

}  // namespace RT

// ---------------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// TMA descriptor of an nx x ny fp64 field (row-major, pitch nx) with a 64 x 64 box
static bool make_tile_map(CUtensorMap *map, double *base, int nx, int ny)
{
    static PFN_encodeTiled enc = nullptr;
    static bool looked = false;
    if (!looked) {
        looked = true;
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            enc = (PFN_encodeTiled)fn;
        (void)cudaGetLastError();
    }
    if (!enc || (nx & 1) || (((size_t)base) & 15)) return false;
    const cuuint64_t gdim[2] = {(cuuint64_t)nx, (cuuint64_t)ny};
    const cuuint64_t gstride[1] = {(cuuint64_t)nx * sizeof(double)};
    const cuuint32_t box[2] = {(cuuint32_t)RT::TS, (cuuint32_t)RT::TS};
    const cuuint32_t estr[2] = {1, 1};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, base, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static int rt_halo_pre(int nu) { return (nu + 2) & ~1; }    // even, >= nu + 1
static int rt_halo_post(int nu) { return (nu + 1) & ~1; }   // even, >= nu

static const size_t RT_SMEM_PRE = (size_t)(2 * RT::TS * RT::TS + 2 * RT::NWARP * 2 * RT::TS) * 8 + 64;
static const size_t RT_SMEM_POST = (size_t)(2 * RT::TS * RT::TS + 2 * RT::NWARP * 2 * RT::TS + RT::PATCH * RT::PS + 32) * 8 + 64;

// The largest rectangle of tiles whose 64 x 64 window holds regular nodes only (the regular nodes form a rectangle, so the
// regular tiles do), and the list of all other tiles of the level's gx x gy tiling with halo h.
static int plan_level(eqgpu_solver *s, const LevelDev &F, int h, Level::RtPlan &P)
{
    const int to = RT::TS - 2 * h;
    P.on = false;
    P.gx = (F.nx + to - 1) / to;
    P.gy = (F.ny + to - 1) / to;
    auto regx = [&](int b) { const int o = b * to - h; return o >= 1 && o + RT::TS - 1 <= F.jreg_hi; };
    auto regy = [&](int b) { const int o = b * to - h; return o >= 1 && o + RT::TS - 1 <= F.ireg_hi; };
    int bx0 = 0, bx1 = P.gx, by0 = 0, by1 = P.gy;
    while (bx0 < bx1 && !regx(bx0)) ++bx0;
    while (bx1 > bx0 && !regx(bx1 - 1)) --bx1;
    while (by0 < by1 && !regy(by0)) ++by0;
    while (by1 > by0 && !regy(by1 - 1)) --by1;
    for (int b = bx0; b < bx1; ++b) if (!regx(b)) return 0;
    for (int b = by0; b < by1; ++b) if (!regy(b)) return 0;
    P.bx0 = bx0; P.by0 = by0; P.nbx = bx1 - bx0; P.nby = by1 - by0;
    if (P.nbx * P.nby < s->rt_min_tiles) return 0;
    std::vector<int> tl;
    for (int by = 0; by < P.gy; ++by)
        for (int bx = 0; bx < P.gx; ++bx)
            if (!(bx >= bx0 && bx < bx1 && by >= by0 && by < by1)) { tl.push_back(bx); tl.push_back(by); }
    P.nperim = (int)tl.size() / 2;
    if (P.nperim > 0) {
        EQ_CUDA(cudaMalloc(&P.d_tlist, sizeof(int) * tl.size()));
        EQ_CUDA(cudaMemcpy(P.d_tlist, tl.data(), sizeof(int) * tl.size(), cudaMemcpyHostToDevice));
    }
    P.on = true;
    return 0;
}

template <class K>
static bool set_smem(K kernel, size_t bytes)
{
    const bool ok = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) == cudaSuccess;
    (void)cudaGetLastError();
    return ok;
}

int rt_setup(eqgpu_solver *s)
{
    s->rt_smooth = !s->slab && s->fused && !s->tensor;
    if (const char *e = getenv("EQGPU_RT")) s->rt_smooth = s->rt_smooth && atoi(e) != 0;
    if (const char *e = getenv("EQGPU_RT_MIN_TILES")) s->rt_min_tiles = std::max(1, atoi(e));
    if (!s->rt_smooth) return 0;
    bool ok = set_smem(RT::k_pre_rt<3, 4>, RT_SMEM_PRE) && set_smem(RT::k_pre_rt<4, 6>, RT_SMEM_PRE) &&
              set_smem(RT::k_post_rt<3, 4, true>, RT_SMEM_POST) && set_smem(RT::k_post_rt<3, 4, false>, RT_SMEM_POST) &&
              set_smem(RT::k_post_rt<4, 4, true>, RT_SMEM_POST) && set_smem(RT::k_post_rt<4, 4, false>, RT_SMEM_POST);
    if (!ok) { s->rt_smooth = false; return 0; }
    s->rt_ctas = 2 * s->num_sms;
    if (const char *e = getenv("EQGPU_RT_CTAS")) s->rt_ctas = std::max(1, atoi(e));
    EQ_CUDA(cudaMalloc(&s->rt_sched, sizeof(unsigned) * 4));
    EQ_CUDA(cudaMemset(s->rt_sched, 0, sizeof(unsigned) * 4));
    EQ_CUDA(cudaStreamCreateWithFlags(&s->rt_stream, cudaStreamNonBlocking));
    EQ_CUDA(cudaEventCreateWithFlags(&s->ev_rt_fork, cudaEventDisableTiming));
    EQ_CUDA(cudaEventCreateWithFlags(&s->ev_rt_join, cudaEventDisableTiming));
    int max_level = 1;   // levels 0 .. max_level may use the register-tile kernels
    if (const char *e = getenv("EQGPU_RT_LEVELS")) max_level = atoi(e) - 1;
    for (size_t l = 0; l + 1 < s->levels.size() && (int)l <= max_level; ++l) {
        Level &lv = s->levels[l];
        const int nu = l == 0 ? s->nu : s->nuc;
        if (nu != 3 && nu != 4) continue;
        if (!make_tile_map(&lv.map_b64, lv.b, lv.dev.nx, lv.dev.ny) || !make_tile_map(&lv.map_t64, lv.t, lv.dev.nx, lv.dev.ny))
            continue;
        int rc = plan_level(s, lv.dev, rt_halo_pre(nu), lv.rt_pre);
        if (rc) return rc;
        rc = plan_level(s, lv.dev, rt_halo_post(nu), lv.rt_post);
        if (rc) return rc;
    }
    return 0;
}

void rt_teardown(eqgpu_solver *s)
{
    for (auto &lv : s->levels) {
        cudaFree(lv.rt_pre.d_tlist); lv.rt_pre.d_tlist = nullptr; lv.rt_pre.on = false;
        cudaFree(lv.rt_post.d_tlist); lv.rt_post.d_tlist = nullptr; lv.rt_post.on = false;
    }
    cudaFree(s->rt_sched); s->rt_sched = nullptr;
    if (s->ev_rt_fork) { cudaEventDestroy(s->ev_rt_fork); s->ev_rt_fork = nullptr; }
    if (s->ev_rt_join) { cudaEventDestroy(s->ev_rt_join); s->ev_rt_join = nullptr; }
    if (s->rt_stream) { cudaStreamDestroy(s->rt_stream); s->rt_stream = nullptr; }
}

static RT::Geom geom_of(const Level::RtPlan &P)
{
    RT::Geom G;
    G.gx = P.gx; G.gy = P.gy; G.bx0 = P.bx0; G.by0 = P.by0; G.nbx = P.nbx; G.nby = P.nby;
    return G;
}

#define RT_LAUNCH(PDL_OK, KERN, SM, ST, ...)                                                              \
    do {                                                                                                  \
        cudaLaunchConfig_t cfg_{};                                                                        \
        cfg_.gridDim = dim3(std::min(s->rt_ctas, G.nbx * G.nby)); cfg_.blockDim = dim3(RT::NT);           \
        cfg_.dynamicSmemBytes = (SM); cfg_.stream = (ST);                                                 \
        cudaLaunchAttribute at_[1];                                                                       \
        if (s->pdl && (PDL_OK) && !s->pdl_block) {                                                        \
            at_[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;                               \
            at_[0].val.programmaticStreamSerializationAllowed = 1;                                        \
            cfg_.attrs = at_; cfg_.numAttrs = 1;                                                          \
        }                                                                                                 \
        s->pdl_block = false;                                                                             \
        cudaLaunchKernelEx(&cfg_, KERN, __VA_ARGS__);                                                     \
    } while (0)

void rt_launch_pre(eqgpu_solver *s, cudaStream_t st, int l, int nu, const SmoothW &sw, bool pdl_ok)
{
    Level &lv = s->levels[l], &cv = s->levels[l + 1];
    const RT::Geom G = geom_of(lv.rt_pre);
    const CGScalars *scc = s->sc;
    if (nu == 3)
        RT_LAUNCH(pdl_ok, (RT::k_pre_rt<3, 4>), RT_SMEM_PRE, st, lv.dev, cv.dev, lv.map_b64, lv.t, cv.b, sw, G, s->rt_sched, scc);
    else
        RT_LAUNCH(pdl_ok, (RT::k_pre_rt<4, 6>), RT_SMEM_PRE, st, lv.dev, cv.dev, lv.map_b64, lv.t, cv.b, sw, G, s->rt_sched, scc);
}

void rt_launch_post(eqgpu_solver *s, cudaStream_t st, int l, int nu, const SmoothW &sw, bool dot, double *out_dot)
{
    Level &lv = s->levels[l], &cv = s->levels[l + 1];
    const RT::Geom G = geom_of(lv.rt_post);
    const double *xc = cv.x;
#define RT_POST(NU, DOT)                                                                                                  \
    RT_LAUNCH(true, (RT::k_post_rt<NU, 4, DOT>), RT_SMEM_POST, st, lv.dev, cv.dev, lv.map_b64, lv.map_t64, lv.x, xc, sw, G, \
              s->rt_sched + 2, s->sc, s->partials, s->counters + 1, out_dot)
    if (nu == 3) { if (dot) RT_POST(3, true); else RT_POST(3, false); }
    else { if (dot) RT_POST(4, true); else RT_POST(4, false); }
#undef RT_POST
}
