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
//   * a tile qualifies when every node of its 64 x 64 window is regular (constant coefficients), outside the grid (TMA fills
//     zeros) or on a Dirichlet wall: the latter two are held at zero by a per-lane column mask folded into the sweep weight
//     and a per-row bit.  Tiles that hold natural-boundary wall nodes or the narrower last cell of a coarse grid stay with
//     the general shared-memory kernels, launched on a tile list beside this kernel (solver.cu: launch_pre / launch_post);
//     with Dirichlet walls all round (the shipped trap) level 0 has none.
//   * levels with an odd row pitch (1025, 513, ...: no TMA) load their rows straight into the registers (TMA = false).
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
    int gx, gy;          // tile grid of the whole level (shared with the tile-list kernel)
    int n;               // tiles handled here ...
    const int *tiles;    // ... as (bx, by) pairs
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

// What a lane knows about its two columns and eight rows of the current tile
struct Lane {
    int ox, oy, bx, by;
    double cma, cmb;     // 1.0: free node column, 0.0: outside the grid or a Dirichlet wall column
    bool ina, inb;       // column inside the grid
    unsigned rowm;       // bit k: row k of the band is inside the grid and not a Dirichlet wall row
    unsigned rowin;      // bit k: row k of the band is inside the grid
};

template <int H>
__device__ __forceinline__ Lane lane_of(const LevelDev &F, const Geom &G, int t, int w, int lane)
{
    constexpr int TO = TS - 2 * H;
    Lane L;
    L.bx = __ldg(G.tiles + 2 * t);
    L.by = __ldg(G.tiles + 2 * t + 1);
    L.ox = L.bx * TO - H;
    L.oy = F.tbase + L.by * TO - H;
    const int ja = L.ox + 2 * lane, jb = ja + 1;
    const bool dl = F.dirmask & 1u, dr = F.dirmask & 2u, dt = F.dirmask & 4u, db = F.dirmask & 8u;
    L.ina = ja >= 0 && ja < F.nx;
    L.inb = jb >= 0 && jb < F.nx;
    L.cma = (L.ina && !(dl && ja == 0) && !(dr && ja == F.nx - 1)) ? 1.0 : 0.0;
    L.cmb = (L.inb && !(dl && jb == 0) && !(dr && jb == F.nx - 1)) ? 1.0 : 0.0;
    unsigned m = 0, in = 0;
#pragma unroll
    for (int k = 0; k < RW; ++k) {
        const int gi = L.oy + w * RW + k;
        const bool inside = gi >= 0 && gi < F.ny;
        in |= inside ? (1u << k) : 0u;
        m |= (inside && !(db && gi == 0) && !(dt && gi == F.ny - 1)) ? (1u << k) : 0u;
    }
    L.rowm = m; L.rowin = in;
    return L;
}

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

// Between two passes a warp only needs its two neighbours' boundary rows.  Pairwise named barriers (boundary q, shared by
// warps q-1 and q, hardware barrier q, 64 threads; lower boundary first, so the chain cannot deadlock) were MEASURED
// SLOWER than one CTA-wide __syncthreads on the B200 (pre-smoother 39.0 against 32.9 us, post 45.3 against 40.9 us at
// 2048^2, profiles/r02_register_tile.md): two barrier instructions per pass and a ripple of dependent releases cost more
// than the slack they give.  Kept behind RT_PAIR_BARRIER for the record.
#ifndef RT_PAIR_BARRIER
#define RT_PAIR_BARRIER 0
#endif
__device__ __forceinline__ void band_sync(int w)
{
#if RT_PAIR_BARRIER
    if (w > 0) asm volatile("bar.sync %0, 64;" ::"r"(w) : "memory");
    if (w < NWARP - 1) asm volatile("bar.sync %0, 64;" ::"r"(w + 1) : "memory");
#else
    __syncthreads();
#endif
}

// One pass over the band.  RESID false: x <- x + wd (b - A x) in place, wda / wdb = wd times the column masks;  true: the
// residual b - A x of the band's rows (times the column masks wda / wdb) goes to rows row0 .. row0+RW-1 of the 64-column
// shared array `rt`, x is left alone.  Rows whose bit in rowm is clear stay (are stored as) zero.
// Lane 0's west partial and lane 31's east partial are their own values (the shuffle has no source): the outermost columns
// of the tile belong to the ring that goes stale, one node per pass, like the outermost rows.
// MASKED false (tiles strictly inside the walls): no masks, wda is the weight of both columns.
// BSM true (the three-CTAs-per-SM instances): the right-hand side is not held in registers but read from the shared tile
// `rt` row by row (ba / bb unused), and the residual overwrites it in place.
template <bool RESID, bool MASKED, bool BSM = false>
__device__ __forceinline__ void band_pass(const Coef &c, double wda, double wdb, unsigned rowm, const double (&ba)[RW],
                                          const double (&bb)[RW], double (&xa)[RW + 2], double (&xb)[RW + 2], double *rt,
                                          int row0, int lane)
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
        double bka, bkb;
        if (BSM) {
            const double2 bv = *reinterpret_cast<const double2 *>(rt + (row0 + k - 1) * TS + 2 * lane);
            bka = bv.x; bkb = bv.y;
        } else {
            bka = ba[k - 1]; bkb = bb[k - 1];
        }
        const double ra = bka - axa, rb = bkb - axb;
        if (!MASKED) {
            if (RESID) {
                *reinterpret_cast<double2 *>(rt + (row0 + k - 1) * TS + 2 * lane) = make_double2(ra, rb);
            } else {
                xa[k] = fma(wda, ra, ca);
                xb[k] = fma(wda, rb, cb);
            }
        } else {
            const bool ok = (rowm >> (k - 1)) & 1u;
            if (RESID) {
                *reinterpret_cast<double2 *>(rt + (row0 + k - 1) * TS + 2 * lane) =
                    make_double2(ok ? wda * ra : 0.0, ok ? wdb * rb : 0.0);
            } else {
                xa[k] = ok ? fma(wda, ra, ca) : 0.0;
                xb[k] = ok ? fma(wdb, rb, cb) : 0.0;
            }
        }
        oa = ca; ob = cb;
    }
}

// the band's rows of a field, straight from global memory (levels without a TMA descriptor); zeros outside the grid
__device__ __forceinline__ void band_load(const LevelDev &F, const Lane &L, const double *__restrict__ g, int w, int lane,
                                          double (&va)[RW], double (&vb)[RW])
{
    const int ja = L.ox + 2 * lane;
#pragma unroll
    for (int k = 0; k < RW; ++k) {
        const bool in = (L.rowin >> k) & 1u;
        const double *q = g + (size_t)(L.oy + w * RW + k) * F.nx + ja;
        va[k] = (in && L.ina) ? __ldg(q) : 0.0;
        vb[k] = (in && L.inb) ? __ldg(q + 1) : 0.0;
    }
}

// owned nodes of the band to global memory (xa / xb rows 1 .. RW)
template <int H, bool TMA>
__device__ __forceinline__ void band_store(const LevelDev &F, const Lane &L, double *__restrict__ x, int w, int lane,
                                           const double (&xa)[RW + 2], const double (&xb)[RW + 2])
{
    if (lane < H / 2 || lane >= 32 - H / 2) return;
#pragma unroll
    for (int k = 0; k < RW; ++k) {
        const int ly = w * RW + k;
        if (ly < H || ly >= TS - H || !((L.rowin >> k) & 1u)) continue;
        double *q = x + (size_t)(L.oy + ly) * F.nx + L.ox + 2 * lane;
        if (TMA) {   // even pitch, even origin: the pair is inside the grid or outside it together, and 16-byte aligned
            if (L.ina) *reinterpret_cast<double2 *>(q) = make_double2(xa[k + 1], xb[k + 1]);
        } else {
            if (L.ina) q[0] = xa[k + 1];
            if (L.inb) q[1] = xb[k + 1];
        }
    }
}

// persistent-CTA tile scheduler: ids are drawn from sched[0]; the last CTA to leave (sched[1]) resets both
__device__ __forceinline__ void sched_leave(unsigned *sched)
{
    if (threadIdx.x == 0) {
        const unsigned prev = atomicAdd(sched + 1, 1u);
        if (prev == gridDim.x - 1) { sched[0] = 0u; sched[1] = 0u; }
    }
}

// the sweeps of one pre-smoothing tile: x = w0 D^-1 b, NU - 1 passes, the residual into rt
template <int NU, bool MASKED>
__device__ __forceinline__ void pre_sweeps(const Coef &c, const SmoothW &sw, double icC, const Lane &L, double *exch, int &buf,
                                           int w, int lane, const double (&ba)[RW], const double (&bb)[RW],
                                           double (&xa)[RW + 2], double (&xb)[RW + 2], double *rt)
{
    const double ma = MASKED ? L.cma : 1.0, mb = MASKED ? L.cmb : 1.0;
    {
        const double wda = sw.w[0] * icC * ma, wdb = sw.w[0] * icC * mb;
#pragma unroll
        for (int k = 0; k < RW; ++k) {
            const bool ok = !MASKED || ((L.rowm >> k) & 1u);
            xa[k + 1] = ok ? wda * ba[k] : 0.0;
            xb[k + 1] = ok ? wdb * bb[k] : 0.0;
        }
    }
#pragma unroll
    for (int s = 1; s < NU; ++s) {
        band_publish(exch, buf, w, lane, xa, xb);
        band_sync(w);
        band_halo(exch, buf, w, lane, xa, xb);
        buf ^= 1;
        band_pass<false, MASKED>(c, sw.w[s] * icC * ma, sw.w[s] * icC * mb, L.rowm, ba, bb, xa, xb, nullptr, 0, lane);
    }
    band_publish(exch, buf, w, lane, xa, xb);
    band_sync(w);
    band_halo(exch, buf, w, lane, xa, xb);
    buf ^= 1;
    band_pass<true, MASKED>(c, ma, mb, L.rowm, ba, bb, xa, xb, rt, w * RW, lane);
}

// =====================================================================================================================
// pre-smoothing: NU sweeps from a zero guess, residual, restriction.        reads b        writes x, b_coarse
// =====================================================================================================================
template <int NU, int H, bool TMA>
__global__ void __launch_bounds__(NT, 2)
k_pre_rt(const LevelDev F, const LevelDev Cc, const __grid_constant__ CUtensorMap map_b, const double *__restrict__ b,
         double *__restrict__ x, double *__restrict__ bc, const SmoothW sw, const Geom G, unsigned *sched,
         const CGScalars *sc)
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
    const int ntiles = G.n;
    // The converged flag is read before the dependency wait (its writers have completed before any prologue that can
    // overlap runs: see k_presmooth in mg_tile.inc).  The first two tiles of a CTA are static (blockIdx.x and blockIdx.x +
    // gridDim.x), so the first TMA copy goes out the moment the predecessor's data is visible; the shared counter hands
    // out the tiles from 2 * gridDim.x on.
    const int done = sc->done;
    if (TMA && tid == 0) { mbar_init(bar, 1); fence_barrier_init(); }
    int cur = blockIdx.x, nxt = blockIdx.x + gridDim.x;
    pdl_wait();
    if (TMA && tid == 0 && cur < ntiles && !done) {
        mbar_expect_tx(bar, TS * TS * 8);
        tma_load_2d(bbuf, &map_b, bar, __ldg(G.tiles + 2 * cur) * TO - H, F.tbase + __ldg(G.tiles + 2 * cur + 1) * TO - H);
    }
    if (done) return;
    __syncthreads();   // the barrier's initialisation is visible to every thread
    unsigned phase = 0;
    int it = 0, buf = 0;
    const Coef c{F.cC, F.cEW, F.cNS, F.cD};
    const double icC = F.icC;
    while (cur < ntiles) {
        const Lane L = lane_of<H>(F, G, cur, w, lane);
        double ba[RW], bb[RW], xa[RW + 2], xb[RW + 2];
        if (TMA) {
            mbar_wait(bar, phase);
            phase ^= 1u;
#pragma unroll
            for (int k = 0; k < RW; ++k) {
                const double2 v = *reinterpret_cast<const double2 *>(bbuf + (w * RW + k) * TS + 2 * lane);
                ba[k] = v.x; bb[k] = v.y;
            }
        } else {
            band_load(F, L, b, w, lane, ba, bb);
        }
        __syncthreads();   // everyone holds its rows: the box may be refilled (and the previous tile's restriction is over)
        unsigned t_nn = 0;
        if (tid == 0) {
            t_nn = 2u * gridDim.x + atomicAdd(sched, 1u);   // the tile after the next one; consumed at the end of this tile
            if (TMA && nxt < ntiles) {
                fence_proxy_async();
                mbar_expect_tx(bar, TS * TS * 8);
                tma_load_2d(bbuf, &map_b, bar, __ldg(G.tiles + 2 * nxt) * TO - H, F.tbase + __ldg(G.tiles + 2 * nxt + 1) * TO - H);
            }
        }
        // (tiles strictly inside the walls take the unmasked instance: a CTA-uniform branch)
        if (L.ox >= 1 && L.ox + TS <= F.nx - 1 && L.oy >= 1 && L.oy + TS <= F.ny - 1)
            pre_sweeps<NU, false>(c, sw, icC, L, exch, buf, w, lane, ba, bb, xa, xb, rt);
        else
            pre_sweeps<NU, true>(c, sw, icC, L, exch, buf, w, lane, ba, bb, xa, xb, rt);
        if (tid == 0) ids[it & 1] = (int)t_nn;
        band_store<H, TMA>(F, L, x, w, lane, xa, xb);   // pre-smoothed iterate of the owned region
        __syncthreads();   // residual tile complete (and ids[] visible)
        // restriction (P^T, full weighting on the six mesh neighbours): coarse nodes = even fine nodes of the owned region
        // (masked residuals are zero, so walls need no special weights; a Dirichlet coarse node gets zero)
        {
            const int I0 = (L.oy + H) >> 1, J0 = (L.ox + H) >> 1;
            for (int q = tid; q < CT * CT; q += NT) {
                const int cy = q / CT, cx = q - cy * CT;
                const int gi = L.oy + H + 2 * cy, gj = L.ox + H + 2 * cx;
                if (gi >= F.ny || gj >= F.nx) continue;
                const int cc = (H + 2 * cy) * TS + H + 2 * cx;
                const double h = rt[cc + 1] + rt[cc - 1] + rt[cc + TS] + rt[cc - TS] + rt[cc + TS + 1] + rt[cc - TS - 1];
                bc[(size_t)(I0 + cy) * Cc.nx + J0 + cx] = is_dirichlet(Cc, I0 + cy, J0 + cx) ? 0.0 : rt[cc] + 0.5 * h;
            }
            // even node counts: the last fine node is odd and has a coarse node of its own, a Dirichlet one here
            const bool lastx = !(F.nx & 1) && L.ox + TS - H >= F.nx, lasty = !(F.ny & 1) && L.oy + TS - H >= F.ny;
            if (lastx)
                for (int cy = tid; cy < CT; cy += NT)
                    if (L.oy + H + 2 * cy < F.ny) bc[(size_t)(I0 + cy) * Cc.nx + Cc.nx - 1] = 0.0;
            if (lasty)
                for (int cx = tid; cx < CT; cx += NT)
                    if (L.ox + H + 2 * cx < F.nx) bc[(size_t)(Cc.ny - 1) * Cc.nx + J0 + cx] = 0.0;
            if (lastx && lasty && tid == 0) bc[(size_t)Cc.ny * Cc.nx - 1] = 0.0;
        }
        const int nn = ids[it & 1];
        cur = nxt; nxt = nn;
        ++it;
        // (the next tile's first write to rt comes after two more __syncthreads: no barrier needed here)
    }
    sched_leave(sched);
}

// =====================================================================================================================
// post-smoothing: x = xin + P x_coarse, NU sweeps (+ x.b).        reads b, xin, x_coarse        writes x
// =====================================================================================================================
// Coarse patch under a tile: PATCH x PATCH coarse nodes from (oy >> 1, ox >> 1), clamped to the coarse grid (values that
// belong to nodes outside the fine grid are masked later), copied global -> shared asynchronously (8-byte LDGSTS: the
// coarse pitch is odd), so the next tile's patch costs no registers while the current tile is swept.
__device__ __forceinline__ void patch_copy_async(const LevelDev &Cc, const double *__restrict__ xc, int ox, int oy, int tid,
                                                 double *patch)
{
    const int J0 = ox >> 1, I0 = oy >> 1;
#pragma unroll
    for (int q = 0; q < 5; ++q) {
        const int e = tid + q * NT;
        const int ci = e / PATCH, cj = e - ci * PATCH;
        if (e < PATCH * PATCH)
            cp_async8(patch + ci * PS + cj,
                      xc + (size_t)min(max(I0 + ci, 0), Cc.ny - 1) * Cc.nx + min(max(J0 + cj, 0), Cc.nx - 1));
    }
    cp_async_commit();
}

// prolongation: tile rows / columns are even <=> coincident with a coarse node (the tile origin is even; the odd last node
// of an even-sized grid is a Dirichlet node here and masked)
template <bool MASKED>
__device__ __forceinline__ void post_prolong(const Lane &L, const double *patch, int w, int lane, double (&xa)[RW + 2],
                                             double (&xb)[RW + 2])
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
        double pa, pb;
        if (k & 1) {   // midpoint row
            pa = 0.5 * (p0[j] + p0[j + 1]);
            pb = 0.5 * (p0[j] + p1[j + 1]);
        } else {
            pa = 0.5 * (p0[j] + p0[j]);
            pb = 0.5 * (p0[j] + p1[j]);
        }
        if (MASKED) {
            const bool ok = (L.rowm >> k) & 1u;
            xa[k + 1] = ok ? L.cma * (xa[k + 1] + pa) : 0.0;
            xb[k + 1] = ok ? L.cmb * (xb[k + 1] + pb) : 0.0;
        } else {
            xa[k + 1] += pa;
            xb[k + 1] += pb;
        }
    }
}

template <int NU, bool MASKED>
__device__ __forceinline__ void post_sweeps(const Coef &c, const SmoothW &sw, double icC, const Lane &L, double *exch, int &buf,
                                            int w, int lane, const double (&ba)[RW], const double (&bb)[RW],
                                            double (&xa)[RW + 2], double (&xb)[RW + 2])
{
    const double ma = MASKED ? L.cma : 1.0, mb = MASKED ? L.cmb : 1.0;
#pragma unroll
    for (int s = 0; s < NU; ++s) {
        band_publish(exch, buf, w, lane, xa, xb);
        band_sync(w);
        band_halo(exch, buf, w, lane, xa, xb);
        buf ^= 1;
        band_pass<false, MASKED>(c, sw.w[s] * icC * ma, sw.w[s] * icC * mb, L.rowm, ba, bb, xa, xb, nullptr, 0, lane);
    }
}

template <int NU, int H, bool DOT, bool TMA>
__global__ void __launch_bounds__(NT, 2)
k_post_rt(const LevelDev F, const LevelDev Cc, const __grid_constant__ CUtensorMap map_b,
          const __grid_constant__ CUtensorMap map_x, const double *__restrict__ b, const double *__restrict__ xin,
          double *__restrict__ x, const double *__restrict__ xc, const SmoothW sw, const Geom G, unsigned *sched,
          CGScalars *sc, double *partials, unsigned *counter, double *out_dot)
{
    static_assert(H % 2 == 0 && H >= NU, "even halo of at least NU nodes");
    constexpr int TO = TS - 2 * H;
    extern __shared__ __align__(128) unsigned char smraw[];
    double *bbuf = reinterpret_cast<double *>(smraw);   // TMA boxes: 64 x 64 each
    double *xbuf = bbuf + TS * TS;
    double *exch = xbuf + TS * TS;                      // 2 x NWARP x 2 x 64
    double *patch2 = exch + 2 * NWARP * 2 * TS;         // two buffers of PATCH rows of stride PS
    double *red = patch2 + 2 * PATCH * PS;              // 32 doubles
    unsigned long long *bar = reinterpret_cast<unsigned long long *>(red + 32);
    int *ids = reinterpret_cast<int *>(bar + 1);
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int ntiles = G.n;
    // (prologue as in k_pre_rt: flag before the dependency wait, static first tiles, first copies at once)
    const int done = sc->done;
    if (TMA && tid == 0) { mbar_init(bar, 1); fence_barrier_init(); }
    int cur = blockIdx.x, nxt = blockIdx.x + gridDim.x;
    pdl_wait();
    if (done) return;
    if (cur < ntiles) {
        const int ox = __ldg(G.tiles + 2 * cur) * TO - H, oy = F.tbase + __ldg(G.tiles + 2 * cur + 1) * TO - H;
        if (TMA && tid == 0) {
            mbar_expect_tx(bar, 2 * TS * TS * 8);
            tma_load_2d(bbuf, &map_b, bar, ox, oy);
            tma_load_2d(xbuf, &map_x, bar, ox, oy);
        }
        patch_copy_async(Cc, xc, ox, oy, tid, patch2);   // the first tile's coarse patch
        cp_async_wait<0>();
    }
    __syncthreads();   // the barrier's initialisation is visible to every thread
    unsigned phase = 0, mine = 0;
    int it = 0, buf = 0;
    const Coef c{F.cC, F.cEW, F.cNS, F.cD};
    const double icC = F.icC;
    while (cur < ntiles) {
        const Lane L = lane_of<H>(F, G, cur, w, lane);
        double ba[RW], bb[RW], xa[RW + 2], xb[RW + 2];
        if (TMA) {
            mbar_wait(bar, phase);
            phase ^= 1u;
#pragma unroll
            for (int k = 0; k < RW; ++k) {
                const double2 v = *reinterpret_cast<const double2 *>(bbuf + (w * RW + k) * TS + 2 * lane);
                const double2 u = *reinterpret_cast<const double2 *>(xbuf + (w * RW + k) * TS + 2 * lane);
                ba[k] = v.x; bb[k] = v.y; xa[k + 1] = u.x; xb[k + 1] = u.y;
            }
        } else {
            double ta[RW], tb[RW];
            band_load(F, L, b, w, lane, ba, bb);
            band_load(F, L, xin, w, lane, ta, tb);
#pragma unroll
            for (int k = 0; k < RW; ++k) { xa[k + 1] = ta[k]; xb[k + 1] = tb[k]; }
        }
        __syncthreads();   // boxes consumed, and the patch written at the end of the previous tile (or above) is visible
        unsigned t_nn = 0;
        if (tid == 0) {
            t_nn = 2u * gridDim.x + atomicAdd(sched, 1u);
            if (TMA && nxt < ntiles) {
                const int ox2 = __ldg(G.tiles + 2 * nxt) * TO - H, oy2 = F.tbase + __ldg(G.tiles + 2 * nxt + 1) * TO - H;
                fence_proxy_async();
                mbar_expect_tx(bar, 2 * TS * TS * 8);
                tma_load_2d(bbuf, &map_b, bar, ox2, oy2);
                tma_load_2d(xbuf, &map_x, bar, ox2, oy2);
            }
        }
        const double *patch = patch2 + (it & 1) * PATCH * PS;
        const bool inner = L.ox >= 1 && L.ox + TS <= F.nx - 1 && L.oy >= 1 && L.oy + TS <= F.ny - 1;   // CTA-uniform
        if (inner) post_prolong<false>(L, patch, w, lane, xa, xb);
        else post_prolong<true>(L, patch, w, lane, xa, xb);
        // the next tile's patch into the other buffer (last read during the previous tile): in flight during the sweeps
        if (nxt < ntiles)
            patch_copy_async(Cc, xc, __ldg(G.tiles + 2 * nxt) * TO - H, F.tbase + __ldg(G.tiles + 2 * nxt + 1) * TO - H, tid,
                             patch2 + ((it + 1) & 1) * PATCH * PS);
        if (inner) post_sweeps<NU, false>(c, sw, icC, L, exch, buf, w, lane, ba, bb, xa, xb);
        else post_sweeps<NU, true>(c, sw, icC, L, exch, buf, w, lane, ba, bb, xa, xb);
        cp_async_wait<0>();   // my share of the next patch has landed; the barrier below publishes it
        if (tid == 0) ids[it & 1] = (int)t_nn;
        band_store<H, TMA>(F, L, x, w, lane, xa, xb);
        if (DOT) {
            double acc = 0.0;
            if (lane >= H / 2 && lane < 32 - H / 2) {
#pragma unroll
                for (int k = 0; k < RW; ++k) {
                    const int ly = w * RW + k;
                    if (ly >= H && ly < TS - H) acc += xa[k + 1] * ba[k] + xb[k + 1] * bb[k];   // zero outside the grid
                }
            }
            const double t = cta_sum(acc, red);   // contains the __syncthreads that publishes ids[] and the patch
            if (tid == 0) partials[L.by * G.gx + L.bx] = t;
            ++mine;
        } else {
            __syncthreads();
        }
        const int nn = ids[it & 1];
        cur = nxt; nxt = nn;
        ++it;
    }
    if (DOT && mine) tiles_arrive(partials, counter, mine, (unsigned)(G.gx * G.gy), out_dot);
    sched_leave(sched);
}

// =====================================================================================================================
// Three CTAs per SM ("lean" instances, TMA levels): the right-hand side tile stays in shared memory for the whole tile
// (read row by row in every pass: 8 LDS.128 per lane and pass) instead of in 32 registers per thread, which brings the
// kernels under 80 registers -- 24 warps per SM instead of 16.  The sweeps are latency-bound at four warps per scheduler
// (profiles/r02_ncu_register_tile.md), so more resident warps buy more than the extra shared-memory reads cost.  The
// price: one box buffer per CTA, so the next tile's copy can only start when the current tile is done with it (no
// prefetch; the other two CTAs of the SM cover that wait), and the post-smoother's incoming iterate comes by 128-bit
// global loads straight into registers.  The residual overwrites the right-hand side in place.
// =====================================================================================================================
template <int NU, bool MASKED>
__device__ __forceinline__ void pre_sweeps3(const Coef &c, const SmoothW &sw, double icC, const Lane &L, double *exch, int &buf,
                                            int w, int lane, double (&xa)[RW + 2], double (&xb)[RW + 2], double *bt)
{
    const double ma = MASKED ? L.cma : 1.0, mb = MASKED ? L.cmb : 1.0;
    const double dummy[RW] = {};
    {
        const double wda = sw.w[0] * icC * ma, wdb = sw.w[0] * icC * mb;
#pragma unroll
        for (int k = 0; k < RW; ++k) {
            const double2 bv = *reinterpret_cast<const double2 *>(bt + (w * RW + k) * TS + 2 * lane);
            const bool ok = !MASKED || ((L.rowm >> k) & 1u);
            xa[k + 1] = ok ? wda * bv.x : 0.0;
            xb[k + 1] = ok ? wdb * bv.y : 0.0;
        }
    }
#pragma unroll
    for (int s = 1; s < NU; ++s) {
        band_publish(exch, buf, w, lane, xa, xb);
        band_sync(w);
        band_halo(exch, buf, w, lane, xa, xb);
        buf ^= 1;
        band_pass<false, MASKED, true>(c, sw.w[s] * icC * ma, sw.w[s] * icC * mb, L.rowm, dummy, dummy, xa, xb, bt, w * RW, lane);
    }
    band_publish(exch, buf, w, lane, xa, xb);
    band_sync(w);
    band_halo(exch, buf, w, lane, xa, xb);
    buf ^= 1;
    band_pass<true, MASKED, true>(c, ma, mb, L.rowm, dummy, dummy, xa, xb, bt, w * RW, lane);
}

template <int NU, int H>
__global__ void __launch_bounds__(NT, 3)
k_pre_rt3(const LevelDev F, const LevelDev Cc, const __grid_constant__ CUtensorMap map_b, double *__restrict__ x,
          double *__restrict__ bc, const SmoothW sw, const Geom G, unsigned *sched, const CGScalars *sc)
{
    static_assert(H % 2 == 0 && H >= NU + 1, "even halo of at least NU + 1 nodes");
    constexpr int TO = TS - 2 * H, CT = TO / 2;
    extern __shared__ __align__(128) unsigned char smraw[];
    double *bt = reinterpret_cast<double *>(smraw);     // TMA box: right-hand side, then the residual in place
    double *exch = bt + TS * TS;                        // 2 x NWARP x 2 x 64
    unsigned long long *bar = reinterpret_cast<unsigned long long *>(exch + 2 * NWARP * 2 * TS);
    int *ids = reinterpret_cast<int *>(bar + 1);
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int ntiles = G.n;
    const int done = sc->done;
    if (tid == 0) { mbar_init(bar, 1); fence_barrier_init(); }
    int cur = blockIdx.x, nxt = blockIdx.x + gridDim.x;
    pdl_wait();
    if (tid == 0 && cur < ntiles && !done) {
        mbar_expect_tx(bar, TS * TS * 8);
        tma_load_2d(bt, &map_b, bar, __ldg(G.tiles + 2 * cur) * TO - H, F.tbase + __ldg(G.tiles + 2 * cur + 1) * TO - H);
    }
    if (done) return;
    __syncthreads();
    unsigned phase = 0;
    int it = 0, buf = 0;
    const Coef c{F.cC, F.cEW, F.cNS, F.cD};
    const double icC = F.icC;
    while (cur < ntiles) {
        const Lane L = lane_of<H>(F, G, cur, w, lane);
        unsigned t_nn = 0;
        if (tid == 0) t_nn = 2u * gridDim.x + atomicAdd(sched, 1u);   // the tile after the next one
        double xa[RW + 2], xb[RW + 2];
        mbar_wait(bar, phase);
        phase ^= 1u;
        if (L.ox >= 1 && L.ox + TS <= F.nx - 1 && L.oy >= 1 && L.oy + TS <= F.ny - 1)
            pre_sweeps3<NU, false>(c, sw, icC, L, exch, buf, w, lane, xa, xb, bt);
        else
            pre_sweeps3<NU, true>(c, sw, icC, L, exch, buf, w, lane, xa, xb, bt);
        if (tid == 0) ids[it & 1] = (int)t_nn;
        band_store<H, true>(F, L, x, w, lane, xa, xb);
        __syncthreads();   // residual tile complete (and ids[] visible)
        {
            const int I0 = (L.oy + H) >> 1, J0 = (L.ox + H) >> 1;
            for (int q = tid; q < CT * CT; q += NT) {
                const int cy = q / CT, cx = q - cy * CT;
                const int gi = L.oy + H + 2 * cy, gj = L.ox + H + 2 * cx;
                if (gi >= F.ny || gj >= F.nx) continue;
                const int cc = (H + 2 * cy) * TS + H + 2 * cx;
                const double h = bt[cc + 1] + bt[cc - 1] + bt[cc + TS] + bt[cc - TS] + bt[cc + TS + 1] + bt[cc - TS - 1];
                bc[(size_t)(I0 + cy) * Cc.nx + J0 + cx] = is_dirichlet(Cc, I0 + cy, J0 + cx) ? 0.0 : bt[cc] + 0.5 * h;
            }
            const bool lastx = !(F.nx & 1) && L.ox + TS - H >= F.nx, lasty = !(F.ny & 1) && L.oy + TS - H >= F.ny;
            if (lastx)
                for (int cy = tid; cy < CT; cy += NT)
                    if (L.oy + H + 2 * cy < F.ny) bc[(size_t)(I0 + cy) * Cc.nx + Cc.nx - 1] = 0.0;
            if (lasty)
                for (int cx = tid; cx < CT; cx += NT)
                    if (L.ox + H + 2 * cx < F.nx) bc[(size_t)(Cc.ny - 1) * Cc.nx + J0 + cx] = 0.0;
            if (lastx && lasty && tid == 0) bc[(size_t)Cc.ny * Cc.nx - 1] = 0.0;
        }
        const int nn = ids[it & 1];
        __syncthreads();   // the box has been read to the end: refill it
        if (tid == 0 && nxt < ntiles) {
            fence_proxy_async();
            mbar_expect_tx(bar, TS * TS * 8);
            tma_load_2d(bt, &map_b, bar, __ldg(G.tiles + 2 * nxt) * TO - H, F.tbase + __ldg(G.tiles + 2 * nxt + 1) * TO - H);
        }
        cur = nxt; nxt = nn;
        ++it;
    }
    sched_leave(sched);
}

template <int NU, int H, bool DOT>
__global__ void __launch_bounds__(NT, 3)
k_post_rt3(const LevelDev F, const LevelDev Cc, const __grid_constant__ CUtensorMap map_b, const double *__restrict__ xin,
           double *__restrict__ x, const double *__restrict__ xc, const SmoothW sw, const Geom G, unsigned *sched,
           CGScalars *sc, double *partials, unsigned *counter, double *out_dot)
{
    static_assert(H % 2 == 0 && H >= NU, "even halo of at least NU nodes");
    constexpr int TO = TS - 2 * H;
    extern __shared__ __align__(128) unsigned char smraw[];
    double *bt = reinterpret_cast<double *>(smraw);     // TMA box: right-hand side
    double *exch = bt + TS * TS;                        // 2 x NWARP x 2 x 64
    double *patch2 = exch + 2 * NWARP * 2 * TS;         // two buffers of PATCH rows of stride PS
    double *red = patch2 + 2 * PATCH * PS;              // 32 doubles
    unsigned long long *bar = reinterpret_cast<unsigned long long *>(red + 32);
    int *ids = reinterpret_cast<int *>(bar + 1);
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int ntiles = G.n;
    const int done = sc->done;
    if (tid == 0) { mbar_init(bar, 1); fence_barrier_init(); }
    int cur = blockIdx.x, nxt = blockIdx.x + gridDim.x;
    pdl_wait();
    if (done) return;
    if (cur < ntiles) {
        const int ox = __ldg(G.tiles + 2 * cur) * TO - H, oy = F.tbase + __ldg(G.tiles + 2 * cur + 1) * TO - H;
        if (tid == 0) {
            mbar_expect_tx(bar, TS * TS * 8);
            tma_load_2d(bt, &map_b, bar, ox, oy);
        }
        patch_copy_async(Cc, xc, ox, oy, tid, patch2);
        cp_async_wait<0>();
    }
    __syncthreads();
    unsigned phase = 0, mine = 0;
    int it = 0, buf = 0;
    const Coef c{F.cC, F.cEW, F.cNS, F.cD};
    const double icC = F.icC;
    const double dummy[RW] = {};
    while (cur < ntiles) {
        const Lane L = lane_of<H>(F, G, cur, w, lane);
        unsigned t_nn = 0;
        if (tid == 0) t_nn = 2u * gridDim.x + atomicAdd(sched, 1u);
        // the incoming iterate: straight into the registers (even pitch and origin: 128-bit, the pair inside the grid or
        // outside it together); in flight while the right-hand side box lands
        double xa[RW + 2], xb[RW + 2];
#pragma unroll
        for (int k = 0; k < RW; ++k) {
            double2 u = make_double2(0.0, 0.0);
            if (((L.rowin >> k) & 1u) && L.ina)
                u = *reinterpret_cast<const double2 *>(xin + (size_t)(L.oy + w * RW + k) * F.nx + L.ox + 2 * lane);
            xa[k + 1] = u.x; xb[k + 1] = u.y;
        }
        const double *patch = patch2 + (it & 1) * PATCH * PS;
        const bool inner = L.ox >= 1 && L.ox + TS <= F.nx - 1 && L.oy >= 1 && L.oy + TS <= F.ny - 1;   // CTA-uniform
        if (inner) post_prolong<false>(L, patch, w, lane, xa, xb);
        else post_prolong<true>(L, patch, w, lane, xa, xb);
        if (nxt < ntiles)   // the next tile's patch into the other buffer (last read during the previous tile)
            patch_copy_async(Cc, xc, __ldg(G.tiles + 2 * nxt) * TO - H, F.tbase + __ldg(G.tiles + 2 * nxt + 1) * TO - H, tid,
                             patch2 + ((it + 1) & 1) * PATCH * PS);
        mbar_wait(bar, phase);
        phase ^= 1u;
        const double ma = inner ? 1.0 : L.cma, mb = inner ? 1.0 : L.cmb;
#pragma unroll
        for (int s = 0; s < NU; ++s) {
            band_publish(exch, buf, w, lane, xa, xb);
            band_sync(w);
            band_halo(exch, buf, w, lane, xa, xb);
            buf ^= 1;
            if (inner) band_pass<false, false, true>(c, sw.w[s] * icC, sw.w[s] * icC, L.rowm, dummy, dummy, xa, xb, bt, w * RW, lane);
            else band_pass<false, true, true>(c, sw.w[s] * icC * ma, sw.w[s] * icC * mb, L.rowm, dummy, dummy, xa, xb, bt, w * RW, lane);
        }
        cp_async_wait<0>();
        if (tid == 0) ids[it & 1] = (int)t_nn;
        band_store<H, true>(F, L, x, w, lane, xa, xb);
        double acc = 0.0;
        if (DOT && lane >= H / 2 && lane < 32 - H / 2) {
#pragma unroll
            for (int k = 0; k < RW; ++k) {
                const int ly = w * RW + k;
                if (ly >= H && ly < TS - H) {
                    const double2 bv = *reinterpret_cast<const double2 *>(bt + ly * TS + 2 * lane);
                    acc += xa[k + 1] * bv.x + xb[k + 1] * bv.y;
                }
            }
        }
        if (DOT) {
            const double t = cta_sum(acc, red);   // contains the __syncthreads after which the box may be refilled
            if (tid == 0) partials[L.by * G.gx + L.bx] = t;
            ++mine;
        } else {
            __syncthreads();
        }
        const int nn = ids[it & 1];
        if (tid == 0 && nxt < ntiles) {
            fence_proxy_async();
            mbar_expect_tx(bar, TS * TS * 8);
            tma_load_2d(bt, &map_b, bar, __ldg(G.tiles + 2 * nxt) * TO - H, F.tbase + __ldg(G.tiles + 2 * nxt + 1) * TO - H);
        }
        cur = nxt; nxt = nn;
        ++it;
    }
    if (DOT && mine) tiles_arrive(partials, counter, mine, (unsigned)(G.gx * G.gy), out_dot);
    sched_leave(sched);
}

}  // namespace RT

// ---------------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// TMA descriptor of an nx x ny fp64 field (row-major, pitch nx) with a 64 x 64 box; elements outside the field arrive as
// zeros.  False when the driver entry point is missing or the pitch is not a multiple of 16 bytes (odd nx).
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
static const size_t RT_SMEM_PRE3 = (size_t)(RT::TS * RT::TS + 2 * RT::NWARP * 2 * RT::TS) * 8 + 64;
static const size_t RT_SMEM_POST3 = (size_t)(RT::TS * RT::TS + 2 * RT::NWARP * 2 * RT::TS + 2 * RT::PATCH * RT::PS + 32) * 8 + 64;
static const size_t RT_SMEM_POST = (size_t)(2 * RT::TS * RT::TS + 2 * RT::NWARP * 2 * RT::TS + 2 * RT::PATCH * RT::PS + 32) * 8 + 64;

// Splits the gx x gy tiling (halo h) of a level into the tiles the register-tile kernel takes -- every node of the 64 x 64
// window regular, outside the grid, or on a Dirichlet wall -- and the rest (tile list of the general kernel).
static int plan_level(eqgpu_solver *s, const LevelDev &F, int h, Level::RtPlan &P)
{
    const int to = RT::TS - 2 * h;
    P.on = false;
    P.gx = (F.nx + to - 1) / to;
    P.gy = (F.ny + to - 1) / to;
    const bool dl = F.dirmask & 1u, dr = F.dirmask & 2u, dt = F.dirmask & 4u, db = F.dirmask & 8u;
    auto line_ok = [&](int o, int n, int reg_hi, bool dlo, bool dhi) {
        for (int q = std::max(o, 0); q <= std::min(o + RT::TS - 1, n - 1); ++q) {
            const bool regular = q >= 1 && q <= reg_hi;
            const bool dirichlet = (q == 0 && dlo) || (q == n - 1 && dhi);
            if (!regular && !dirichlet) return false;
        }
        return true;
    };
    std::vector<int> mine, rest;
    for (int by = 0; by < P.gy; ++by)
        for (int bx = 0; bx < P.gx; ++bx) {
            const bool ok = line_ok(bx * to - h, F.nx, F.jreg_hi, dl, dr) && line_ok(by * to - h, F.ny, F.ireg_hi, db, dt);
            std::vector<int> &v = ok ? mine : rest;
            v.push_back(bx); v.push_back(by);
        }
    P.n = (int)mine.size() / 2;
    P.nperim = (int)rest.size() / 2;
    if (P.n < s->rt_min_tiles) return 0;
    EQ_CUDA(cudaMalloc(&P.d_tiles, sizeof(int) * mine.size()));
    EQ_CUDA(cudaMemcpy(P.d_tiles, mine.data(), sizeof(int) * mine.size(), cudaMemcpyHostToDevice));
    if (P.nperim > 0) {
        EQ_CUDA(cudaMalloc(&P.d_tlist, sizeof(int) * rest.size()));
        EQ_CUDA(cudaMemcpy(P.d_tlist, rest.data(), sizeof(int) * rest.size(), cudaMemcpyHostToDevice));
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
    s->rt_smooth = !s->slab && s->fused;
    if (const char *e = getenv("EQGPU_RT")) s->rt_smooth = s->rt_smooth && atoi(e) != 0;
    if (const char *e = getenv("EQGPU_RT_MIN_TILES")) s->rt_min_tiles = std::max(1, atoi(e));
    if (!s->rt_smooth) return 0;
    bool ok = true;
#define RT_SET(TMA)                                                                                                        \
    ok = ok && set_smem(RT::k_pre_rt<3, 4, TMA>, RT_SMEM_PRE) && set_smem(RT::k_pre_rt<4, 6, TMA>, RT_SMEM_PRE) &&         \
         set_smem(RT::k_post_rt<3, 4, true, TMA>, RT_SMEM_POST) && set_smem(RT::k_post_rt<3, 4, false, TMA>, RT_SMEM_POST) && \
         set_smem(RT::k_post_rt<4, 4, true, TMA>, RT_SMEM_POST) && set_smem(RT::k_post_rt<4, 4, false, TMA>, RT_SMEM_POST)
    RT_SET(true);
    RT_SET(false);
#undef RT_SET
    if (!ok) { s->rt_smooth = false; return 0; }
    // three CTAs per SM: the lean instances (right-hand side in shared memory, under 80 registers) on TMA levels.  Opt-in
    // (EQGPU_RT_LEAN=1): measured on the B200 at 2048^2, 24 warps per SM buy nothing -- 29.4 / 41.7 us against 28.0 / 37.0 us
    // for the two-CTA instances with their prefetched boxes (profiles/r02_ncu_register_tile.md)
    s->rt_lean = getenv("EQGPU_RT_LEAN") != nullptr && atoi(getenv("EQGPU_RT_LEAN")) != 0;
    if (s->rt_lean)
        s->rt_lean = set_smem(RT::k_pre_rt3<3, 4>, RT_SMEM_PRE3) && set_smem(RT::k_pre_rt3<4, 6>, RT_SMEM_PRE3) &&
                     set_smem(RT::k_post_rt3<3, 4, true>, RT_SMEM_POST3) && set_smem(RT::k_post_rt3<3, 4, false>, RT_SMEM_POST3) &&
                     set_smem(RT::k_post_rt3<4, 4, true>, RT_SMEM_POST3) && set_smem(RT::k_post_rt3<4, 4, false>, RT_SMEM_POST3);
    s->rt_ctas = 2 * s->num_sms;
    if (const char *e = getenv("EQGPU_RT_CTAS")) s->rt_ctas = std::max(1, atoi(e));
    EQ_CUDA(cudaMalloc(&s->rt_sched, sizeof(unsigned) * 4));
    EQ_CUDA(cudaMemset(s->rt_sched, 0, sizeof(unsigned) * 4));
    {
        int prio_lo = 0, prio_hi = 0;   // highest priority: the few general tiles should start first, beside the others
        cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
        EQ_CUDA(cudaStreamCreateWithPriority(&s->rt_stream, cudaStreamNonBlocking, prio_hi));
    }
    EQ_CUDA(cudaEventCreateWithFlags(&s->ev_rt_fork, cudaEventDisableTiming));
    EQ_CUDA(cudaEventCreateWithFlags(&s->ev_rt_join, cudaEventDisableTiming));
    // levels 0 .. max_level may use the register-tile kernels.  Level 1 of the 2048^2 hierarchy (1025^2: odd pitch, no TMA,
    // a narrower last cell, ~1 tile per CTA) measured no faster than the tile kernels: level 0 only unless asked
    int max_level = 0;
    if (const char *e = getenv("EQGPU_RT_LEVELS")) max_level = atoi(e) - 1;
    const bool no_tma = getenv("EQGPU_NO_TMA") != nullptr;
    for (size_t l = 0; l + 1 < s->levels.size() && (int)l <= max_level; ++l) {
        Level &lv = s->levels[l];
        const int nu = l == 0 ? s->nu : s->nuc;
        if (nu != 3 && nu != 4) continue;
        lv.rt_tma = !no_tma && make_tile_map(&lv.map_b64, lv.b, lv.dev.nx, lv.dev.ny) &&
                    make_tile_map(&lv.map_t64, lv.t, lv.dev.nx, lv.dev.ny);
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
        for (Level::RtPlan *P : {&lv.rt_pre, &lv.rt_post}) {
            cudaFree(P->d_tlist); P->d_tlist = nullptr;
            cudaFree(P->d_tiles); P->d_tiles = nullptr;
            P->on = false;
        }
    }
    cudaFree(s->rt_sched); s->rt_sched = nullptr;
    if (s->ev_rt_fork) { cudaEventDestroy(s->ev_rt_fork); s->ev_rt_fork = nullptr; }
    if (s->ev_rt_join) { cudaEventDestroy(s->ev_rt_join); s->ev_rt_join = nullptr; }
    if (s->rt_stream) { cudaStreamDestroy(s->rt_stream); s->rt_stream = nullptr; }
}

static RT::Geom geom_of(const Level::RtPlan &P)
{
    RT::Geom G;
    G.gx = P.gx; G.gy = P.gy; G.n = P.n; G.tiles = P.d_tiles;
    return G;
}

#define RT_LAUNCH(PDL_OK, KERN, SM, ST, ...)                                                              \
    do {                                                                                                  \
        cudaLaunchConfig_t cfg_{};                                                                        \
        cfg_.gridDim = dim3(std::min(ctas, G.n)); cfg_.blockDim = dim3(RT::NT);                           \
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
    const double *b = lv.b;
    const bool lean = s->rt_lean && lv.rt_tma;
    const int ctas = lean ? (s->rt_ctas / 2) * 3 : s->rt_ctas;
    if (lean) {
        if (nu == 3) RT_LAUNCH(pdl_ok, (RT::k_pre_rt3<3, 4>), RT_SMEM_PRE3, st, lv.dev, cv.dev, lv.map_b64, lv.t, cv.b, sw, G, s->rt_sched, scc);
        else RT_LAUNCH(pdl_ok, (RT::k_pre_rt3<4, 6>), RT_SMEM_PRE3, st, lv.dev, cv.dev, lv.map_b64, lv.t, cv.b, sw, G, s->rt_sched, scc);
        return;
    }
#define RT_PRE(NU, H, TMA) \
    RT_LAUNCH(pdl_ok, (RT::k_pre_rt<NU, H, TMA>), RT_SMEM_PRE, st, lv.dev, cv.dev, lv.map_b64, b, lv.t, cv.b, sw, G, s->rt_sched, scc)
    if (nu == 3) { if (lv.rt_tma) RT_PRE(3, 4, true); else RT_PRE(3, 4, false); }
    else { if (lv.rt_tma) RT_PRE(4, 6, true); else RT_PRE(4, 6, false); }
#undef RT_PRE
}

void rt_launch_post(eqgpu_solver *s, cudaStream_t st, int l, int nu, const SmoothW &sw, bool dot, double *out_dot)
{
    Level &lv = s->levels[l], &cv = s->levels[l + 1];
    const RT::Geom G = geom_of(lv.rt_post);
    const double *xc = cv.x, *b = lv.b, *xin = lv.t;
    const bool lean = s->rt_lean && lv.rt_tma;
    const int ctas = lean ? (s->rt_ctas / 2) * 3 : s->rt_ctas;
    if (lean) {
#define RT_POST3(NU, DOT)                                                                                                  \
    RT_LAUNCH(true, (RT::k_post_rt3<NU, 4, DOT>), RT_SMEM_POST3, st, lv.dev, cv.dev, lv.map_b64, xin, lv.x, xc, sw, G,       \
              s->rt_sched + 2, s->sc, s->partials, s->counters + 1, out_dot)
        if (nu == 3) { if (dot) RT_POST3(3, true); else RT_POST3(3, false); }
        else { if (dot) RT_POST3(4, true); else RT_POST3(4, false); }
#undef RT_POST3
        return;
    }
#define RT_POST(NU, DOT, TMA)                                                                                              \
    RT_LAUNCH(true, (RT::k_post_rt<NU, 4, DOT, TMA>), RT_SMEM_POST, st, lv.dev, cv.dev, lv.map_b64, lv.map_t64, b, xin, lv.x, \
              xc, sw, G, s->rt_sched + 2, s->sc, s->partials, s->counters + 1, out_dot)
#define RT_POST2(NU, DOT) do { if (lv.rt_tma) RT_POST(NU, DOT, true); else RT_POST(NU, DOT, false); } while (0)
    if (nu == 3) { if (dot) RT_POST2(3, true); else RT_POST2(3, false); }
    else { if (dot) RT_POST2(4, true); else RT_POST2(4, false); }
#undef RT_POST2
#undef RT_POST
}
