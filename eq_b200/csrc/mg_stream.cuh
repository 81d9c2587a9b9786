// Streaming multigrid smoothers (isotropic operator, one GPU): the round-2 replacement of the 64x64 shared-memory
// tile kernels of mg_tile.inc on the levels that carry the bytes.
//
// Why: ncu on the tile kernels (profiles/r01_ncu_iteration_final.md) shows them bound by shared-memory wavefronts --
// every node update of every sweep costs 2 LDS.64 + 1 STS.64 plus a __syncthreads per sweep -- at 0.31 of HBM peak.
// Here a node update touches NO shared memory:
//   * one WARP owns a strip of 64 columns (two adjacent columns per lane) and walks it upwards row by row;
//   * all NU sweeps (+ residual + restriction, or prolongation + sweeps) run as a software pipeline skewed by one row
//     per sweep: at walk step i sweep k works on row i-k, so the whole temporally blocked smoother is ONE pass over
//     the rows, with a 3-row sliding window per sweep kept in REGISTERS (sweep k lags two rows behind sweep k-1, so
//     that the update chains of one walk step are independent of each other);
//   * west/east neighbours come from the neighbouring lanes by warp shuffles, north/south from the window;
//   * the inputs (right-hand side b, incoming iterate) are staged ahead of the walk into a small per-warp ring of
//     shared-memory row blocks by TMA (cp.async.bulk.tensor.2d + mbarrier, out-of-grid rows/columns zero-filled by the
//     hardware) where the row pitch is 16-byte aligned, else by 8-byte cp.async; results leave through 128-bit stores;
//   * no __syncthreads anywhere in the walk: warps are independent tasks (strip x row chunk).
// Redundancy: the halo is 2 x NU' columns of 64 per strip and NU' + NU + 1 rows per chunk of HS rows.
//
// Boundaries cost nothing on the fast path: every lane carries the seven (negated) stencil coefficients of ITS two
// columns for a regular row -- wall columns, Robin columns and the narrower last cell of a coarse grid are just
// different numbers in the same registers, Dirichlet and out-of-grid columns are all-zero coefficient sets -- and rows
// that are irregular (row 0, the rows above ireg_hi) take a slow path that evaluates the general row (stencil_iso).
#pragma once
#include <cuda.h>
#include <cuda/std/type_traits>
#include "eqgpu_internal.cuh"
#include "mg_fused.cuh"

namespace STRM {

constexpr int SWID = 64;                 // strip width in nodes: two per lane
constexpr int RB = 6;                    // rows per staged block == walk steps per unrolled loop body
constexpr int NSLOT = 3;                 // ring slots per input stream per warp
constexpr int BLK_BYTES = RB * SWID * 8; // 3072: one staged block of one stream
constexpr int CROWS = RB / 2 + 1;        // coarse rows under one block (post-smoothing)
constexpr int CCOLS = SWID / 2 + 2;      // coarse columns under one strip (+1, padded to even)
constexpr int CBLK_BYTES = CROWS * CCOLS * 8;   // 1088
constexpr int WPC = 2;                   // warps (independent tasks) per CTA
#ifndef EQ_UNI_CTAS
#define EQ_UNI_CTAS 6
#endif
constexpr int UNI_CTAS = EQ_UNI_CTAS;    // resident CTAs per SM the UNI smoother instances are compiled for (register cap)

// negated row of A and the weighted inverse diagonals of the sweeps, for one node column
template <int NU>
struct Coef {
    double nC, nE, nW, nN, nS, nNE, nSW;
    double wic[NU];
};

// Uniform (regular-node) coefficient set, negated and weighted on the host: a kernel parameter, i.e. constant-bank
// operands of the FMAs instead of 2 x (7 + NU) doubles of registers per lane.  Used by the UNI instances, whose strips
// contain only regular and Dirichlet / out-of-grid columns (the latter carry a zero update weight and a zero mask).
template <int NU>
struct UniCoef {
    double nC, nE, nW, nN, nS, nNE, nSW;
    double wic[NU];
};

struct StreamGeom {
    int nstrips, nchunks, hs;   // strips across, chunks up, owned rows per chunk
    int halo;                   // even halo (columns per side, rows below)
    int nblk;                   // staged blocks (of RB rows) per task
};

// ---- mbarrier / TMA / cp.async primitives ------------------------------------------------------------------------------
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

// One input stream of a warp: NSLOT blocks of RB rows x 64 columns in shared memory.
struct Ring {
    double *base;               // NSLOT * RB * 64 doubles
    unsigned long long *bar;    // NSLOT mbarriers (TMA path)
};

// Stage block `blk` (global rows y .. y+RB-1, columns ox .. ox+63) of vector g into its ring slot.  TMA: lane 0
// arms the slot's barrier and issues one tensor copy (rows/columns outside the tensor arrive as zeros).  Otherwise
// every lane issues 8-byte cp.async copies and stores zeros where the source is outside [0,nx) x [0,ny).
template <bool TMA>
__device__ __forceinline__ void stage_block(const Ring &r, int blk, const CUtensorMap *map, const double *__restrict__ g,
                                            int nx, int ny, int ox, int y, bool also_second, const Ring &r2,
                                            const CUtensorMap *map2, const double *__restrict__ g2)
{
    const int slot = blk % NSLOT, lane = threadIdx.x & 31;
    if (TMA) {
        if (lane == 0) {
            fence_proxy_async();   // the slot's previous contents were read through the generic proxy
            mbar_expect_tx(r.bar + slot, also_second ? 2 * BLK_BYTES : BLK_BYTES);
            tma_load_2d(r.base + slot * (RB * SWID), map, r.bar + slot, ox, y);
            if (also_second) tma_load_2d(r2.base + slot * (RB * SWID), map2, r.bar + slot, ox, y);
        }
    } else {
        double *dst = r.base + slot * (RB * SWID), *dst2 = also_second ? r2.base + slot * (RB * SWID) : nullptr;
#pragma unroll
        for (int q = 0; q < RB * SWID / 32; ++q) {
            const int e = q * 32 + lane, row = e >> 6, col = e & 63;
            const int gi = y + row, gj = ox + col;
            if (gi >= 0 && gi < ny && gj >= 0 && gj < nx) {
                cp_async8(dst + e, g + (size_t)gi * nx + gj);
                if (also_second) cp_async8(dst2 + e, g2 + (size_t)gi * nx + gj);
            } else {
                dst[e] = 0.0;
                if (also_second) dst2[e] = 0.0;
            }
        }
    }
}

// Coarse patch under block `blk` of a post-smoothing strip: coarse rows I0 .. I0+CROWS-1, columns J0 .. J0+CCOLS-1
// (clamped loads; values of coarse nodes outside the grid are never used with a non-zero weight, zeros keep them finite)
__device__ __forceinline__ void stage_coarse(double *cring, int blk, const double *__restrict__ xc, int cnx, int cny,
                                             int J0, int I0)
{
    const int slot = blk % NSLOT, lane = threadIdx.x & 31;
    double *dst = cring + slot * (CROWS * CCOLS);
    for (int e = lane; e < CROWS * CCOLS; e += 32) {
        const int row = e / CCOLS, col = e - row * CCOLS;
        const int I = I0 + row, J = J0 + col;
        if (I >= 0 && I < cny && J >= 0 && J < cnx) cp_async8(dst + e, xc + (size_t)I * cnx + J);
        else dst[e] = 0.0;
    }
}

// ---- coefficient sets ----------------------------------------------------------------------------------------------------
template <int NU>
__device__ __forceinline__ void zero_coef(Coef<NU> &c)
{
    c.nC = c.nE = c.nW = c.nN = c.nS = c.nNE = c.nSW = 0.0;
#pragma unroll
    for (int k = 0; k < NU; ++k) c.wic[k] = 0.0;
}

// general row (i, j): zero for nodes outside the grid and Dirichlet nodes (their value stays 0 in every sweep and
// their residual is their right-hand side, which is 0)
template <int NU>
__device__ __noinline__ void node_coef(const LevelDev &L, int i, int j, const SmoothW &sw, Coef<NU> &c)
{
    zero_coef<NU>(c);
    if (i < 0 || i >= L.ny || j < 0 || j >= L.nx || is_dirichlet(L, i, j)) return;
    double cf[NBAND];
    // padded cell sizes around the node in closed form (west/east of j, south/north of i): no global loads
    double hx[2], ihx[2], hy[2], ihy[2];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const int kx = j + q, ky = i + q;   // size of the cell WEST of node kx / SOUTH of node ky
        hx[q] = (kx == 0 || kx == L.nx) ? 0.0 : (kx == L.nx - 1 ? L.hxl : L.hxr);
        hy[q] = (ky == 0 || ky == L.ny) ? 0.0 : (ky == L.ny - 1 ? L.hyl : L.hyr);
        ihx[q] = hx[q] > 0.0 ? 1.0 / hx[q] : 0.0;
        ihy[q] = hy[q] > 0.0 ? 1.0 / hy[q] : 0.0;
    }
    Spacing S;
    S.hx = hx; S.ihx = ihx; S.hy = hy; S.ihy = ihy; S.jo = j; S.io = i;
    stencil_iso(L, S, i, j, cf);
    c.nC = -cf[B_C]; c.nE = -cf[B_E]; c.nW = -cf[B_W]; c.nN = -cf[B_N]; c.nS = -cf[B_S]; c.nNE = -cf[B_NE]; c.nSW = -cf[B_SW];
    const double id = 1.0 / cf[B_C];
#pragma unroll
    for (int k = 0; k < NU; ++k) c.wic[k] = sw.w[k] * id;
}

// The rows of a level that are not regular are row 0 and the rows above ireg_hi, i.e. (uniform mesh, coarse grids with
// a narrower last cell) at most ny-2 and ny-1.  Their coefficient sets are evaluated ONCE per task, before the walk,
// into a small per-lane table (local memory): a call inside the walk would spill the ~150 live registers of the sweep
// windows around it, which made the boundary tasks 3x slower than the interior ones in the first version of this file.
template <int NU>
struct IrrTab { Coef<NU> c[3][2]; };   // rows 0, ny-2, ny-1  x  the lane's two columns

template <int NU>
__device__ __noinline__ void fill_irr(const LevelDev &L, int j0, const SmoothW &sw, int row_lo, int row_hi, IrrTab<NU> &t)
{
    const int rows[3] = {0, L.ny - 2, L.ny - 1};
#pragma unroll 1
    for (int q = 0; q < 3; ++q) {
        const int r = rows[q];
        if (r >= row_lo && r <= row_hi && !(r >= 1 && r <= L.ireg_hi)) {
            node_coef<NU>(L, r, j0, sw, t.c[q][0]);
            node_coef<NU>(L, r, j0 + 1, sw, t.c[q][1]);
        } else {
            zero_coef<NU>(t.c[q][0]);
            zero_coef<NU>(t.c[q][1]);
        }
    }
}

// coefficient sets of row r for the lane's two columns (walk-time, no calls)
template <int NU>
__device__ __forceinline__ void row_coef(const LevelDev &L, int r, const Coef<NU> &ca, const Coef<NU> &cb,
                                         const IrrTab<NU> &t, Coef<NU> &la, Coef<NU> &lb)
{
    if (r >= 1 && r <= L.ireg_hi) { la = ca; lb = cb; }
    else if (r < 0 || r >= L.ny) { zero_coef<NU>(la); zero_coef<NU>(lb); }
    else {
        const int q = r == 0 ? 0 : (r == L.ny - 1 ? 2 : 1);
        la = t.c[q][0];
        lb = t.c[q][1];
    }
}

// the lane's regular-row sets, by value (their address must not escape: they live in registers for the whole walk)
template <int NU>
__device__ __forceinline__ void regular_coefs(const LevelDev &L, int j0, const SmoothW &sw, Coef<NU> &ca, Coef<NU> &cb)
{
    Coef<NU> ta, tb;
    const int ireg = L.ireg_hi >= 1 ? 1 : -1;
    node_coef<NU>(L, ireg, j0, sw, ta);
    node_coef<NU>(L, ireg, j0 + 1, sw, tb);
    ca = ta;
    cb = tb;
}

// UNI instances: the uniform set with the lane's column masks folded into the update weights
template <int NU>
__device__ __forceinline__ void uniform_coefs(const LevelDev &L, int j0, const UniCoef<NU> &U, Coef<NU> &ca, Coef<NU> &cb,
                                              double &ma, double &mb)
{
    auto free_col = [&](int j) {
        return j >= 0 && j < L.nx && !(((L.dirmask & 1u) && j == 0) || ((L.dirmask & 2u) && j == L.nx - 1));
    };
    ma = free_col(j0) ? 1.0 : 0.0;
    mb = free_col(j0 + 1) ? 1.0 : 0.0;
    ca.nC = cb.nC = U.nC; ca.nE = cb.nE = U.nE; ca.nW = cb.nW = U.nW; ca.nN = cb.nN = U.nN; ca.nS = cb.nS = U.nS;
    ca.nNE = cb.nNE = U.nNE; ca.nSW = cb.nSW = U.nSW;
#pragma unroll
    for (int k = 0; k < NU; ++k) { ca.wic[k] = U.wic[k] * ma; cb.wic[k] = U.wic[k] * mb; }
}

// one weighted-Jacobi update of the lane's two nodes of the window's middle row.
// win[r][0..3] = W, a, b, E of rows (oldest, middle, newest); returns the residuals, writes the new values.
template <int NU>
__device__ __forceinline__ void node_pair(const double (&win)[3][4], const Coef<NU> &ca, const Coef<NU> &cb, double ba,
                                          double bb, double &ra, double &rb)
{
    // oldest row first, the newest row (the one the previous walk step produced) last: the head of this chain does not
    // wait for the tail of the previous step's
    ra = fma(ca.nS, win[0][1], ba);
    rb = fma(cb.nS, win[0][2], bb);
    ra = fma(ca.nSW, win[0][0], ra);
    rb = fma(cb.nSW, win[0][1], rb);
    ra = fma(ca.nC, win[1][1], ra);
    rb = fma(cb.nC, win[1][2], rb);
    ra = fma(ca.nE, win[1][2], ra);
    rb = fma(cb.nE, win[1][3], rb);
    ra = fma(ca.nW, win[1][0], ra);
    rb = fma(cb.nW, win[1][1], rb);
    ra = fma(ca.nN, win[2][1], ra);
    rb = fma(cb.nN, win[2][2], rb);
    ra = fma(ca.nNE, win[2][2], ra);
    rb = fma(cb.nNE, win[2][3], rb);
}

// push a new row (a, b) into a window: the west / east values come from the neighbouring lanes
__device__ __forceinline__ void push_row(double (&win)[3][4], double a, double b)
{
#pragma unroll
    for (int q = 0; q < 4; ++q) { win[0][q] = win[1][q]; win[1][q] = win[2][q]; }
    win[2][1] = a;
    win[2][2] = b;
    win[2][0] = __shfl_up_sync(0xffffffffu, b, 1);
    win[2][3] = __shfl_down_sync(0xffffffffu, a, 1);
}

// ---- shared set-up of a task ---------------------------------------------------------------------------------------------
struct Task {
    int ox, y0;          // first column of the strip (even, may be negative), first walked row (even, may be negative)
    int lo_x, hi_x;      // owned columns [lo_x, hi_x)
    int lo_y, hi_y;      // owned rows    [lo_y, hi_y)
    int j0;              // the lane's first column
    bool own_cols;       // both of the lane's columns are owned and inside the grid (or the first is and the second is
                         // past the last column: handled at the store)
};

__device__ __forceinline__ Task make_task(const LevelDev &L, const StreamGeom &G, int strip, int chunk)
{
    Task t;
    const int so = SWID - 2 * G.halo, lane = threadIdx.x & 31;
    t.lo_x = strip * so;
    t.hi_x = min(t.lo_x + so, L.nx);
    t.ox = t.lo_x - G.halo;
    t.lo_y = chunk * G.hs;
    t.hi_y = min(t.lo_y + G.hs, L.ny);
    t.y0 = t.lo_y - G.halo;
    t.j0 = t.ox + 2 * lane;
    t.own_cols = t.j0 >= t.lo_x && t.j0 < t.hi_x;
    return t;
}

// store the lane's two values of row gi (both columns owned by construction; the second may be past the grid)
__device__ __forceinline__ void store_pair(double *__restrict__ x, const LevelDev &L, int gi, int j0, double a, double b)
{
    double *q = x + (size_t)gi * L.nx + j0;
    if (j0 + 1 < L.nx) {
        if ((((size_t)q) & 15) == 0) *reinterpret_cast<double2 *>(q) = make_double2(a, b);
        else { q[0] = a; q[1] = b; }
    } else q[0] = a;
}

// ==========================================================================================================================
// pre-smoothing: NU sweeps from a zero guess, residual, restriction.        reads b        writes x, b_coarse
// ==========================================================================================================================
template <int NU, bool TMA, bool UNI>
__global__ void __launch_bounds__(32 * WPC, UNI ? UNI_CTAS : 4)
ks_presmooth(LevelDev F, LevelDev Cc, const __grid_constant__ CUtensorMap map_b, const double *__restrict__ b,
             double *__restrict__ x, double *__restrict__ bc, SmoothW sw, UniCoef<NU> U, StreamGeom G, const CGScalars *sc)
{
    pdl_trigger();
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    Ring rb;
    rb.base = reinterpret_cast<double *>(smem_raw) + warp * (NSLOT * RB * SWID);
    rb.bar = reinterpret_cast<unsigned long long *>(smem_raw + WPC * NSLOT * BLK_BYTES) + warp * NSLOT;
    const int strip = blockIdx.x * WPC + warp, chunk = blockIdx.y;
    const bool active = strip < G.nstrips;
    const Task T = make_task(F, G, strip, chunk);
    if (TMA && lane == 0) {
#pragma unroll
        for (int q = 0; q < NSLOT; ++q) mbar_init(rb.bar + q, 1);
        fence_barrier_init();
    }
    __syncwarp();
    // the lane's coefficient sets for a regular row (set-up constants only: safe before the dependency wait)
    const bool has_reg = F.ireg_hi >= 1;
    Coef<NU> ca, cb;
    double ma = 1.0, mb = 1.0;
    if (UNI) uniform_coefs<NU>(F, T.j0, U, ca, cb, ma, mb);
    else regular_coefs<NU>(F, T.j0, sw, ca, cb);
    IrrTab<NU> irr;
    fill_irr<NU>(F, T.j0, sw, T.y0 - 2 * NU - 2, T.y0 + G.nblk * RB, irr);
    // restriction weights of the lane's columns: the west / east neighbours of its coarse column j0 are midpoints?
    const bool cj0 = T.j0 >= 0 && T.j0 < F.nx;                               // j0 is even: a coarse column
    const double wE = (T.j0 + 1 < F.nx && is_mid(T.j0 + 1, F.nx)) ? 0.5 : 0.0;
    const double wW = (T.j0 - 1 >= 0 && is_mid(T.j0 - 1, F.nx)) ? 0.5 : 0.0;
    const bool cj1 = (T.j0 + 1 == F.nx - 1) && !is_mid(T.j0 + 1, F.nx) && T.j0 + 1 >= T.lo_x && T.j0 + 1 < T.hi_x;   // odd last column
    const int J0 = T.j0 >> 1;
    pdl_wait();
    if (sc->done || !active) return;
    // one cp.async group per block, committed whether or not it holds copies, so that "all but the NSLOT-1 newest
    // groups have landed" always means "this block has landed"
    for (int q = 0; q < NSLOT; ++q) {
        if (q < G.nblk) stage_block<TMA>(rb, q, &map_b, b, F.nx, F.ny, T.ox, T.y0 + q * RB, false, rb, nullptr, nullptr);
        if (!TMA) cp_async_commit();
    }

    double xw[NU][3][4];        // xw[k]: window of x_{k+1}
    double rw[3][3];            // residual rows (W, a, b)
    double bw[2 * NU + 1][2];   // b rows i, i-1, .., i-2NU
#pragma unroll
    for (int k = 0; k < NU; ++k)
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int q = 0; q < 4; ++q) xw[k][r][q] = 0.0;
#pragma unroll
    for (int r = 0; r < 3; ++r) { rw[r][0] = rw[r][1] = rw[r][2] = 0.0; }
#pragma unroll
    for (int k = 0; k <= 2 * NU; ++k) { bw[k][0] = bw[k][1] = 0.0; }

    // One walk step: row i enters.  Every sweep reads its predecessor's window AS THE PREVIOUS STEP LEFT IT and only then
    // do the windows take their new rows, so the NU + 1 update chains of a step (and the shuffles that publish the new
    // rows) are independent of each other: sweep k works on row i - 2k.  (The first version ran sweep k on row i - k,
    // each sweep waiting for its predecessor's shuffle within the step: ncu showed one instruction issued every 5.8
    // cycles per warp, the fixed-latency `wait` stall on top.)  PAR = parity of i where the compiler knows it, else -1.
    auto step = [&](int i, const double2 bv, const bool fast, const int PAR) {
#pragma unroll
        for (int k = 2 * NU; k > 0; --k) { bw[k][0] = bw[k - 1][0]; bw[k][1] = bw[k - 1][1]; }
        bw[0][0] = bv.x; bw[0][1] = bv.y;
        Coef<NU> la, lb;
        double na[NU], nb[NU];
        // sweep 1: x1 = w0 D^-1 b on row i
        if (fast) { la = ca; lb = cb; } else row_coef<NU>(F, i, ca, cb, irr, la, lb);
        na[0] = la.wic[0] * bv.x;
        nb[0] = lb.wic[0] * bv.y;
        // sweeps 2..NU: x_{k+1}(i-2k) from the window of x_k, which holds rows i-2k-1 .. i-2k+1
#pragma unroll
        for (int k = 1; k < NU; ++k) {
            const int r = i - 2 * k;
            if (fast) { la = ca; lb = cb; } else row_coef<NU>(F, r, ca, cb, irr, la, lb);
            double ra, rbv;
            node_pair<NU>(xw[k - 1], la, lb, bw[2 * k][0], bw[2 * k][1], ra, rbv);
            na[k] = fma(la.wic[k], ra, xw[k - 1][1][1]);
            nb[k] = fma(lb.wic[k], rbv, xw[k - 1][1][2]);
        }
        // residual of x_NU on row i-2NU; the iterate itself is the middle row of its window
        double ra, rbv;
        {
            const int r = i - 2 * NU;
            if (fast) { la = ca; lb = cb; } else row_coef<NU>(F, r, ca, cb, irr, la, lb);
            node_pair<NU>(xw[NU - 1], la, lb, bw[2 * NU][0], bw[2 * NU][1], ra, rbv);
            if (UNI) { ra *= ma; rbv *= mb; }   // uniform rows do not vanish on Dirichlet columns: their residual is 0
            if (T.own_cols && r >= T.lo_y && r < T.hi_y) store_pair(x, F, r, T.j0, xw[NU - 1][1][1], xw[NU - 1][1][2]);
        }
        // now the windows move on
#pragma unroll
        for (int k = 0; k < NU; ++k) push_row(xw[k], na[k], nb[k]);
#pragma unroll
        for (int q = 0; q < 3; ++q) { rw[0][q] = rw[1][q]; rw[1][q] = rw[2][q]; }
        rw[2][1] = ra; rw[2][2] = rbv;
        rw[2][0] = __shfl_up_sync(0xffffffffu, rbv, 1);
        // restriction at fine row rc = i-2NU-1 (middle residual row) when it is a coarse row; it reads this step's
        // residual row only through the lane's own two values
        {
            const int rc = i - 2 * NU - 1;
            const bool even = PAR >= 0 ? ((PAR & 1) == 1) : ((rc & 1) == 0);
            if ((even || (!fast && rc == F.ny - 1)) && rc >= T.lo_y && rc < T.hi_y) {
                const int I = coarse_lo(rc, F.ny, Cc.ny);
                // fast blocks lie strictly inside the regular rows: both row neighbours are midpoints
                const double wN = fast ? 0.5 : ((rc + 1 < F.ny && is_mid(rc + 1, F.ny)) ? 0.5 : 0.0);
                const double wS = fast ? 0.5 : ((rc - 1 >= 0 && is_mid(rc - 1, F.ny)) ? 0.5 : 0.0);
                if (T.own_cols && cj0) {
                    double v = rw[1][1] + wE * rw[1][2] + wW * rw[1][0] + wN * rw[2][1] + wS * rw[0][1] +
                               (2.0 * wN * wE) * rw[2][2] + (2.0 * wS * wW) * rw[0][0];
                    if (is_dirichlet(Cc, I, J0)) v = 0.0;
                    bc[(size_t)I * Cc.nx + J0] = v;
                }
                if (cj1) {   // the odd last column is a coarse column of its own: no east/west midpoints
                    double v = rw[1][2] + wN * rw[2][2] + wS * rw[0][2];
                    if (is_dirichlet(Cc, I, Cc.nx - 1)) v = 0.0;
                    bc[(size_t)I * Cc.nx + Cc.nx - 1] = v;
                }
            }
        }
    };

    for (int blk = 0; blk < G.nblk; ++blk) {
        const int slot = blk % NSLOT, i0 = T.y0 + blk * RB;
        if (TMA) mbar_wait(rb.bar + slot, (blk / NSLOT) & 1);
        else {
            cp_async_wait<NSLOT - 1>();
            __syncwarp();
        }
        const double2 *rowp = reinterpret_cast<const double2 *>(rb.base + slot * (RB * SWID)) + lane;
        // every row any sweep touches in this block is regular: rows i0-2NU-2 .. i0+RB-1
        const bool fast = has_reg && i0 - 2 * NU - 2 >= 1 && i0 + RB - 1 <= F.ireg_hi;
        if (fast) {
#pragma unroll
            for (int u = 0; u < RB; ++u) step(i0 + u, rowp[u * (SWID / 2)], true, u & 1);   // y0 even: parity of row = parity of u
        } else {
#pragma unroll 1
            for (int u = 0; u < RB; ++u) step(i0 + u, rowp[u * (SWID / 2)], false, -1);
        }
        __syncwarp();
        if (blk + NSLOT < G.nblk) {
            stage_block<TMA>(rb, blk + NSLOT, &map_b, b, F.nx, F.ny, T.ox, T.y0 + (blk + NSLOT) * RB, false, rb, nullptr, nullptr);
        }
        if (!TMA) cp_async_commit();
    }
}

// ==========================================================================================================================
// post-smoothing: x = xin + P xc, NU sweeps (+ x.b).        reads b, xin, x_coarse        writes x
// ==========================================================================================================================
template <int NU, bool DOT, bool TMA, bool UNI>
__global__ void __launch_bounds__(32 * WPC, UNI ? UNI_CTAS : 4)
ks_postsmooth(LevelDev F, LevelDev Cc, const __grid_constant__ CUtensorMap map_b, const __grid_constant__ CUtensorMap map_x,
              const double *__restrict__ b, const double *__restrict__ xin, double *__restrict__ x,
              const double *__restrict__ xc, SmoothW sw, UniCoef<NU> U, StreamGeom G, CGScalars *sc, double *partials,
              unsigned *counter, double *out_dot)
{
    pdl_trigger();
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    Ring rb, rx;
    rb.base = reinterpret_cast<double *>(smem_raw) + warp * (NSLOT * RB * SWID);
    rx.base = reinterpret_cast<double *>(smem_raw + WPC * NSLOT * BLK_BYTES) + warp * (NSLOT * RB * SWID);
    double *cring = reinterpret_cast<double *>(smem_raw + 2 * WPC * NSLOT * BLK_BYTES) + warp * (NSLOT * CROWS * CCOLS);
    rb.bar = reinterpret_cast<unsigned long long *>(smem_raw + 2 * WPC * NSLOT * BLK_BYTES + WPC * NSLOT * CBLK_BYTES) +
             warp * NSLOT;
    rx.bar = rb.bar;
    const int strip = blockIdx.x * WPC + warp, chunk = blockIdx.y;
    const bool active = strip < G.nstrips;
    const Task T = make_task(F, G, strip, chunk);
    if (TMA && lane == 0) {
#pragma unroll
        for (int q = 0; q < NSLOT; ++q) mbar_init(rb.bar + q, 1);
        fence_barrier_init();
    }
    __syncwarp();
    const bool has_reg = F.ireg_hi >= 1;
    Coef<NU> ca, cb;
    double ma = 1.0, mb = 1.0;
    if (UNI) uniform_coefs<NU>(F, T.j0, U, ca, cb, ma, mb);
    else regular_coefs<NU>(F, T.j0, sw, ca, cb);
    IrrTab<NU> irr;
    fill_irr<NU>(F, T.j0, sw, T.y0 - 2 * NU - 1, T.y0 + G.nblk * RB, irr);
    // prolongation along x for the lane's columns: j0 is even (coincides with coarse column J0, or lies outside the
    // grid); j0+1 is a midpoint (half of J0 and J0+1) or the odd last column (coarse column J0+1 itself)
    const int J0 = T.j0 >> 1, CJ0 = T.ox >> 1;   // coarse column of the lane, first coarse column of the strip's patch
    const bool in0 = T.j0 >= 0 && T.j0 < F.nx, in1 = T.j0 + 1 >= 0 && T.j0 + 1 < F.nx;
    const bool mid1 = in1 && is_mid(T.j0 + 1, F.nx);
    // column masks: Dirichlet and out-of-grid columns stay 0
    const double m0 = (in0 && !(((F.dirmask & 1u) && T.j0 == 0) || ((F.dirmask & 2u) && T.j0 == F.nx - 1))) ? 1.0 : 0.0;
    const double m1 = (in1 && !(((F.dirmask & 1u) && T.j0 + 1 == 0) || ((F.dirmask & 2u) && T.j0 + 1 == F.nx - 1))) ? 1.0 : 0.0;
    pdl_wait();
    double dot = 0.0;
    if (!sc->done && active) {
        for (int q = 0; q < NSLOT; ++q) {   // one cp.async group per block (see ks_presmooth)
            if (q < G.nblk) {
                stage_block<TMA>(rb, q, &map_b, b, F.nx, F.ny, T.ox, T.y0 + q * RB, true, rx, &map_x, xin);
                stage_coarse(cring, q, xc, Cc.nx, Cc.ny, CJ0, (T.y0 + q * RB) >> 1);
            }
            cp_async_commit();
        }
        double xw[NU][3][4];        // xw[0]: prolongated iterate, xw[k]: after k sweeps (the last sweep is stored, not kept)
        double bw[2 * NU + 1][2];   // b rows i .. i-2NU
#pragma unroll
        for (int k = 0; k < NU; ++k)
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int q = 0; q < 4; ++q) xw[k][r][q] = 0.0;
#pragma unroll
        for (int k = 0; k <= 2 * NU; ++k) { bw[k][0] = bw[k][1] = 0.0; }
        // one walk step (see ks_presmooth): sweep k works on row i - 2k from the windows as the previous step left them
        auto step = [&](int i, const double2 bv, const double2 xv, const double *crow, const bool fast, const int PAR) {
#pragma unroll
            for (int k = 2 * NU; k > 0; --k) { bw[k][0] = bw[k - 1][0]; bw[k][1] = bw[k - 1][1]; }
            bw[0][0] = bv.x; bw[0][1] = bv.y;
            double na[NU], nb[NU];
            // prolongation on row i: crow points at the coarse row under it (row (i - i0)/2 of the block's patch), the
            // lane's coarse column first; a midpoint row also takes the coarse row above
            {
                const bool rowin = fast || (i >= 0 && i < F.ny);
                const bool odd = PAR >= 0 ? (PAR & 1) : (i & 1);
                const bool mi = odd && (fast || i != F.ny - 1);    // is_mid(i, ny)
                // coarse row of a non-midpoint row: i/2, or (odd last row) (i+1)/2 = the row above in the patch
                const double *c0 = crow + ((odd && !mi) ? CCOLS : 0);
                double pa, pb;
                if (!mi) {
                    pa = c0[0];
                    pb = mid1 ? 0.5 * (c0[0] + c0[1]) : c0[1];
                } else {
                    pa = 0.5 * (c0[0] + c0[CCOLS]);
                    pb = mid1 ? 0.5 * (c0[0] + c0[CCOLS + 1]) : 0.5 * (c0[1] + c0[CCOLS + 1]);
                }
                double rm = rowin ? 1.0 : 0.0;
                if (!fast && rowin && (((F.dirmask & 8u) && i == 0) || ((F.dirmask & 4u) && i == F.ny - 1))) rm = 0.0;
                na[0] = rm * m0 * (xv.x + pa);
                nb[0] = rm * m1 * (xv.y + pb);
            }
            Coef<NU> la, lb;
#pragma unroll
            for (int k = 1; k <= NU; ++k) {
                const int r = i - 2 * k;
                if (fast) { la = ca; lb = cb; } else row_coef<NU>(F, r, ca, cb, irr, la, lb);
                double ra, rbv;
                node_pair<NU>(xw[k - 1], la, lb, bw[2 * k][0], bw[2 * k][1], ra, rbv);
                const double va = fma(la.wic[k - 1], ra, xw[k - 1][1][1]), vb = fma(lb.wic[k - 1], rbv, xw[k - 1][1][2]);
                if (k < NU) { na[k] = va; nb[k] = vb; }
                else if (T.own_cols && r >= T.lo_y && r < T.hi_y) {
                    store_pair(x, F, r, T.j0, va, vb);
                    if (DOT) dot += va * bw[2 * k][0] + (T.j0 + 1 < F.nx ? vb * bw[2 * k][1] : 0.0);
                }
            }
#pragma unroll
            for (int k = 0; k < NU; ++k) push_row(xw[k], na[k], nb[k]);
        };
        for (int blk = 0; blk < G.nblk; ++blk) {
            const int slot = blk % NSLOT, i0 = T.y0 + blk * RB;
            cp_async_wait<NSLOT - 1>();
            if (TMA) mbar_wait(rb.bar + slot, (blk / NSLOT) & 1);
            __syncwarp();
            const double2 *rowb = reinterpret_cast<const double2 *>(rb.base + slot * (RB * SWID)) + lane;
            const double2 *rowx = reinterpret_cast<const double2 *>(rx.base + slot * (RB * SWID)) + lane;
            const double *cp = cring + slot * (CROWS * CCOLS) + (J0 - CJ0);
            const bool fast = has_reg && i0 - 2 * NU - 1 >= 1 && i0 + RB - 1 <= F.ireg_hi;
            if (fast) {
#pragma unroll
                for (int u = 0; u < RB; ++u)
                    step(i0 + u, rowb[u * (SWID / 2)], rowx[u * (SWID / 2)], cp + (u >> 1) * CCOLS, true, u & 1);
            } else {
#pragma unroll 1
                for (int u = 0; u < RB; ++u)
                    step(i0 + u, rowb[u * (SWID / 2)], rowx[u * (SWID / 2)], cp + (u >> 1) * CCOLS, false, -1);
            }
            __syncwarp();
            if (blk + NSLOT < G.nblk) {
                stage_block<TMA>(rb, blk + NSLOT, &map_b, b, F.nx, F.ny, T.ox, T.y0 + (blk + NSLOT) * RB, true, rx, &map_x, xin);
                stage_coarse(cring, blk + NSLOT, xc, Cc.nx, Cc.ny, CJ0, (T.y0 + (blk + NSLOT) * RB) >> 1);
            }
            cp_async_commit();
        }
    }
    if (DOT) {
        double v[1] = {dot}, tot[1];
        if (grid_reduce<1>(v, partials, counter, tot) && !sc->done) *out_dot = tot[0];
    }
}

// ==========================================================================================================================
// p' = z + beta p ; Ap = A p' ; p'.Ap        reads z, p        writes p', Ap        (one layer: halo 2 columns, 2+1 rows)
// ==========================================================================================================================
template <bool TMA, bool UNI>
__global__ void __launch_bounds__(32 * WPC)
ks_apply_p(LevelDev F, const __grid_constant__ CUtensorMap map_z, const __grid_constant__ CUtensorMap map_p,
           const double *__restrict__ z, const double *__restrict__ pin, double *__restrict__ p, double *__restrict__ Ap,
           UniCoef<1> U, StreamGeom G, CGScalars *sc, double *partials, unsigned *counter, double *out_pAp)
{
    pdl_trigger();
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    Ring rz, rp;
    rz.base = reinterpret_cast<double *>(smem_raw) + warp * (NSLOT * RB * SWID);
    rp.base = reinterpret_cast<double *>(smem_raw + WPC * NSLOT * BLK_BYTES) + warp * (NSLOT * RB * SWID);
    rz.bar = reinterpret_cast<unsigned long long *>(smem_raw + 2 * WPC * NSLOT * BLK_BYTES) + warp * NSLOT;
    rp.bar = rz.bar;
    const int strip = blockIdx.x * WPC + warp, chunk = blockIdx.y;
    const bool active = strip < G.nstrips;
    const Task T = make_task(F, G, strip, chunk);
    if (TMA && lane == 0) {
#pragma unroll
        for (int q = 0; q < NSLOT; ++q) mbar_init(rz.bar + q, 1);
        fence_barrier_init();
    }
    __syncwarp();
    const bool has_reg = F.ireg_hi >= 1;
    SmoothW sw1;
    sw1.w[0] = sw1.w[1] = sw1.w[2] = sw1.w[3] = 0.0;
    Coef<1> ca, cb;
    double ma = 1.0, mb = 1.0;
    if (UNI) uniform_coefs<1>(F, T.j0, U, ca, cb, ma, mb);
    else regular_coefs<1>(F, T.j0, sw1, ca, cb);
    IrrTab<1> irr;
    fill_irr<1>(F, T.j0, sw1, T.y0 - 2, T.y0 + G.nblk * RB, irr);
    pdl_wait();
    // a deferred x update (k_update_x) has completed before this kernel starts: clear its pending mark
    if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) sc->x_applied = sc->x_stamp;
    double dot = 0.0;
    if (!sc->done && active) {
        const double beta = sc->iters == 0 ? 0.0 : sc->rz_new / sc->rz_old;
        for (int q = 0; q < NSLOT; ++q) {
            if (q < G.nblk) stage_block<TMA>(rz, q, &map_z, z, F.nx, F.ny, T.ox, T.y0 + q * RB, true, rp, &map_p, pin);
            if (!TMA) cp_async_commit();
        }
        double pw[3][4];
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int q = 0; q < 4; ++q) pw[r][q] = 0.0;
        auto step = [&](int i, const double2 zv, const double2 pv, const bool fast) {
            // A p' on the middle row of the window as the previous step left it (row i-2), then row i enters
            const int r = i - 2;
            Coef<1> la, lb;
            if (fast) { la = ca; lb = cb; } else row_coef<1>(F, r, ca, cb, irr, la, lb);
            double ra, rbv;
            node_pair<1>(pw, la, lb, 0.0, 0.0, ra, rbv);   // = -(A p')
            if (UNI) { ra *= ma; rbv *= mb; }
            if (T.own_cols && r >= T.lo_y && r < T.hi_y) {
                const double pa = pw[1][1], pb = pw[1][2];
                store_pair(p, F, r, T.j0, pa, pb);
                store_pair(Ap, F, r, T.j0, -ra, -rbv);
                dot -= pa * ra + (T.j0 + 1 < F.nx ? pb * rbv : 0.0);
            }
            push_row(pw, fma(beta, pv.x, zv.x), fma(beta, pv.y, zv.y));
        };
        for (int blk = 0; blk < G.nblk; ++blk) {
            const int slot = blk % NSLOT, i0 = T.y0 + blk * RB;
            if (TMA) mbar_wait(rz.bar + slot, (blk / NSLOT) & 1);
            else {
                cp_async_wait<NSLOT - 1>();
                __syncwarp();
            }
            const double2 *rowz = reinterpret_cast<const double2 *>(rz.base + slot * (RB * SWID)) + lane;
            const double2 *rowp = reinterpret_cast<const double2 *>(rp.base + slot * (RB * SWID)) + lane;
            const bool fast = has_reg && i0 - 2 >= 1 && i0 + RB - 1 <= F.ireg_hi;
            if (fast) {
#pragma unroll
                for (int u = 0; u < RB; ++u) step(i0 + u, rowz[u * (SWID / 2)], rowp[u * (SWID / 2)], true);
            } else {
#pragma unroll 1
                for (int u = 0; u < RB; ++u) step(i0 + u, rowz[u * (SWID / 2)], rowp[u * (SWID / 2)], false);
            }
            __syncwarp();
            if (blk + NSLOT < G.nblk)
                stage_block<TMA>(rz, blk + NSLOT, &map_z, z, F.nx, F.ny, T.ox, T.y0 + (blk + NSLOT) * RB, true, rp, &map_p, pin);
            if (!TMA) cp_async_commit();
        }
    }
    double v[1] = {dot}, tot[1];
    if (grid_reduce<1>(v, partials, counter, tot) && !sc->done) *out_pAp = tot[0];
}

}  // namespace STRM

// ============================================================================================================================
// Warp-specialised sweep pipeline (the form the streaming smoothers ship in on TMA-capable levels).
//
// Measured on the B200 (scripts/micro/dfma_bench.cu): a dependent DFMA issues after 9 cycles, one warp alone sustains
// 0.42 DFMA per cycle and four warps x four independent chains already saturate the SM's FP64 pipe (1.82 warp-DFMA per
// cycle).  The one-warp-does-all-sweeps kernels above nevertheless issue one instruction per ~5 cycles per warp: with
// every sweep's window in the same thread they need 230-255 registers, which leaves ptxas no room to interleave the
// chains and the SM room for only eight warps.  Here the sweeps of a task (one 64-column strip x one row chunk) are
// spread over the WARPS of a CTA instead:
//     warp 0            rows of b (TMA ring)           -> sweeps 1+2 -> ring H0
//     warp k            rows of x_{k+1} (ring H_{k-1}) -> sweep k+2  -> ring H_k
//     last warp         rows of x_NU                   -> residual, restriction, stores        (pre-smoothing)
// Each warp keeps ONE three-row window (12 doubles) and its own coefficient sets, takes its input rows from shared
// memory, publishes its output rows to shared memory, and has no loop-carried dependency through its arithmetic at all,
// so consecutive walk steps overlap freely.  West/east neighbours still come from warp shuffles.  Hand-off between the
// warps is by blocks of RB rows through mbarrier full/empty pairs (no __syncthreads in the walk); every warp reads the
// rows of b it needs from the shared TMA ring, which the last warp releases.
//
// Row bookkeeping: element s of warp k's INPUT sequence is fine row y0 + s - 2k; at step s the warp emits the update of
// the middle row of its window BEFORE the entering row is pushed, i.e. row y0 + s - 2(k+1), whose right-hand side is
// element s - LG of the b sequence, LG = 2(k+1).  Block m of any sequence is elements 6m .. 6m+5.
// ============================================================================================================================
namespace PIPE {
using namespace STRM;

constexpr int NHS = 2;          // hand-off ring slots (blocks of RB rows) between two consecutive sweeps
constexpr int PFB = 2;          // blocks of b (and xin) staged ahead of the first warp

__device__ __forceinline__ void mbar_arrive(unsigned long long *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// shared-memory layout of one task (CTA)
template <int NU, bool POST>
struct Layout {
    static constexpr int NSB = NU + 1 + PFB;                       // b ring: blocks m-NU .. m live, PFB ahead
    static constexpr int NSX = 1 + PFB;                            // xin ring (first warp only)
    static constexpr int CSLOT = (CBLK_BYTES + 127) / 128 * 128;
    static constexpr int B_OFF = 0;
    static constexpr int X_OFF = B_OFF + NSB * BLK_BYTES;
    static constexpr int C_OFF = X_OFF + (POST ? NSX * BLK_BYTES : 0);
    static constexpr int H_OFF = C_OFF + (POST ? NSX * CSLOT : 0);
    static constexpr int NH = NU - 1;                              // hand-off rings
    static constexpr int BAR_OFF = H_OFF + NH * NHS * BLK_BYTES;
    // barriers: full_b[NSB], empty_b[NSB], full_x[NSX], then per hand-off ring full[NHS], empty[NHS]
    static constexpr int NBAR = 2 * NSB + NSX + NH * 2 * NHS;
    static constexpr int BYTES = BAR_OFF + NBAR * 8;
};

struct Ctx {
    unsigned char *smem;
    unsigned long long *bars;
    int lane, nblk;
    bool has_reg;
};

// the lane's two values of row `row` of a block
__device__ __forceinline__ double2 blk_ld(const unsigned char *blk, int row, int lane)
{
    return *reinterpret_cast<const double2 *>(blk + row * (SWID * 8) + lane * 16);
}
__device__ __forceinline__ void blk_st(unsigned char *blk, int row, int lane, double a, double b)
{
    *reinterpret_cast<double2 *>(blk + row * (SWID * 8) + lane * 16) = make_double2(a, b);
}

// right-hand side of the row a warp with lag LG emits at step u of block m: element 6m + u - LG of the b sequence
// (zero before the first staged row).  bcur / bprev / bprev2: the b blocks m, m-1, m-2.
template <int LG>
__device__ __forceinline__ double2 b_of(int u, int m, const unsigned char *bcur, const unsigned char *bprev,
                                        const unsigned char *bprev2, int lane)
{
    const int e = u - LG;                       // relative to the first element of block m
    if (e >= 0) return blk_ld(bcur, e, lane);
    if (e >= -RB) return m >= 1 ? blk_ld(bprev, e + RB, lane) : make_double2(0.0, 0.0);
    return m >= 2 ? blk_ld(bprev2, e + 2 * RB, lane) : make_double2(0.0, 0.0);
}

template <int NU, bool POST>
__device__ __forceinline__ const unsigned char *b_block(const Ctx &c, int m)
{
    using L = Layout<NU, POST>;
    return c.smem + L::B_OFF + (size_t)((m + L::NSB) % L::NSB) * BLK_BYTES;
}

// ---- middle warp k (1 <= k <= NU-2): x_{k+1} -> x_{k+2} (pre) / sweep k+1 (post); also the post-smoother's last warp ----
template <int NU, int K, bool POST, bool LAST, bool DOT>
__device__ __forceinline__ void role_mid(const Ctx &c, const LevelDev &F, const Task &T, const Coef<NU> &ca, const Coef<NU> &cb,
                                         const IrrTab<NU> &irr, double *__restrict__ x, double &dot)
{
    using L = Layout<NU, POST>;
    constexpr int LG = 2 * (K + 1), WI = POST ? K : K + 1;   // weight index of this warp's sweep
    unsigned long long *full_b = c.bars, *empty_b = c.bars + L::NSB;
    unsigned long long *full_in = c.bars + 2 * L::NSB + L::NSX + (K - 1) * 2 * NHS, *empty_in = full_in + NHS;
    unsigned long long *full_out = c.bars + 2 * L::NSB + L::NSX + K * 2 * NHS, *empty_out = full_out + NHS;
    unsigned char *hin = c.smem + L::H_OFF + (size_t)(K - 1) * NHS * BLK_BYTES;
    unsigned char *hout = c.smem + L::H_OFF + (size_t)K * NHS * BLK_BYTES;
    const int lane = c.lane;
    double win[3][4];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int q = 0; q < 4; ++q) win[r][q] = 0.0;
    for (int m = 0; m < c.nblk; ++m) {
        const int hs = m % NHS, i0 = T.y0 + m * RB - 2 * K;   // fine row of the first entering element
        mbar_wait(full_in + hs, (m / NHS) & 1);
        if (!LAST) mbar_wait(empty_out + hs, ((m / NHS) & 1) ^ 1);
        mbar_wait(full_b + (m % L::NSB), (m / L::NSB) & 1);   // rows of b up to block m are in place
        const unsigned char *bi = hin + (size_t)hs * BLK_BYTES;
        unsigned char *bo_ = hout + (size_t)hs * BLK_BYTES;
        const unsigned char *b0 = b_block<NU, POST>(c, m), *b1 = b_block<NU, POST>(c, m - 1), *b2 = b_block<NU, POST>(c, m - 2);
        // every input row of the block is loaded before the first output row is stored: the hand-off rings and the b ring
        // are the same shared array to the compiler, so a load after a store could not be moved above it and the six
        // steps of a block would run one after the other (950 cycles per step in the first version of this kernel)
        double2 xin[RB], bin[RB];
#pragma unroll
        for (int u = 0; u < RB; ++u) { xin[u] = blk_ld(bi, u, lane); bin[u] = b_of<LG>(u, m, b0, b1, b2, lane); }
        auto body = [&](auto fc) {
            constexpr bool fast = decltype(fc)::value;
#pragma unroll
            for (int u = 0; u < RB; ++u) {
                const int r = i0 + u - 2;
                Coef<NU> la, lb;
                if (fast) { la = ca; lb = cb; } else row_coef<NU>(F, r, ca, cb, irr, la, lb);
                const double2 bv = bin[u];
                double ra, rbv;
                node_pair<NU>(win, la, lb, bv.x, bv.y, ra, rbv);
                const double va = fma(la.wic[WI], ra, win[1][1]), vb = fma(lb.wic[WI], rbv, win[1][2]);
                if (!LAST) blk_st(bo_, u, lane, va, vb);
                else if (T.own_cols && r >= T.lo_y && r < T.hi_y) {
                    store_pair(x, F, r, T.j0, va, vb);
                    if (DOT) dot += va * bv.x + (T.j0 + 1 < F.nx ? vb * bv.y : 0.0);
                }
                push_row(win, xin[u].x, xin[u].y);
            }
        };
        if (c.has_reg && i0 - 3 >= 1 && i0 + RB - 1 <= F.ireg_hi) body(cuda::std::true_type{});
        else body(cuda::std::false_type{});
        __syncwarp();
        if (lane == 0) {
            if (!LAST) mbar_arrive(full_out + hs);
            mbar_arrive(empty_in + hs);
            // the last warp has read every row of b before block m + 1 - ceil(2NU/6): release one block per iteration
            if (LAST && m >= (2 * NU + RB - 1) / RB) mbar_arrive(empty_b + ((m - (2 * NU + RB - 1) / RB) % L::NSB));
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------------
// pre-smoothing pipeline: NU warps
// ---------------------------------------------------------------------------------------------------------------------------
template <int NU>
__global__ void __launch_bounds__(32 * NU)
kp_presmooth(LevelDev F, LevelDev Cc, const __grid_constant__ CUtensorMap map_b, double *__restrict__ x,
             double *__restrict__ bc, SmoothW sw, StreamGeom G, const CGScalars *sc)
{
    pdl_trigger();
    using L = Layout<NU, false>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int stage = threadIdx.x >> 5, lane = threadIdx.x & 31;
    Ctx c;
    c.smem = smem_raw;
    c.bars = reinterpret_cast<unsigned long long *>(smem_raw + L::BAR_OFF);
    c.lane = lane;
    c.nblk = G.nblk;
    c.has_reg = F.ireg_hi >= 1;
    unsigned long long *full_b = c.bars, *empty_b = c.bars + L::NSB;
    const Task T = make_task(F, G, blockIdx.x, blockIdx.y);
    if (threadIdx.x == 0) {
        for (int q = 0; q < L::NBAR; ++q) mbar_init(c.bars + q, 1);
        fence_barrier_init();
    }
    // the lane's coefficient sets (regular row) and the irregular-row table: set-up constants only
    Coef<NU> ca, cb;
    regular_coefs<NU>(F, T.j0, sw, ca, cb);
    IrrTab<NU> irr;
    fill_irr<NU>(F, T.j0, sw, T.y0 - 2 * NU - 4, T.y0 + G.nblk * RB, irr);
    __syncthreads();   // barriers initialised (the only block-wide barrier of the kernel)
    pdl_wait();
    if (sc->done) return;
    double dummy = 0.0;

    if (stage == 0) {
        // ---- warp 0: b -> x1 (pointwise) -> x2 ------------------------------------------------------------------------------
        unsigned long long *full_out = c.bars + 2 * L::NSB + L::NSX, *empty_out = full_out + NHS;
        unsigned char *hout = smem_raw + L::H_OFF;
        if (lane == 0)
            for (int q = 0; q <= PFB && q < c.nblk; ++q) {
                mbar_expect_tx(full_b + q, BLK_BYTES);
                tma_load_2d(smem_raw + L::B_OFF + (size_t)q * BLK_BYTES, &map_b, full_b + q, T.ox, T.y0 + q * RB);
            }
        double win[3][4];
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int q = 0; q < 4; ++q) win[r][q] = 0.0;
        for (int m = 0; m < c.nblk; ++m) {
            const int slot = m % L::NSB, hs = m % NHS, i0 = T.y0 + m * RB;
            mbar_wait(full_b + slot, (m / L::NSB) & 1);
            mbar_wait(empty_out + hs, ((m / NHS) & 1) ^ 1);
            const unsigned char *b0 = b_block<NU, false>(c, m), *b1 = b_block<NU, false>(c, m - 1);
            unsigned char *bo_ = hout + (size_t)hs * BLK_BYTES;
            double2 bin[RB], bin2[RB];   // all loads of the block before its first store (see role_mid)
#pragma unroll
            for (int u = 0; u < RB; ++u) { bin[u] = blk_ld(b0, u, lane); bin2[u] = b_of<2>(u, m, b0, b1, b1, lane); }
            auto body = [&](auto fc) {
                constexpr bool fast = decltype(fc)::value;
#pragma unroll
                for (int u = 0; u < RB; ++u) {
                    const int i = i0 + u;
                    Coef<NU> la, lb;
                    // x2 on row i-2 from the x1 window (rows i-3 .. i-1)
                    if (fast) { la = ca; lb = cb; } else row_coef<NU>(F, i - 2, ca, cb, irr, la, lb);
                    const double2 bv2 = bin2[u];
                    double ra, rbv;
                    node_pair<NU>(win, la, lb, bv2.x, bv2.y, ra, rbv);
                    blk_st(bo_, u, lane, fma(la.wic[1], ra, win[1][1]), fma(lb.wic[1], rbv, win[1][2]));
                    // x1 on the entering row i
                    if (!fast) row_coef<NU>(F, i, ca, cb, irr, la, lb);
                    push_row(win, la.wic[0] * bin[u].x, lb.wic[0] * bin[u].y);
                }
            };
            if (c.has_reg && i0 - 3 >= 1 && i0 + RB - 1 <= F.ireg_hi) body(cuda::std::true_type{});
            else body(cuda::std::false_type{});
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(full_out + hs);
                const int mp = m + PFB + 1;   // next block to stage: its slot was last used by block mp - NSB
                if (mp < c.nblk) {
                    const int sp = mp % L::NSB;
                    mbar_wait(empty_b + sp, ((mp / L::NSB) & 1) ^ 1);
                    fence_proxy_async();
                    mbar_expect_tx(full_b + sp, BLK_BYTES);
                    tma_load_2d(smem_raw + L::B_OFF + (size_t)sp * BLK_BYTES, &map_b, full_b + sp, T.ox, T.y0 + mp * RB);
                }
            }
            __syncwarp();
        }
    } else if (stage < NU - 1) {
        if (NU == 3 || stage == 1) role_mid<NU, 1, false, false, false>(c, F, T, ca, cb, irr, nullptr, dummy);
        else role_mid<NU, (NU > 3 ? 2 : 1), false, false, false>(c, F, T, ca, cb, irr, nullptr, dummy);
    } else {
        // ---- last warp: x_NU -> residual, restriction, stores ------------------------------------------------------------------
        constexpr int K = NU - 1, LG = 2 * NU;
        unsigned long long *full_in = c.bars + 2 * L::NSB + L::NSX + (K - 1) * 2 * NHS, *empty_in = full_in + NHS;
        unsigned char *hin = smem_raw + L::H_OFF + (size_t)(K - 1) * NHS * BLK_BYTES;
        const bool cj0 = T.j0 >= 0 && T.j0 < F.nx;
        const double wE = (T.j0 + 1 < F.nx && is_mid(T.j0 + 1, F.nx)) ? 0.5 : 0.0;
        const double wW = (T.j0 - 1 >= 0 && is_mid(T.j0 - 1, F.nx)) ? 0.5 : 0.0;
        const bool cj1 = (T.j0 + 1 == F.nx - 1) && !is_mid(T.j0 + 1, F.nx) && T.j0 + 1 >= T.lo_x && T.j0 + 1 < T.hi_x;
        const int J0 = T.j0 >> 1;
        double win[3][4], rw[3][3];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            rw[r][0] = rw[r][1] = rw[r][2] = 0.0;
#pragma unroll
            for (int q = 0; q < 4; ++q) win[r][q] = 0.0;
        }
        for (int m = 0; m < c.nblk; ++m) {
            const int hs = m % NHS, i0 = T.y0 + m * RB - 2 * K;
            mbar_wait(full_in + hs, (m / NHS) & 1);
            mbar_wait(full_b + (m % L::NSB), (m / L::NSB) & 1);
            const unsigned char *bi = hin + (size_t)hs * BLK_BYTES;
            const unsigned char *b0 = b_block<NU, false>(c, m), *b1 = b_block<NU, false>(c, m - 1), *b2 = b_block<NU, false>(c, m - 2);
            double2 xin[RB], bin[RB];
#pragma unroll
            for (int u = 0; u < RB; ++u) { xin[u] = blk_ld(bi, u, lane); bin[u] = b_of<LG>(u, m, b0, b1, b2, lane); }
            auto body = [&](auto fc) {
                constexpr bool fast = decltype(fc)::value;
#pragma unroll
                for (int u = 0; u < RB; ++u) {
                    const int r = i0 + u - 2;   // the row whose residual is formed
                    Coef<NU> la, lb;
                    if (fast) { la = ca; lb = cb; } else row_coef<NU>(F, r, ca, cb, irr, la, lb);
                    const double2 bv = bin[u];
                    double ra, rbv;
                    node_pair<NU>(win, la, lb, bv.x, bv.y, ra, rbv);
                    if (T.own_cols && r >= T.lo_y && r < T.hi_y) store_pair(x, F, r, T.j0, win[1][1], win[1][2]);
#pragma unroll
                    for (int q = 0; q < 3; ++q) { rw[0][q] = rw[1][q]; rw[1][q] = rw[2][q]; }
                    rw[2][1] = ra; rw[2][2] = rbv;
                    rw[2][0] = __shfl_up_sync(0xffffffffu, rbv, 1);
                    // restriction at rc = r - 1 (middle residual row) when it is a coarse row; y0 is even, so rc is even
                    // exactly when u is odd
                    const int rc = r - 1;
                    const bool even = fast ? ((u & 1) == 1) : ((rc & 1) == 0);
                    if ((even || (!fast && rc == F.ny - 1)) && rc >= T.lo_y && rc < T.hi_y) {
                        const int I = coarse_lo(rc, F.ny, Cc.ny);
                        const double wN = fast ? 0.5 : ((rc + 1 < F.ny && is_mid(rc + 1, F.ny)) ? 0.5 : 0.0);
                        const double wS = fast ? 0.5 : ((rc - 1 >= 0 && is_mid(rc - 1, F.ny)) ? 0.5 : 0.0);
                        if (T.own_cols && cj0) {
                            double v = rw[1][1] + wE * rw[1][2] + wW * rw[1][0] + wN * rw[2][1] + wS * rw[0][1] +
                                       (2.0 * wN * wE) * rw[2][2] + (2.0 * wS * wW) * rw[0][0];
                            if (is_dirichlet(Cc, I, J0)) v = 0.0;
                            bc[(size_t)I * Cc.nx + J0] = v;
                        }
                        if (cj1) {
                            double v = rw[1][2] + wN * rw[2][2] + wS * rw[0][2];
                            if (is_dirichlet(Cc, I, Cc.nx - 1)) v = 0.0;
                            bc[(size_t)I * Cc.nx + Cc.nx - 1] = v;
                        }
                    }
                    push_row(win, xin[u].x, xin[u].y);
                }
            };
            if (c.has_reg && i0 - 4 >= 1 && i0 + RB - 1 <= F.ireg_hi) body(cuda::std::true_type{});
            else body(cuda::std::false_type{});
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(empty_in + hs);
                if (m >= (LG + RB - 1) / RB) mbar_arrive(empty_b + ((m - (LG + RB - 1) / RB) % L::NSB));
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------------
// post-smoothing pipeline: NU warps.  Warp 0 prolongates (x0 = xin + P xc) and does sweep 1, warp k sweep k+1, the last
// warp stores the result and sums x.b
// ---------------------------------------------------------------------------------------------------------------------------
template <int NU, bool DOT>
__global__ void __launch_bounds__(32 * NU)
kp_postsmooth(LevelDev F, LevelDev Cc, const __grid_constant__ CUtensorMap map_b, const __grid_constant__ CUtensorMap map_x,
              double *__restrict__ x, const double *__restrict__ xc, SmoothW sw, StreamGeom G, CGScalars *sc,
              double *partials, unsigned *counter, double *out_dot)
{
    pdl_trigger();
    using L = Layout<NU, true>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int stage = threadIdx.x >> 5, lane = threadIdx.x & 31;
    Ctx c;
    c.smem = smem_raw;
    c.bars = reinterpret_cast<unsigned long long *>(smem_raw + L::BAR_OFF);
    c.lane = lane;
    c.nblk = G.nblk;
    c.has_reg = F.ireg_hi >= 1;
    unsigned long long *full_b = c.bars, *empty_b = c.bars + L::NSB, *full_x = c.bars + 2 * L::NSB;
    const Task T = make_task(F, G, blockIdx.x, blockIdx.y);
    if (threadIdx.x == 0) {
        for (int q = 0; q < L::NBAR; ++q) mbar_init(c.bars + q, 1);
        fence_barrier_init();
    }
    Coef<NU> ca, cb;
    regular_coefs<NU>(F, T.j0, sw, ca, cb);
    IrrTab<NU> irr;
    fill_irr<NU>(F, T.j0, sw, T.y0 - 2 * NU - 4, T.y0 + G.nblk * RB, irr);
    __syncthreads();
    pdl_wait();
    double dot = 0.0;
    if (!sc->done) {
        if (stage == 0) {
            unsigned long long *full_out = c.bars + 2 * L::NSB + L::NSX, *empty_out = full_out + NHS;
            unsigned char *hout = smem_raw + L::H_OFF, *rx = smem_raw + L::X_OFF, *rcs = smem_raw + L::C_OFF;
            const int J0 = T.j0 >> 1, CJ0 = T.ox >> 1;
            const bool in0 = T.j0 >= 0 && T.j0 < F.nx, in1 = T.j0 + 1 >= 0 && T.j0 + 1 < F.nx;
            const bool mid1 = in1 && is_mid(T.j0 + 1, F.nx);
            const double m0 = (in0 && !(((F.dirmask & 1u) && T.j0 == 0) || ((F.dirmask & 2u) && T.j0 == F.nx - 1))) ? 1.0 : 0.0;
            const double m1 = (in1 && !(((F.dirmask & 1u) && T.j0 + 1 == 0) || ((F.dirmask & 2u) && T.j0 + 1 == F.nx - 1))) ? 1.0 : 0.0;
            auto stage_in = [&](int q) {   // block q of b, xin (TMA) and of the coarse patch (cp.async)
                const int sb = q % L::NSB, sx = q % L::NSX;
                if (lane == 0) {
                    mbar_expect_tx(full_b + sb, BLK_BYTES);
                    tma_load_2d(smem_raw + L::B_OFF + (size_t)sb * BLK_BYTES, &map_b, full_b + sb, T.ox, T.y0 + q * RB);
                    mbar_expect_tx(full_x + sx, BLK_BYTES);
                    tma_load_2d(rx + (size_t)sx * BLK_BYTES, &map_x, full_x + sx, T.ox, T.y0 + q * RB);
                }
                double *dst = reinterpret_cast<double *>(rcs + (size_t)sx * L::CSLOT);
                const int I0 = (T.y0 + q * RB) >> 1;
                for (int e = lane; e < CROWS * CCOLS; e += 32) {
                    const int row = e / CCOLS, col = e - row * CCOLS;
                    const int I = I0 + row, J = CJ0 + col;
                    if (I >= 0 && I < Cc.ny && J >= 0 && J < Cc.nx) cp_async8(dst + e, xc + (size_t)I * Cc.nx + J);
                    else dst[e] = 0.0;
                }
            };
            for (int q = 0; q <= PFB; ++q) {
                if (q < c.nblk) stage_in(q);
                cp_async_commit();
            }
            double win[3][4];
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int q = 0; q < 4; ++q) win[r][q] = 0.0;
            for (int m = 0; m < c.nblk; ++m) {
                const int sx = m % L::NSX, hs = m % NHS, i0 = T.y0 + m * RB;
                cp_async_wait<PFB>();
                mbar_wait(full_b + (m % L::NSB), (m / L::NSB) & 1);
                mbar_wait(full_x + sx, (m / L::NSX) & 1);
                mbar_wait(empty_out + hs, ((m / NHS) & 1) ^ 1);
                __syncwarp();
                const double *cp = reinterpret_cast<const double *>(rcs + (size_t)sx * L::CSLOT) + (J0 - CJ0);
                const unsigned char *b0 = b_block<NU, true>(c, m), *b1 = b_block<NU, true>(c, m - 1);
                const unsigned char *xi = rx + (size_t)sx * BLK_BYTES;
                unsigned char *bo_ = hout + (size_t)hs * BLK_BYTES;
                double2 xin[RB], bin[RB];   // all loads of the block before its first store (see role_mid)
                double cl[CROWS][2];        // the lane's coarse column and its east neighbour, the rows under the block
#pragma unroll
                for (int u = 0; u < RB; ++u) { xin[u] = blk_ld(xi, u, lane); bin[u] = b_of<2>(u, m, b0, b1, b1, lane); }
#pragma unroll
                for (int q = 0; q < CROWS; ++q) { cl[q][0] = cp[q * CCOLS]; cl[q][1] = cp[q * CCOLS + 1]; }
                auto body = [&](auto fc) {
                    constexpr bool fast = decltype(fc)::value;
#pragma unroll
                    for (int u = 0; u < RB; ++u) {
                        const int i = i0 + u;
                        Coef<NU> la, lb;
                        // sweep 1 on row i-2 from the window of the prolongated iterate
                        if (fast) { la = ca; lb = cb; } else row_coef<NU>(F, i - 2, ca, cb, irr, la, lb);
                        const double2 bv = bin[u];
                        double ra, rbv;
                        node_pair<NU>(win, la, lb, bv.x, bv.y, ra, rbv);
                        blk_st(bo_, u, lane, fma(la.wic[0], ra, win[1][1]), fma(lb.wic[0], rbv, win[1][2]));
                        // prolongation on the entering row i
                        const bool rowin = fast || (i >= 0 && i < F.ny);
                        const bool odd = fast ? ((u & 1) != 0) : ((i & 1) != 0);
                        const bool mi = odd && (fast || i != F.ny - 1);
                        // coarse row under row i: patch row u/2; an odd last row sits on the patch row above
                        const int q0 = (u >> 1) + ((odd && !mi) ? 1 : 0);
                        double pa, pb;
                        if (!mi) {
                            pa = cl[q0][0];
                            pb = mid1 ? 0.5 * (cl[q0][0] + cl[q0][1]) : cl[q0][1];
                        } else {
                            pa = 0.5 * (cl[q0][0] + cl[q0 + 1][0]);
                            pb = mid1 ? 0.5 * (cl[q0][0] + cl[q0 + 1][1]) : 0.5 * (cl[q0][1] + cl[q0 + 1][1]);
                        }
                        double rm = rowin ? 1.0 : 0.0;
                        if (!fast && rowin && (((F.dirmask & 8u) && i == 0) || ((F.dirmask & 4u) && i == F.ny - 1))) rm = 0.0;
                        push_row(win, rm * m0 * (xin[u].x + pa), rm * m1 * (xin[u].y + pb));
                    }
                };
                if (c.has_reg && i0 - 3 >= 1 && i0 + RB - 1 <= F.ireg_hi) body(cuda::std::true_type{});
                else body(cuda::std::false_type{});
                __syncwarp();
                if (lane == 0) mbar_arrive(full_out + hs);
                const int mp = m + PFB + 1;
                if (mp < c.nblk) {
                    if (lane == 0) {
                        mbar_wait(empty_b + (mp % L::NSB), ((mp / L::NSB) & 1) ^ 1);
                        fence_proxy_async();
                    }
                    __syncwarp();
                    stage_in(mp);
                }
                cp_async_commit();
            }
        } else if (stage == NU - 1) {
            role_mid<NU, NU - 1, true, true, DOT>(c, F, T, ca, cb, irr, x, dot);
        } else {
            if (NU == 3 || stage == 1) role_mid<NU, 1, true, false, false>(c, F, T, ca, cb, irr, nullptr, dot);
            else role_mid<NU, (NU > 3 ? 2 : 1), true, false, false>(c, F, T, ca, cb, irr, nullptr, dot);
        }
    }
    if (DOT) {
        double v[1] = {dot}, tot[1];
        if (grid_reduce<1>(v, partials, counter, tot) && !sc->done) *out_dot = tot[0];
    }
}

}  // namespace PIPE
