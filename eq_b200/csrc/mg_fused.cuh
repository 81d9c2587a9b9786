// Fused multigrid kernels (isotropic operator).
//
// HBM traffic bounds the V-cycle, so every level is touched by exactly two
// kernels per cycle, each doing all of its smoothing sweeps on a shared-memory
// tile whose halo is as deep as the sweep count (temporal blocking):
//   k_presmooth : NU Chebyshev-weighted Jacobi sweeps from a zero guess +
//                 residual + restriction.   reads b        writes x, b_coarse
//   k_postsmooth: prolongate-add + NU sweeps (+ r.z).
//                                           reads x, b, x_coarse   writes x
// The deepest levels, small enough for one CTA's shared memory, are all handled
// by the single-CTA k_tail.  p = z + beta*p is fused into the operator apply
// (k_apply_p).
//
// Tile layout: a 64x64-node shared tile (TS) with a one-node pad ring, 512
// threads = 64 columns x 8 row chunks.  A thread owns one column of an 8-row
// chunk, keeps its b values in registers for all sweeps, and walks the column
// upwards with the 3x3 neighbourhood in registers, so one node update costs
// 3 LDS (row above) + 1 STS.  Two shared arrays (ping-pong) per CTA.  Tiles that contain
// only regular interior nodes (all but the perimeter tiles) take a branch-free
// constant-coefficient path.
#pragma once
#include "eqgpu_internal.cuh"

#define TAIL_THREADS 1024
#define MAX_LEVELS 16
#define MAX_CHEB 24

struct CoarseW { int n; double w[MAX_CHEB]; };

// coarse node J <-> fine node min(2J, nf-1)
__device__ __forceinline__ int fine_of(int J, int nf) { return min(2 * J, nf - 1); }
// fine index f is a midpoint node (odd and not the last node)?
__device__ __forceinline__ bool is_mid(int f, int nf) { return (f & 1) && f != nf - 1; }
// coarse index of the coincident node (or of the lower neighbour of a midpoint)
__device__ __forceinline__ int coarse_lo(int f, int nf, int nc) { return f == nf - 1 ? nc - 1 : f >> 1; }

__device__ __forceinline__ double inv_diag(const LevelDev &L, int i, int j, const double c[NBAND])
{
    return (i >= 1 && i <= L.ireg_hi && j >= 1 && j <= L.jreg_hi) ? L.icC : 1.0 / c[B_C];
}

// value of P*xc at fine node (gi,gj), coarse values read from global memory
__device__ __forceinline__ double prolong_at(const LevelDev &F, const LevelDev &Cc,
                                             const double *__restrict__ xc, int gi, int gj)
{
    const bool mi = is_mid(gi, F.ny), mj = is_mid(gj, F.nx);
    const double *c0 = xc + (size_t)coarse_lo(gi, F.ny, Cc.ny) * Cc.nx + coarse_lo(gj, F.nx, Cc.nx);
    if (!mi && !mj) return __ldg(c0);
    if (!mi && mj) return 0.5 * (__ldg(c0) + __ldg(c0 + 1));
    if (mi && !mj) return 0.5 * (__ldg(c0) + __ldg(c0 + Cc.nx));
    return 0.5 * (__ldg(c0) + __ldg(c0 + Cc.nx + 1));
}

// Debug build only (-DEQ_KTRACE): clock64() marks of CTA 0 of the pre-smoothing kernels, slot base by level width
#ifdef EQ_KTRACE
__device__ unsigned long long g_ktrace[256];
#define KT_BASE(F) ((F).nx <= 130 ? 0 : (F).nx <= 258 ? 32 : (F).nx <= 514 ? 64 : (F).nx <= 1026 ? 96 : 128)
#define KT(F, i) do { if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) g_ktrace[KT_BASE(F) + (i)] = clock64(); } while (0)
#else
#define KT(F, i) do { } while (0)
#endif

#define TS 64
#define TSL 6
namespace T64 {
#include "mg_tile.inc"
}  // namespace T64
#undef TS
#undef TSL
#define TS 32
#define TSL 5
namespace T32 {
#include "mg_tile.inc"
}  // namespace T32
#undef TS
#undef TSL

// ---------------------------------------------------------------------------
// single-CTA tail: the whole V-cycle below level `first`, in shared memory
// ---------------------------------------------------------------------------
struct TailDesc {
    int first, last;            // level range [first, last]
    int off[MAX_LEVELS];        // start of the level's three arrays in the smem block (doubles)
    int soff[MAX_LEVELS];       // start of the level's cell-size copies (hx, ihx, hy, ihy)
    int total;                  // doubles of shared memory in use
};

// Tail levels live in shared memory as padded arrays: row stride nx+2, a zero
// ring around the grid, node (i,j) at (i+1)*(nx+2) + j+1.
__device__ __forceinline__ int lidx(const LevelDev &L, int i, int j) { return (i + 1) * (L.nx + 2) + j + 1; }

// One sweep over a whole tail level.  All nodes take the constant-coefficient
// update (warps walk rows; lanes walk columns), then the irregular nodes (rows
// 0, ireg_hi+1.., columns 0, jreg_hi+1.., Dirichlet nodes) are recomputed with
// the general row.
template <int MODE>
__device__ __forceinline__ void lvl_sweep(const LevelDev &L, const Spacing &S, const double *b, const double *src,
                                          double *dst, double w)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, sx = L.nx + 2;
    const double cC = L.cC, cEW = L.cEW, cNS = L.cNS, cD = L.cD, wd = w * L.icC;
    for (int i = warp; i < L.ny; i += TAIL_THREADS / 32) {
        for (int j = lane; j < L.nx; j += 32) {
            const int c = (i + 1) * sx + j + 1;
            double out;
            if (MODE == 0) out = wd * b[c];
            else {
                const double ax = cC * src[c] + cEW * (src[c + 1] + src[c - 1]) + cNS * (src[c + sx] + src[c - sx]) +
                                  cD * (src[c + sx + 1] + src[c - sx - 1]);
                const double res = b[c] - ax;
                out = MODE == 2 ? res : src[c] + wd * res;
            }
            dst[c] = out;
        }
    }
    __syncthreads();
    const int nr = L.ny - L.ireg_hi, ncol = L.nx - L.jreg_hi;
    const int span = max(L.nx, L.ny);
    const int items = (nr + ncol) * span;
    for (int it = threadIdx.x; it < items; it += TAIL_THREADS) {
        const int k = it / span, e = it - k * span;
        int i, j;
        if (k < nr) { i = k == 0 ? 0 : L.ireg_hi + k; j = e; }
        else { const int kk = k - nr; j = kk == 0 ? 0 : L.jreg_hi + kk; i = e; }
        if (i >= L.ny || j >= L.nx) continue;
        const int c = (i + 1) * sx + j + 1;
        double out = 0.0;
        if (!is_dirichlet(L, i, j)) {
            double cf[NBAND];
            stencil_iso(L, S, i, j, cf);
            const double id = inv_diag(L, i, j, cf);
            if (MODE == 0) out = w * id * b[c];
            else {
                const double ax = cf[B_C] * src[c] + cf[B_E] * src[c + 1] + cf[B_W] * src[c - 1] +
                                  cf[B_N] * src[c + sx] + cf[B_S] * src[c - sx] + cf[B_NE] * src[c + sx + 1] +
                                  cf[B_SW] * src[c - sx - 1];
                const double res = b[c] - ax;
                out = MODE == 2 ? res : src[c] + w * id * res;
            }
        }
        dst[c] = out;
    }
    __syncthreads();
}

// `sweeps` weighted-Jacobi sweeps with weights w[0..]; FROM_ZERO starts with the
// w[0]*D^-1 b sweep.  x holds the result.
template <bool FROM_ZERO>
__device__ __forceinline__ void lvl_smooth(const LevelDev &L, const Spacing &S, const double *b, double *x,
                                           double *t, int sweeps, const double *w)
{
    double *cur = x, *oth = t;
    int k = 0;
    if (FROM_ZERO) {
        double *first = (sweeps & 1) ? x : t;  // so that the last write lands in x
        lvl_sweep<0>(L, S, b, nullptr, first, w[0]);
        cur = first; oth = (first == x) ? t : x;
        k = 1;
    } else if (sweeps & 1) {
        const int n = (L.nx + 2) * (L.ny + 2);
        for (int g = threadIdx.x; g < n; g += TAIL_THREADS) t[g] = x[g];
        __syncthreads();
        cur = t; oth = x;
    }
    for (; k < sweeps; ++k) {
        lvl_sweep<1>(L, S, b, cur, oth, w[k]);
        double *s = cur; cur = oth; oth = s;
    }
}

__global__ void __launch_bounds__(TAIL_THREADS)
k_tail(const LevelDev *__restrict__ levels, TailDesc td, const double *__restrict__ b_in,
       double *__restrict__ x_out, int nu, SmoothW sw, CoarseW cw, const CGScalars *sc)
{
    if (sc->done) return;
    extern __shared__ double sm[];
    auto PN = [&](int l) { return (levels[l].nx + 2) * (levels[l].ny + 2); };
    auto X = [&](int l) { return sm + td.off[l]; };
    auto B = [&](int l) { return sm + td.off[l] + PN(l); };
    auto T = [&](int l) { return sm + td.off[l] + 2 * PN(l); };
    auto SP = [&](int l) {
        const LevelDev &L = levels[l];
        double *q = sm + td.soff[l];
        Spacing S; S.hx = q; S.ihx = q + L.nx + 1; S.hy = q + 2 * (L.nx + 1); S.ihy = S.hy + L.ny + 1;
        S.jo = 0; S.io = 0;
        return S;
    };
    // zero everything once (pad rings must be zero), stage the cell sizes and the right-hand side
    for (int g = threadIdx.x; g < td.total; g += TAIL_THREADS) sm[g] = 0.0;
    __syncthreads();
    for (int l = td.first; l <= td.last; ++l) {
        const LevelDev &L = levels[l];
        double *q = sm + td.soff[l];
        for (int k = threadIdx.x; k <= L.nx; k += TAIL_THREADS) { q[k] = L.hx[k]; q[L.nx + 1 + k] = L.ihx[k]; }
        double *qy = q + 2 * (L.nx + 1);
        for (int k = threadIdx.x; k <= L.ny; k += TAIL_THREADS) { qy[k] = L.hy[k]; qy[L.ny + 1 + k] = L.ihy[k]; }
    }
    {
        const LevelDev &L = levels[td.first];
        double *b0 = B(td.first);
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        for (int i = warp; i < L.ny; i += TAIL_THREADS / 32)
            for (int j = lane; j < L.nx; j += 32) b0[lidx(L, i, j)] = __ldg(b_in + (size_t)i * L.nx + j);
        __syncthreads();
    }
    for (int l = td.first; l < td.last; ++l) {
        const LevelDev &F = levels[l], &Cc = levels[l + 1];
        lvl_smooth<true>(F, SP(l), B(l), X(l), T(l), nu, sw.w);
        lvl_sweep<2>(F, SP(l), B(l), X(l), T(l), 0.0);
        const double *rf = T(l);
        double *bc = B(l + 1);
        const int nc = Cc.nx * Cc.ny, fs = F.nx + 2;
        for (int g = threadIdx.x; g < nc; g += TAIL_THREADS) {
            const int I = g / Cc.nx, J = g - I * Cc.nx;
            double out = 0.0;
            if (!is_dirichlet(Cc, I, J)) {
                const int fi = fine_of(I, F.ny), fj = fine_of(J, F.nx);
                const int c = lidx(F, fi, fj);
                const bool e = fj + 1 < F.nx && is_mid(fj + 1, F.nx), w = fj >= 1 && is_mid(fj - 1, F.nx);
                const bool n = fi + 1 < F.ny && is_mid(fi + 1, F.ny), s = fi >= 1 && is_mid(fi - 1, F.ny);
                double h = 0.0;
                if (e) h += rf[c + 1];
                if (w) h += rf[c - 1];
                if (n) h += rf[c + fs];
                if (s) h += rf[c - fs];
                if (n && e) h += rf[c + fs + 1];
                if (s && w) h += rf[c - fs - 1];
                out = rf[c] + 0.5 * h;
            }
            bc[lidx(Cc, I, J)] = out;
        }
        __syncthreads();
    }
    lvl_smooth<true>(levels[td.last], SP(td.last), B(td.last), X(td.last), T(td.last), cw.n, cw.w);
    for (int l = td.last - 1; l >= td.first; --l) {
        const LevelDev &F = levels[l], &Cc = levels[l + 1];
        double *xf = X(l);
        const double *xc = X(l + 1);
        const int n = F.nx * F.ny, cs = Cc.nx + 2;
        for (int g = threadIdx.x; g < n; g += TAIL_THREADS) {
            const int i = g / F.nx, j = g - i * F.nx;
            if (is_dirichlet(F, i, j)) continue;
            const bool mi = is_mid(i, F.ny), mj = is_mid(j, F.nx);
            const double *c0 = xc + lidx(Cc, coarse_lo(i, F.ny, Cc.ny), coarse_lo(j, F.nx, Cc.nx));
            double add;
            if (!mi && !mj) add = c0[0];
            else if (!mi && mj) add = 0.5 * (c0[0] + c0[1]);
            else if (mi && !mj) add = 0.5 * (c0[0] + c0[cs]);
            else add = 0.5 * (c0[0] + c0[cs + 1]);
            xf[lidx(F, i, j)] += add;
        }
        __syncthreads();
        lvl_smooth<false>(F, SP(l), B(l), X(l), T(l), nu, sw.w);
    }
    {
        const LevelDev &L = levels[td.first];
        const double *x0 = X(td.first);
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        for (int i = warp; i < L.ny; i += TAIL_THREADS / 32)
            for (int j = lane; j < L.nx; j += 32) x_out[(size_t)i * L.nx + j] = x0[lidx(L, i, j)];
    }
}
