// Cluster tail: the deepest multigrid levels (<= 257^2 at the 2048^2 mesh) run
// as ONE kernel on one thread-block cluster of 16 CTAs.  Every level is split
// into row slabs, one per CTA, resident in that CTA's shared memory; a sweep
// reads the row above/below a slab edge straight out of the neighbouring CTA's
// shared memory (DSMEM) and sweeps are separated by cluster barriers (~0.2 us)
// instead of kernel boundaries (~10 us each for these latency-bound levels).
#pragma once
#include <cooperative_groups.h>
#include "eqgpu_internal.cuh"
#include "mg_fused.cuh"

namespace cg = cooperative_groups;

#define CT_THREADS 1024
#define CT_MAX_CTAS 16

struct CTailDesc {
    int first, last;          // level range [first, last]
    int ncta;                 // CTAs in the cluster
    int rp[MAX_LEVELS];       // rows per CTA of each level
    int off[MAX_LEVELS];      // per-CTA offset (doubles) of the level's x|b|t slabs
    int soff[MAX_LEVELS];     // per-CTA offset of the level's cell-size copies
    int zoff;                 // per-CTA offset of a zero row
    int total;                // doubles of shared memory per CTA
};

struct CTCtx {
    cg::cluster_group cl;
    double *sm;               // this CTA's dynamic shared memory
    const LevelDev *levels;
    const CTailDesc *td;
    unsigned rank;
};

// Pointer to element j = -1 (the left pad) of row i of array `arr` (0 x, 1 b, 2 t)
// of level l, wherever in the cluster it lives; rows outside the grid read zeros.
__device__ __forceinline__ const double *ct_row(const CTCtx &c, int l, int arr, int i)
{
    const LevelDev &L = c.levels[l];
    if (i < 0 || i >= L.ny) return c.sm + c.td->zoff;
    const int rp = c.td->rp[l], owner = i / rp, sx = L.nx + 2;
    const double *p = c.sm + c.td->off[l] + (arr * rp + (i - owner * rp)) * sx;
    return owner == (int)c.rank ? p : c.cl.map_shared_rank(p, owner);
}
__device__ __forceinline__ double *ct_own(const CTCtx &c, int l, int arr, int i)
{
    const LevelDev &L = c.levels[l];
    const int rp = c.td->rp[l];
    return c.sm + c.td->off[l] + (arr * rp + (i - (int)c.rank * rp)) * (L.nx + 2);
}
__device__ __forceinline__ Spacing ct_spacing(const CTCtx &c, int l)
{
    const LevelDev &L = c.levels[l];
    double *q = c.sm + c.td->soff[l];
    Spacing S; S.hx = q; S.ihx = q + L.nx + 1; S.hy = q + 2 * (L.nx + 1); S.ihy = S.hy + L.ny + 1;
    S.jo = 0; S.io = 0;
    return S;
}

// One sweep of level l over this CTA's rows: dst = f(b, src); arrays by index.
template <int MODE>
__device__ __forceinline__ void ct_sweep(const CTCtx &c, int l, int src, int dst, double w)
{
    const LevelDev &L = c.levels[l];
    const int rp = c.td->rp[l], i0 = c.rank * rp, i1 = min(i0 + rp, L.ny);
    const double cC = L.cC, cEW = L.cEW, cNS = L.cNS, cD = L.cD, wd = w * L.icC;
    // thread -> (column j, row chunk ch): a thread walks its column over rc rows with the 3x3
    // neighbourhood in registers; only the rows just outside the slab can be remote (DSMEM) or zero
    const int sx = L.nx + 2, CW = (L.nx + 31) & ~31, nch = max(CT_THREADS / CW, 1);
    const int j = threadIdx.x % CW, ch = threadIdx.x / CW;
    const int nown = max(i1 - i0, 0), rc = (nown + nch - 1) / nch;
    const int a = i0 + ch * rc, b = min(a + rc, i1);
    const bool regj = j >= 1 && j <= L.jreg_hi;
    if (ch < nch && regj && a < b) {
        const int q = j + 1;
        const double *__restrict__ bb = ct_own(c, l, 1, i0) + q;
        double *__restrict__ dd = ct_own(c, l, dst, i0) + q;
        if (MODE == 0) {
            for (int r = a; r < b; ++r)
                if (r >= 1 && r <= L.ireg_hi) dd[(r - i0) * sx] = wd * bb[(r - i0) * sx];
        } else {
            const double *base = ct_own(c, l, src, i0) + q;
            const double *rowm1 = (a - 1 >= i0) ? base + (a - 1 - i0) * sx : ct_row(c, l, src, a - 1) + q;
            const double *rowb = (b < i1) ? base + (b - i0) * sx : ct_row(c, l, src, b) + q;
            double sw_ = rowm1[-1], s_ = rowm1[0];
            const double *m = base + (a - i0) * sx;
            double w_ = m[-1], c_ = m[0], e_ = m[1];
            for (int r = a; r < b; ++r) {
                const double *up = (r + 1 < b) ? base + (r + 1 - i0) * sx : rowb;
                const double nw = up[-1], n_ = up[0], ne = up[1];
                const double ax = cC * c_ + cEW * (e_ + w_) + cNS * (n_ + s_) + cD * (ne + sw_);
                const double res = bb[(r - i0) * sx] - ax;
                if (r >= 1 && r <= L.ireg_hi) dd[(r - i0) * sx] = MODE == 2 ? res : c_ + wd * res;
                sw_ = w_; s_ = c_; w_ = nw; c_ = n_; e_ = ne;
            }
        }
    }
    // irregular nodes of my rows (disjoint from the nodes written above, so no barrier in between):
    // rows 0, ireg_hi+1..ny-1 (whole row) and columns 0, jreg_hi+1..nx-1
    const Spacing S = ct_spacing(c, l);
    const int nr = L.ny - L.ireg_hi, ncol = L.nx - L.jreg_hi;
    const int items = nr * L.nx + ncol * nown;
    for (int it = threadIdx.x; it < items; it += CT_THREADS) {
        int i, j;
        if (it < nr * L.nx) { const int k = it / L.nx; j = it - k * L.nx; i = k == 0 ? 0 : L.ireg_hi + k; }
        else { const int e = it - nr * L.nx; const int kk = e / nown; i = i0 + (e - kk * nown); j = kk == 0 ? 0 : L.jreg_hi + kk; }
        if (i < i0 || i >= i1 || j >= L.nx) continue;
        const int q = j + 1;
        double out = 0.0;
        if (!is_dirichlet(L, i, j)) {
            double cf[NBAND];
            stencil_iso(L, S, i, j, cf);
            const double id = inv_diag(L, i, j, cf);
            const double bq = ct_own(c, l, 1, i)[q];
            if (MODE == 0) out = w * id * bq;
            else {
                const double *mid = ct_own(c, l, src, i), *up = ct_row(c, l, src, i + 1), *dn = ct_row(c, l, src, i - 1);
                const double ax = cf[B_C] * mid[q] + cf[B_E] * mid[q + 1] + cf[B_W] * mid[q - 1] + cf[B_N] * up[q] +
                                  cf[B_S] * dn[q] + cf[B_NE] * up[q + 1] + cf[B_SW] * dn[q - 1];
                const double res = bq - ax;
                out = MODE == 2 ? res : mid[q] + w * id * res;
            }
        }
        ct_own(c, l, dst, i)[q] = out;
    }
    c.cl.sync();
}

// `sweeps` weighted sweeps; FROM_ZERO starts from x = 0.  Returns the array index (0 or 2) holding the result.
template <bool FROM_ZERO>
__device__ __forceinline__ int ct_smooth(const CTCtx &c, int l, int xin, int sweeps, const double *w)
{
    int cur = xin, oth = xin == 0 ? 2 : 0, k = 0;
    if (FROM_ZERO) {
        ct_sweep<0>(c, l, cur, cur, w[0]);
        k = 1;
    }
    for (; k < sweeps; ++k) {
        ct_sweep<1>(c, l, cur, oth, w[k]);
        const int t = cur; cur = oth; oth = t;
    }
    return cur;
}

__global__ void __launch_bounds__(CT_THREADS, 1)
k_ctail(const LevelDev *__restrict__ levels, CTailDesc td, const double *__restrict__ b_in,
        double *__restrict__ x_out, int nu, SmoothW sw, CoarseW cw, const CGScalars *sc)
{
    extern __shared__ double sm[];
    CTCtx c{cg::this_cluster(), sm, levels, &td, 0};
    c.rank = c.cl.block_rank();
    if (sc->done) return;  // uniform across the cluster
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int g = threadIdx.x; g < td.total; g += CT_THREADS) sm[g] = 0.0;
    __syncthreads();
    for (int l = td.first; l <= td.last; ++l) {
        const LevelDev &L = levels[l];
        double *q = sm + td.soff[l];
        for (int k = threadIdx.x; k <= L.nx; k += CT_THREADS) { q[k] = L.hx[k]; q[L.nx + 1 + k] = L.ihx[k]; }
        double *qy = q + 2 * (L.nx + 1);
        for (int k = threadIdx.x; k <= L.ny; k += CT_THREADS) { qy[k] = L.hy[k]; qy[L.ny + 1 + k] = L.ihy[k]; }
    }
    {   // right-hand side of the first level: my rows
        const LevelDev &L = levels[td.first];
        const int rp = td.rp[td.first], i0 = c.rank * rp, i1 = min(i0 + rp, L.ny);
        for (int i = i0 + warp; i < i1; i += CT_THREADS / 32) {
            double *bl = ct_own(c, td.first, 1, i);
            for (int j = lane; j < L.nx; j += 32) bl[j + 1] = __ldg(b_in + (size_t)i * L.nx + j);
        }
    }
    c.cl.sync();
    int xarr[MAX_LEVELS];
    for (int l = td.first; l < td.last; ++l) {
        const LevelDev &F = levels[l], &Cc = levels[l + 1];
        const int xa = ct_smooth<true>(c, l, 0, nu, sw.w);
        const int ra = xa == 0 ? 2 : 0;
        ct_sweep<2>(c, l, xa, ra, 0.0);
        xarr[l] = xa;
        // restriction into my rows of the coarse right-hand side
        const int rp = td.rp[l + 1], I0 = c.rank * rp, I1 = min(I0 + rp, Cc.ny);
        for (int I = I0 + warp; I < I1; I += CT_THREADS / 32) {
            const int fi = fine_of(I, F.ny);
            const bool n = fi + 1 < F.ny && is_mid(fi + 1, F.ny), s = fi >= 1 && is_mid(fi - 1, F.ny);
            const double *mid = ct_row(c, l, ra, fi), *up = ct_row(c, l, ra, fi + 1), *dn = ct_row(c, l, ra, fi - 1);
            double *bc = ct_own(c, l + 1, 1, I);
            for (int J = lane; J < Cc.nx; J += 32) {
                double out = 0.0;
                if (!is_dirichlet(Cc, I, J)) {
                    const int fj = fine_of(J, F.nx), q = fj + 1;
                    const bool e = fj + 1 < F.nx && is_mid(fj + 1, F.nx), w = fj >= 1 && is_mid(fj - 1, F.nx);
                    double h = 0.0;
                    if (e) h += mid[q + 1];
                    if (w) h += mid[q - 1];
                    if (n) h += up[q];
                    if (s) h += dn[q];
                    if (n && e) h += up[q + 1];
                    if (s && w) h += dn[q - 1];
                    out = mid[q] + 0.5 * h;
                }
                bc[J + 1] = out;
            }
        }
        c.cl.sync();
    }
    xarr[td.last] = ct_smooth<true>(c, td.last, 0, cw.n, cw.w);
    for (int l = td.last - 1; l >= td.first; --l) {
        const LevelDev &F = levels[l], &Cc = levels[l + 1];
        const int rp = td.rp[l], i0 = c.rank * rp, i1 = min(i0 + rp, F.ny);
        for (int i = i0 + warp; i < i1; i += CT_THREADS / 32) {
            const bool mi = is_mid(i, F.ny);
            const int I = coarse_lo(i, F.ny, Cc.ny);
            const double *c0 = ct_row(c, l + 1, xarr[l + 1], I), *c1 = mi ? ct_row(c, l + 1, xarr[l + 1], I + 1) : c0;
            double *xf = ct_own(c, l, xarr[l], i);
            for (int j = lane; j < F.nx; j += 32) {
                if (is_dirichlet(F, i, j)) continue;
                const bool mj = is_mid(j, F.nx);
                const int J = coarse_lo(j, F.nx, Cc.nx) + 1;
                double add;
                if (!mi && !mj) add = c0[J];
                else if (!mi && mj) add = 0.5 * (c0[J] + c0[J + 1]);
                else if (mi && !mj) add = 0.5 * (c0[J] + c1[J]);
                else add = 0.5 * (c0[J] + c1[J + 1]);
                xf[j + 1] += add;
            }
        }
        c.cl.sync();
        xarr[l] = ct_smooth<false>(c, l, xarr[l], nu, sw.w);
    }
    {
        const LevelDev &L = levels[td.first];
        const int rp = td.rp[td.first], i0 = c.rank * rp, i1 = min(i0 + rp, L.ny);
        for (int i = i0 + warp; i < i1; i += CT_THREADS / 32) {
            const double *xl = ct_own(c, td.first, xarr[td.first], i);
            for (int j = lane; j < L.nx; j += 32) x_out[(size_t)i * L.nx + j] = xl[j + 1];
        }
    }
    c.cl.sync();  // nobody may exit while a neighbour can still read its shared memory
}
