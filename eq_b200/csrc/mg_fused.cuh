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

#define TS 64                 // shared tile edge (nodes), halo included
#define TP (TS + 2)           // padded row stride
#define TN (TP * TP)          // doubles per shared array
#define TAIL_THREADS 1024
#define MAX_LEVELS 16
#define MAX_CHEB 16

struct SmoothW { double w[4]; };          // per-sweep Jacobi weights (Chebyshev roots)
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

// padded shared index of tile node (ly, lx), 0 <= ly,lx < TS
__device__ __forceinline__ int tidx(int ly, int lx) { return (ly + 1) * TP + lx + 1; }

// does the whole shared tile [ox, ox+TS) x [oy, oy+TS) consist of regular nodes?
__device__ __forceinline__ bool tile_regular(const LevelDev &L, int ox, int oy)
{
    return ox >= 1 && ox + TS - 1 <= L.jreg_hi && oy >= 1 && oy + TS - 1 <= L.ireg_hi && oy >= L.slo &&
           oy + TS <= L.shi;
}

// Per-thread slice of a global field: the thread's column, its R rows.  REG
// tiles lie strictly inside the grid (no clamping, no masking).
template <int R, bool REG>
__device__ __forceinline__ void column_load(const LevelDev &L, int ox, int oy, const double *__restrict__ g,
                                            double (&v)[R], double *mirror = nullptr)
{
    const int lx = threadIdx.x & (TS - 1), ly0 = (threadIdx.x >> 6) * R;
    const int gj = ox + lx;
    if (REG) {
        const double *q = g + (size_t)(oy + ly0) * L.nx + gj;
#pragma unroll
        for (int k = 0; k < R; ++k) v[k] = __ldg(q + (size_t)k * L.nx);
    } else {
        const bool okx = gj >= 0 && gj < L.nx;
        const int gjc = min(max(gj, 0), L.nx - 1);
#pragma unroll
        for (int k = 0; k < R; ++k) v[k] = __ldg(g + (size_t)min(max(oy + ly0 + k, L.slo), L.shi - 1) * L.nx + gjc);
#pragma unroll
        for (int k = 0; k < R; ++k) {
            const int gi = oy + ly0 + k;
            if (!(okx && gi >= 0 && gi < L.ny)) v[k] = 0.0;
            if (mirror) mirror[tidx(ly0 + k, lx)] = v[k];
        }
    }
}

// Shared copy of the padded cell sizes seen by a tile: entries for nodes
// [ox, ox+TS] in x and [oy, oy+TS] in y (clamped to the grid; entries outside
// the grid are never used because those nodes are skipped).  buf: 4*(TS+2).
__device__ __forceinline__ Spacing tile_spacing(const LevelDev &L, int ox, int oy, double *buf)
{
    constexpr int n = TS + 2;
    for (int k = threadIdx.x; k < n; k += blockDim.x) {
        const int j = min(max(ox + k, 0), L.nx), i = min(max(oy + k, 0), L.ny);
        buf[k] = L.hx[j]; buf[n + k] = L.ihx[j]; buf[2 * n + k] = L.hy[i]; buf[3 * n + k] = L.ihy[i];
    }
    Spacing S; S.hx = buf; S.ihx = buf + n; S.hy = buf + 2 * n; S.ihy = buf + 3 * n; S.jo = ox; S.io = oy;
    return S;
}

__device__ __forceinline__ void tile_zero_pads(double *s)
{
    for (int k = threadIdx.x; k < TP; k += blockDim.x) {
        s[k] = 0.0; s[(TP - 1) * TP + k] = 0.0; s[k * TP] = 0.0; s[k * TP + TP - 1] = 0.0;
    }
}

// Generic (any node) update used by the boundary fix-up.
template <int MODE>
__device__ __forceinline__ double node_generic(const LevelDev &L, const Spacing &S, int gi, int gj, int c,
                                               double bval, const double *src, double w)
{
    if (gi < 0 || gi >= L.ny || gj < 0 || gj >= L.nx || is_dirichlet(L, gi, gj)) return 0.0;
    double cf[NBAND];
    stencil_iso(L, S, gi, gj, cf);
    const double id = inv_diag(L, gi, gj, cf);
    if (MODE == 0) return w * id * bval;
    const double ax = cf[B_C] * src[c] + cf[B_E] * src[c + 1] + cf[B_W] * src[c - 1] + cf[B_N] * src[c + TP] +
                      cf[B_S] * src[c - TP] + cf[B_NE] * src[c + TP + 1] + cf[B_SW] * src[c - TP - 1];
    const double res = bval - ax;
    return MODE == 2 ? res : src[c] + w * id * res;
}

// One pass over the whole shared tile; bv[] holds the thread's b values.
// MODE 0: dst = w*D^-1 b ; MODE 1: dst = src + w*D^-1 (b - A src) ; MODE 2: dst = b - A src
// REG tiles: branch-free constant-coefficient update, the 3x3 neighbourhood is
// carried up the column in registers (3 LDS + 1 STS per node).  Other tiles run
// the same update with nodes outside the grid masked to zero, then recompute
// the few irregular nodes (boundary rows/columns, Dirichlet nodes, the narrower
// last cell of a coarse grid) with the general row.
// Values in the outer rings become stale sweep by sweep; callers only consume
// nodes at least (number of MODE 1/2 passes) inside the tile.
template <int MODE, int R, bool REG>
__device__ __forceinline__ void tile_pass(const LevelDev &L, const Spacing &S, int ox, int oy,
                                          const double (&bv)[R], const double *__restrict__ bglob,
                                          const double *src, double *dst, double w, const double *sbm = nullptr)
{
    constexpr int NT = TS * (TS / R);
    const int lx = threadIdx.x & (TS - 1), ly0 = (threadIdx.x >> 6) * R;
    const double cC = L.cC, cEW = L.cEW, cNS = L.cNS, cD = L.cD, wd = w * L.icC;
    int c = tidx(ly0, lx);
    if (REG) {
        if (MODE == 0) {
#pragma unroll
            for (int k = 0; k < R; ++k) dst[c + k * TP] = wd * bv[k];
        } else {
            double sw = src[c - TP - 1], sc = src[c - TP];
            double ww = src[c - 1], cc = src[c], ee = src[c + 1];
#pragma unroll
            for (int k = 0; k < R; ++k, c += TP) {
                const double nw = src[c + TP - 1], nc = src[c + TP], ne = src[c + TP + 1];
                const double ax = cC * cc + cEW * (ee + ww) + cNS * (nc + sc) + cD * (ne + sw);
                const double res = bv[k] - ax;
                dst[c] = MODE == 2 ? res : cc + wd * res;
                sw = ww; sc = cc; ww = nw; cc = nc; ee = ne;
            }
        }
    } else {
        const int gj = ox + lx;
        const bool okx = gj >= 0 && gj < L.nx;
        if (MODE == 0) {
#pragma unroll
            for (int k = 0; k < R; ++k) {
                const bool ok = okx && (unsigned)(oy + ly0 + k) < (unsigned)L.ny;
                dst[c + k * TP] = ok ? wd * bv[k] : 0.0;
            }
        } else {
            double sw = src[c - TP - 1], sc = src[c - TP];
            double ww = src[c - 1], cc = src[c], ee = src[c + 1];
#pragma unroll
            for (int k = 0; k < R; ++k, c += TP) {
                const double nw = src[c + TP - 1], nc = src[c + TP], ne = src[c + TP + 1];
                const double ax = cC * cc + cEW * (ee + ww) + cNS * (nc + sc) + cD * (ne + sw);
                const double res = bv[k] - ax;
                const bool ok = okx && (unsigned)(oy + ly0 + k) < (unsigned)L.ny;
                dst[c] = ok ? (MODE == 2 ? res : cc + wd * res) : 0.0;
                sw = ww; sc = cc; ww = nw; cc = nc; ee = ne;
            }
        }
        __syncthreads();
        // irregular in-grid rows: 0 and ireg_hi+1 .. ny-1 ; columns likewise
        const int nr = L.ny - L.ireg_hi, ncol = L.nx - L.jreg_hi;
        const int items = (nr + ncol) * TS;
        for (int it = threadIdx.x; it < items; it += NT) {
            const int k = it >> 6, e = it & (TS - 1);
            int ly, lxx;
            if (k < nr) { const int gi = k == 0 ? 0 : L.ireg_hi + k; ly = gi - oy; lxx = e; }
            else { const int kk = k - nr; const int gjj = kk == 0 ? 0 : L.jreg_hi + kk; lxx = gjj - ox; ly = e; }
            if (ly < 0 || ly >= TS || lxx < 0 || lxx >= TS) continue;
            const int gi = oy + ly, gjj = ox + lxx;
            const bool in = gi >= L.slo && gi < L.shi && gjj >= 0 && gjj < L.nx;
            const int cc2 = tidx(ly, lxx);
            const double bval = sbm ? sbm[cc2] : (in ? __ldg(bglob + (size_t)gi * L.nx + gjj) : 0.0);
            dst[cc2] = node_generic<MODE>(L, S, gi, gjj, cc2, bval, src, w);
        }
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------
// pre-smoothing: owned region T = TS - 2*(NU+1) nodes per side
// ---------------------------------------------------------------------------
template <int NU, int R, bool REG>
__device__ __forceinline__ void presmooth_body(const LevelDev &F, const LevelDev &Cc, const Spacing &S, int ox,
                                               int oy, const double *__restrict__ b, double *__restrict__ x,
                                               double *__restrict__ bc, const SmoothW &sw, double *xa, double *xb,
                                               double *sb)
{
    constexpr int H = NU + 1, TO = TS - 2 * H, NT = TS * (TS / R);
    const int lx = threadIdx.x & (TS - 1), ly0 = (threadIdx.x >> 6) * R;
    double bv[R];
    column_load<R, REG>(F, ox, oy, b, bv, sb);
    if (sb) __syncthreads();
    tile_pass<0, R, REG>(F, S, ox, oy, bv, b, xa, xa, sw.w[0], sb);
    double *cur = xa, *oth = xb;
#pragma unroll
    for (int k = 1; k < NU; ++k) {
        tile_pass<1, R, REG>(F, S, ox, oy, bv, b, cur, oth, sw.w[k], sb);
        double *t = cur; cur = oth; oth = t;
    }
    tile_pass<2, R, REG>(F, S, ox, oy, bv, b, cur, oth, 0.0, sb);  // residual, valid on T+1
    const int gj = ox + lx;
    if (REG) {
        // x on T -> global
        if (lx >= H && lx < TS - H) {
            double *q = x + (size_t)(oy + ly0) * F.nx + gj;
#pragma unroll
            for (int k = 0; k < R; ++k) {
                const int ly = ly0 + k, gi = oy + ly;
                if (ly >= H && ly < TS - H && gi >= F.wlo && gi < F.whi) q[(size_t)k * F.nx] = cur[tidx(ly, lx)];
            }
        }
        // restriction: coarse nodes = even fine nodes of T (tile origin is even)
        constexpr int CT = TO / 2;
        const int I0 = (oy + H) >> 1, J0 = (ox + H) >> 1;
        for (int q = threadIdx.x; q < CT * CT; q += NT) {
            const int cy = q / CT, cx = q - cy * CT;
            const int gfi = oy + H + 2 * cy;
            if (gfi < F.wlo || gfi >= F.whi) continue;
            const int c = tidx(H + 2 * cy, H + 2 * cx);
            const double h = oth[c + 1] + oth[c - 1] + oth[c + TP] + oth[c - TP] + oth[c + TP + 1] + oth[c - TP - 1];
            bc[(size_t)(I0 + cy) * Cc.nx + J0 + cx] = oth[c] + 0.5 * h;
        }
    } else {
        if (lx >= H && lx < TS - H && gj < F.nx) {
            const bool cj = !(gj & 1) || gj == F.nx - 1;
#pragma unroll 2
            for (int k = 0; k < R; ++k) {
                const int ly = ly0 + k, gi = oy + ly;
                if (ly < H || ly >= TS - H || gi < F.wlo || gi >= F.whi) continue;
                const int c = tidx(ly, lx);
                x[(size_t)gi * F.nx + gj] = cur[c];
                const bool ci = !(gi & 1) || gi == F.ny - 1;
                if (ci && cj) {
                    const int I = coarse_lo(gi, F.ny, Cc.ny), J = coarse_lo(gj, F.nx, Cc.nx);
                    double out = 0.0;
                    if (!is_dirichlet(Cc, I, J)) {
                        const bool e = gj + 1 < F.nx && is_mid(gj + 1, F.nx), w = gj >= 1 && is_mid(gj - 1, F.nx);
                        const bool n = gi + 1 < F.ny && is_mid(gi + 1, F.ny), s = gi >= 1 && is_mid(gi - 1, F.ny);
                        double h = 0.0;
                        if (e) h += oth[c + 1];
                        if (w) h += oth[c - 1];
                        if (n) h += oth[c + TP];
                        if (s) h += oth[c - TP];
                        if (n && e) h += oth[c + TP + 1];
                        if (s && w) h += oth[c - TP - 1];
                        out = oth[c] + 0.5 * h;
                    }
                    bc[(size_t)I * Cc.nx + J] = out;
                }
            }
        }
    }
}

template <int NU, int R>
__global__ void __launch_bounds__(TS *(TS / R), R == 8 ? 2 : 1)
k_presmooth(LevelDev F, LevelDev Cc, const double *__restrict__ b, double *__restrict__ x,
            double *__restrict__ bc, SmoothW sw, const CGScalars *sc)
{
    if (sc->done) return;
    constexpr int H = NU + 1, TO = TS - 2 * H;
    extern __shared__ double sm[];
    double *xa = sm, *xb = sm + TN;
    __shared__ double spc[4 * (TS + 2)];
    const int ox = blockIdx.x * TO - H, oy = F.tbase + blockIdx.y * TO - H;
    const bool regular = tile_regular(F, ox, oy);
    tile_zero_pads(xa);
    tile_zero_pads(xb);
    if (regular) {
        __syncthreads();
        presmooth_body<NU, R, true>(F, Cc, Spacing{}, ox, oy, b, x, bc, sw, xa, xb, nullptr);
    } else {
        const Spacing S = tile_spacing(F, ox, oy, spc);
        __syncthreads();
        presmooth_body<NU, R, false>(F, Cc, S, ox, oy, b, x, bc, sw, xa, xb, R == 4 ? sm + 2 * TN : nullptr);
    }
}

// ---------------------------------------------------------------------------
// coarsest level: all NC Chebyshev-Jacobi sweeps of the coarse solve on one tile
// pass (halo NC-1), x = p_NC(D^-1 A) b from a zero guess.  Replaces a chain of
// NC latency-bound whole-level sweeps by a single wave of a few CTAs.
// ---------------------------------------------------------------------------
template <int NC, int R, bool REG>
__device__ __forceinline__ void coarsest_body(const LevelDev &F, const Spacing &S, int ox, int oy,
                                              const double *__restrict__ b, double *__restrict__ x,
                                              const CoarseW &cw, double *xa, double *xb, double *sb)
{
    constexpr int H = NC - 1;
    const int lx = threadIdx.x & (TS - 1), ly0 = (threadIdx.x >> 6) * R;
    double bv[R];
    column_load<R, REG>(F, ox, oy, b, bv, sb);
    if (sb) __syncthreads();
    tile_pass<0, R, REG>(F, S, ox, oy, bv, b, xa, xa, cw.w[0], sb);
    double *cur = xa, *oth = xb;
#pragma unroll 1
    for (int k = 1; k < NC; ++k) {
        tile_pass<1, R, REG>(F, S, ox, oy, bv, b, cur, oth, cw.w[k], sb);
        double *t = cur; cur = oth; oth = t;
    }
    const int gj = ox + lx;
    if (lx >= H && lx < TS - H && gj >= 0 && gj < F.nx) {
#pragma unroll
        for (int k = 0; k < R; ++k) {
            const int ly = ly0 + k, gi = oy + ly;
            if (ly >= H && ly < TS - H && gi >= F.wlo && gi < F.whi) x[(size_t)gi * F.nx + gj] = cur[tidx(ly, lx)];
        }
    }
}

template <int NC, int R>
__global__ void __launch_bounds__(TS *(TS / R), 1)
k_coarsest(LevelDev F, const double *__restrict__ b, double *__restrict__ x, CoarseW cw, const CGScalars *sc)
{
    if (sc->done) return;
    constexpr int H = NC - 1, TO = TS - 2 * H;
    extern __shared__ double sm[];
    double *xa = sm, *xb = sm + TN;
    __shared__ double spc[4 * (TS + 2)];
    const int ox = blockIdx.x * TO - H, oy = F.tbase + blockIdx.y * TO - H;
    const bool regular = tile_regular(F, ox, oy);
    tile_zero_pads(xa);
    tile_zero_pads(xb);
    if (regular) {
        __syncthreads();
        coarsest_body<NC, R, true>(F, Spacing{}, ox, oy, b, x, cw, xa, xb, nullptr);
    } else {
        const Spacing S = tile_spacing(F, ox, oy, spc);
        __syncthreads();
        coarsest_body<NC, R, false>(F, S, ox, oy, b, x, cw, xa, xb, R == 4 ? sm + 2 * TN : nullptr);
    }
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

// ---------------------------------------------------------------------------
// post-smoothing: owned region T = TS - 2*NU nodes per side
// ---------------------------------------------------------------------------
template <int NU, bool DOT, int R, bool REG>
__device__ __forceinline__ double postsmooth_body(const LevelDev &F, const LevelDev &Cc, const Spacing &S, int ox,
                                                  int oy, const double *__restrict__ b,
                                                  const double *__restrict__ xin, double *__restrict__ x,
                                                  const double *__restrict__ xc, const SmoothW &sw, double *xa,
                                                  double *xb, double *sb)
{
    constexpr int H = NU, CP = TS / 2 + 2;
    const int lx = threadIdx.x & (TS - 1), ly0 = (threadIdx.x >> 6) * R;
    const int gj = ox + lx;
    double bv[R], xv[R];
    column_load<R, REG>(F, ox, oy, b, bv, sb);
    column_load<R, REG>(F, ox, oy, xin, xv);
    // coarse patch covering the tile -> xb (as scratch): coarse nodes [J0, J0+CP) x [I0, I0+CP)
    const int J0 = coarse_lo(min(max(ox, 0), F.nx - 1), F.nx, Cc.nx);
    const int I0 = coarse_lo(min(max(oy, 0), F.ny - 1), F.ny, Cc.ny);
    for (int k = threadIdx.x; k < CP * CP; k += blockDim.x) {
        const int ci = k / CP, cj = k - ci * CP;
        xb[k] = __ldg(xc + (size_t)min(max(I0 + ci, Cc.slo), Cc.shi - 1) * Cc.nx + min(J0 + cj, Cc.nx - 1));
    }
    __syncthreads();
    if (REG) {   // interior tile: midpoint <=> odd index, nothing is Dirichlet
        const bool mj = gj & 1;
        const int cj = (gj >> 1) - J0;
#pragma unroll
        for (int k = 0; k < R; ++k) {
            const int gi = oy + ly0 + k;
            const double *c0 = xb + ((gi >> 1) - I0) * CP + cj;
            const bool mi = gi & 1;
            const double a0 = c0[0], a1 = c0[mi ? (mj ? CP + 1 : CP) : (mj ? 1 : 0)];
            xa[tidx(ly0 + k, lx)] = xv[k] + 0.5 * (a0 + a1);
        }
    } else {
        const bool okx = gj >= 0 && gj < F.nx;
        const bool mj = okx && is_mid(gj, F.nx);
        const int cj = okx ? coarse_lo(gj, F.nx, Cc.nx) - J0 : 0;
#pragma unroll 2
        for (int k = 0; k < R; ++k) {
            const int gi = oy + ly0 + k;
            double v = 0.0;
            if (okx && gi >= 0 && gi < F.ny && !is_dirichlet(F, gi, gj)) {
                const bool mi = is_mid(gi, F.ny);
                const double *c0 = xb + (coarse_lo(gi, F.ny, Cc.ny) - I0) * CP + cj;
                double add;
                if (!mi && !mj) add = c0[0];
                else if (!mi && mj) add = 0.5 * (c0[0] + c0[1]);
                else if (mi && !mj) add = 0.5 * (c0[0] + c0[CP]);
                else add = 0.5 * (c0[0] + c0[CP + 1]);
                v = xv[k] + add;
            }
            xa[tidx(ly0 + k, lx)] = v;
        }
    }
    __syncthreads();
    tile_zero_pads(xb);  // the patch scratch overlapped xb's pad ring
    __syncthreads();
    double *cur = xa, *oth = xb;
#pragma unroll
    for (int k = 0; k < NU; ++k) {
        tile_pass<1, R, REG>(F, S, ox, oy, bv, b, cur, oth, sw.w[k], sb);
        double *t = cur; cur = oth; oth = t;
    }
    double acc = 0.0;
    if (lx >= H && lx < TS - H && (REG || gj < F.nx)) {
#pragma unroll
        for (int k = 0; k < R; ++k) {
            const int ly = ly0 + k, gi = oy + ly;
            if (ly >= H && ly < TS - H && gi >= F.wlo && gi < F.whi) {
                const double v = cur[tidx(ly, lx)];
                x[(size_t)gi * F.nx + gj] = v;
                if (DOT) acc += v * bv[k];
            }
        }
    }
    return acc;
}

template <int NU, bool DOT, int R>
__global__ void __launch_bounds__(TS *(TS / R), R == 8 ? 2 : 1)
k_postsmooth(LevelDev F, LevelDev Cc, const double *__restrict__ b, const double *__restrict__ xin,
             double *__restrict__ x, const double *__restrict__ xc, SmoothW sw, CGScalars *sc,
             double *partials, unsigned *counter, double *out_dot)
{
    if (sc->done) return;
    constexpr int H = NU, TO = TS - 2 * H;
    extern __shared__ double sm[];
    double *xa = sm, *xb = sm + TN;
    __shared__ double spc[4 * (TS + 2)];
    const int ox = blockIdx.x * TO - H, oy = F.tbase + blockIdx.y * TO - H;
    const bool regular = tile_regular(F, ox, oy);
    tile_zero_pads(xa);
    double v[1];
    if (regular) {
        v[0] = postsmooth_body<NU, DOT, R, true>(F, Cc, Spacing{}, ox, oy, b, xin, x, xc, sw, xa, xb, nullptr);
    } else {
        const Spacing S = tile_spacing(F, ox, oy, spc);
        v[0] = postsmooth_body<NU, DOT, R, false>(F, Cc, S, ox, oy, b, xin, x, xc, sw, xa, xb,
                                                   R == 4 ? sm + 2 * TN : nullptr);
    }
    if (DOT) {
        double tot[1];
        if (grid_reduce<1>(v, partials, counter, tot)) *out_dot = tot[0];
    }
}

// ---------------------------------------------------------------------------
// p' = z + beta p ; Ap = A p' ; p'.Ap      (1-node halo; owned TS-2 per side)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_apply_p(LevelDev L, const double *__restrict__ z, const double *__restrict__ pin, double *__restrict__ p,
          double *__restrict__ Ap, CGScalars *sc, double *partials, unsigned *counter, double *out_pAp)
{
    if (sc->done) return;
    constexpr int TO = TS - 2, TROWS = 16;
    __shared__ double sp[TN];
    const double beta = sc->iters == 0 ? 0.0 : sc->rz_new / sc->rz_old;
    const int ox = blockIdx.x * TO - 1, oy = L.tbase + blockIdx.y * TO - 1;
    const int lx = threadIdx.x & (TS - 1), ly0 = (threadIdx.x >> 6) * TROWS;
    const int gj = ox + lx;
    const bool regular = tile_regular(L, ox, oy);
    __shared__ double spc[4 * (TS + 2)];
    const Spacing S = tile_spacing(L, ox, oy, spc);
    tile_zero_pads(sp);
    {
        const bool okx = gj >= 0 && gj < L.nx;
        const int gjc = min(max(gj, 0), L.nx - 1);
        double vz[TROWS], vp[TROWS];
#pragma unroll
        for (int k = 0; k < TROWS; ++k) {
            const size_t g = (size_t)min(max(oy + ly0 + k, L.slo), L.shi - 1) * L.nx + gjc;
            vz[k] = __ldg(z + g);
            vp[k] = __ldg(pin + g);
        }
#pragma unroll
        for (int k = 0; k < TROWS; ++k) {
            const int gi = oy + ly0 + k;
            sp[tidx(ly0 + k, lx)] = (okx && gi >= 0 && gi < L.ny) ? vz[k] + beta * vp[k] : 0.0;
        }
    }
    __syncthreads();
    double v[1] = {0.0};
    if (lx >= 1 && lx < TS - 1 && gj < L.nx) {
        const double cC = L.cC, cEW = L.cEW, cNS = L.cNS, cD = L.cD;
        const bool regx = regular || (gj >= 1 && gj <= L.jreg_hi);
        int c = tidx(ly0, lx);
        double sw = sp[c - TP - 1], s0 = sp[c - TP];
        double ww = sp[c - 1], cc = sp[c], ee = sp[c + 1];
#pragma unroll
        for (int k = 0; k < TROWS; ++k, c += TP) {
            const double nw = sp[c + TP - 1], nc = sp[c + TP], ne = sp[c + TP + 1];
            const int ly = ly0 + k, gi = oy + ly;
            if (ly >= 1 && ly < TS - 1 && gi >= L.wlo && gi < L.whi) {
                const size_t g = (size_t)gi * L.nx + gj;
                p[g] = cc;
                if (regx && (regular || (gi >= 1 && gi <= L.ireg_hi))) {
                    const double ax = cC * cc + cEW * (ee + ww) + cNS * (nc + s0) + cD * (ne + sw);
                    Ap[g] = ax;
                    v[0] += ax * cc;
                }
            }
            sw = ww; s0 = cc; ww = nw; cc = nc; ee = ne;
        }
    }
    if (!regular) {
        // irregular owned nodes: rows 0, ireg_hi+1..ny-1 and columns 0, jreg_hi+1..nx-1
        const int nr = L.ny - L.ireg_hi, ncol = L.nx - L.jreg_hi;
        const int items = (nr + ncol) * TS;
        for (int it = threadIdx.x; it < items; it += 256) {
            const int k = it >> 6, e = it & (TS - 1);
            int ly, lxx;
            bool rowitem = k < nr;
            if (rowitem) { const int gi = k == 0 ? 0 : L.ireg_hi + k; ly = gi - oy; lxx = e; }
            else { const int kk = k - nr; const int gjj = kk == 0 ? 0 : L.jreg_hi + kk; lxx = gjj - ox; ly = e; }
            if (ly < 1 || ly >= TS - 1 || lxx < 1 || lxx >= TS - 1) continue;
            const int gi = oy + ly, gjj = ox + lxx;
            if (gi < L.wlo || gi >= L.whi || gjj >= L.nx) continue;
            // a node on an irregular row AND an irregular column is handled by its row item only
            if (!rowitem && !(gi >= 1 && gi <= L.ireg_hi)) continue;
            const int c = tidx(ly, lxx);
            const double pc = sp[c];
            double out = 0.0;
            if (!is_dirichlet(L, gi, gjj)) {
                double cf[NBAND];
                stencil_iso(L, S, gi, gjj, cf);
                out = cf[B_C] * pc + cf[B_E] * sp[c + 1] + cf[B_W] * sp[c - 1] + cf[B_N] * sp[c + TP] +
                      cf[B_S] * sp[c - TP] + cf[B_NE] * sp[c + TP + 1] + cf[B_SW] * sp[c - TP - 1];
                v[0] += out * pc;
            }
            Ap[(size_t)gi * L.nx + gjj] = out;
        }
    }
    double tot[1];
    if (grid_reduce<1>(v, partials, counter, tot)) *out_pAp = tot[0];
}

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
