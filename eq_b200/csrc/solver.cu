// Matrix-free P1-FEM operator, load vector, and the multigrid-preconditioned CG
// that replaces LinearVariationalSolver::solve() (src/fHSL.cpp:104-108).
//
// System solved each step (fenics/hslD.ufl:36-42):
//   (M + dt*K(D) + dt*R) u = M u0 + dt*r*s*e,  Dirichlet rows u_i = g_i.
// DOLFIN assembles it and runs a sparse LU; here A is never formed: every
// kernel evaluates the 7-point row of the "right"-diagonal mesh in registers.
#include "eqgpu_internal.cuh"
#include "mg_fused.cuh"
#include "mg_cluster.cuh"
#include "mg_stream.cuh"
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>

#define BX 32
#define BY 8
constexpr int TN64 = 66 * 66, TN32 = 34 * 34;  // doubles per shared tile array (mg_tile.inc TN)
// dynamic shared memory of the tile kernels: two or three tile arrays + the irregular-node table (mg_tile.inc TAB_DOUBLES)
constexpr size_t SM64_2 = (2 * TN64 + 6 * 64 * 8) * sizeof(double), SM64_3 = (3 * TN64 + 6 * 64 * 8) * sizeof(double);
constexpr size_t SM32_3 = (3 * TN32 + 6 * 32 * 8) * sizeof(double);
// Can the deep-halo tile kernel do an n-sweep coarsest solve?  (64-node tile, halo n-1, owned region >= 8;
// row slabs carry 6 halo rows.)
static inline bool coarsest_tileable(const eqgpu_solver *s, int n)
{
    return n >= 2 && n <= MAX_CHEB && 64 - 2 * (n - 1) >= 8 && (!s->slab || n <= 7);
}

struct DirData {
    double val[4];
    const double *top, *bot;  // per-node channel values or null
};

__device__ __forceinline__ double dir_value(const LevelDev &L, const DirData &d, int i, int j)
{
    // DirichletBC objects are applied in the order left,right,top,bottom
    // (src/fHSL.cpp:468-539); the last one applied wins at a corner.
    const unsigned m = L.dirmask;
    if (L.lumped) {   // ApplyBoundaryConditions writes top/bottom first, then right/left (diffuclass.cpp:218-272)
        if ((m & 2u) && j == L.nx - 1) return d.val[1];
        if ((m & 1u) && j == 0) return d.val[0];
    }
    if ((m & 8u) && i == 0) return d.bot ? d.bot[j] : d.val[3];
    if ((m & 4u) && i == L.ny - 1) return d.top ? d.top[j] : d.val[2];
    if ((m & 2u) && j == L.nx - 1) return d.val[1];
    return d.val[0];
}

// ---------------------------------------------------------------------------
// operator apply: y = A x on free rows (Dirichlet rows -> 0), optional x.y
// ---------------------------------------------------------------------------
template <int TENSOR, bool DOT>
__global__ void __launch_bounds__(BX *BY)
k_apply(LevelDev L, const double *__restrict__ x, double *__restrict__ y, CGScalars *sc,
        double *partials, unsigned *counter, double *out_pAp)
{
    if (DOT && sc->done) return;
    const int j = blockIdx.x * BX + threadIdx.x, i = blockIdx.y * BY + threadIdx.y;
    double v[1] = {0.0};
    if (i >= L.own0 && i < L.own1 && j < L.nx) {
        const size_t g = (size_t)i * L.nx + j;
        double out = 0.0;
        if (!is_dirichlet(L, i, j)) {
            double c[NBAND];
            stencil_row<TENSOR>(L, i, j, c);
            out = stencil_dot(L, i, j, c, x);
            if (DOT) v[0] = out * __ldg(x + g);
        }
        y[g] = out;
    }
    if (DOT) {
        double tot[1];
        if (grid_reduce<1>(v, partials, counter, tot)) *out_pAp = tot[0];
    }
}

// Variable tensor with assembled rows, one GPU: p' = z + beta p and Ap = A p' in ONE pass (the unfused pair k_update_p +
// k_apply<2> reads and writes p twice).  Kept opt-in: measured slower than the pair (see the call site).  p' is needed at the six neighbours too, so each thread forms it there from z and
// the OLD p -- which is why the new direction goes to a second buffer (the p ping-pong of the isotropic path); the
// neighbours' loads hit L1/L2 (a 32 x 8 block re-reads a one-node rim).  Dirichlet rows: p' = 0 (r and z vanish there).
__global__ void __launch_bounds__(BX *BY)
k_apply_p_asm(LevelDev L, const double *__restrict__ z, const double *__restrict__ pin, double *__restrict__ p,
              double *__restrict__ Ap, CGScalars *sc, double *partials, unsigned *counter, double *out_pAp)
{
    if (sc->done) return;
    const double beta = sc->iters == 0 ? 0.0 : sc->rz_new / sc->rz_old;
    const int j = blockIdx.x * BX + threadIdx.x, i = blockIdx.y * BY + threadIdx.y;
    double v[1] = {0.0};
    if (i >= L.own0 && i < L.own1 && j < L.nx) {
        const size_t g = (size_t)i * L.nx + j, nx = (size_t)L.nx;
        auto pn = [&](size_t q) { return __ldg(z + q) + beta * __ldg(pin + q); };
        const double pc = pn(g);
        double out = 0.0;
        if (!is_dirichlet(L, i, j)) {
            double c[NBAND];
            stencil_row<2>(L, i, j, c);
            const bool hasW = j > 0, hasE = j < L.nx - 1, hasS = i > 0, hasN = i < L.ny - 1;
            out = c[B_C] * pc;
            if (hasE) out += c[B_E] * pn(g + 1);
            if (hasW) out += c[B_W] * pn(g - 1);
            if (hasN) out += c[B_N] * pn(g + nx);
            if (hasS) out += c[B_S] * pn(g - nx);
            if (hasN && hasE) out += c[B_NE] * pn(g + nx + 1);
            if (hasS && hasW) out += c[B_SW] * pn(g - nx - 1);
            v[0] = out * pc;
        }
        p[g] = pc;
        Ap[g] = out;
    }
    double tot[1];
    if (grid_reduce<1>(v, partials, counter, tot)) *out_pAp = tot[0];
}

// Verification hook: y = A x with explicit handling of constrained rows/cols.
// mode 0: unconstrained operator.  mode 1: Dirichlet rows = identity, columns
// to Dirichlet nodes dropped (the symmetric elimination CG works with).
template <bool TENSOR>
__global__ void k_apply_check(LevelDev L, const double *__restrict__ x, double *__restrict__ y, int mode)
{
    const int j = blockIdx.x * BX + threadIdx.x, i = blockIdx.y * BY + threadIdx.y;
    if (i >= L.ny || j >= L.nx) return;
    const size_t g = (size_t)i * L.nx + j;
    if (mode == 1 && is_dirichlet(L, i, j)) { y[g] = x[g]; return; }
    double c[NBAND];
    stencil_row<TENSOR>(L, i, j, c);
    if (mode == 1) {
        if (j + 1 < L.nx && is_dirichlet(L, i, j + 1)) c[B_E] = 0.0;
        if (j > 0 && is_dirichlet(L, i, j - 1)) c[B_W] = 0.0;
        if (i + 1 < L.ny && is_dirichlet(L, i + 1, j)) c[B_N] = 0.0;
        if (i > 0 && is_dirichlet(L, i - 1, j)) c[B_S] = 0.0;
        if (i + 1 < L.ny && j + 1 < L.nx && is_dirichlet(L, i + 1, j + 1)) c[B_NE] = 0.0;
        if (i > 0 && j > 0 && is_dirichlet(L, i - 1, j - 1)) c[B_SW] = 0.0;
    }
    y[g] = stencil_dot(L, i, j, c, x);
}

// b = M u0 + Robin load (unconstrained load vector L(phi_g), fenics/hslD.h:3466-3689)
__device__ __forceinline__ double load_row(const LevelDev &L, int i, int j,
                                           const double *__restrict__ u0, double rs_l, double rs_r)
{
    double c[NBAND];
    if (i >= 1 && i <= L.ireg_hi && j >= 1 && j <= L.jreg_hi) {  // regular node: constant consistent-mass row
        const double m = L.mO;  // hx*hy/12 (0 with the lumped mass)
        c[B_C] = L.mC; c[B_E] = m; c[B_W] = m; c[B_N] = m; c[B_S] = m; c[B_NE] = m; c[B_SW] = m;
    } else stencil_mass(L, i, j, c);
    double b = stencil_dot(L, i, j, c, u0);
    if (j == 0) b += rs_l * 0.5 * (L.hy[i] + L.hy[i + 1]);
    if (j == L.nx - 1) b += rs_r * 0.5 * (L.hy[i] + L.hy[i + 1]);
    return b;
}

__global__ void k_rhs_check(LevelDev L, const double *__restrict__ u0, double *__restrict__ b,
                            double rs_l, double rs_r)
{
    const int j = blockIdx.x * BX + threadIdx.x, i = blockIdx.y * BY + threadIdx.y;
    if (i >= L.ny || j >= L.nx) return;
    b[(size_t)i * L.nx + j] = load_row(L, i, j, u0, rs_l, rs_r);
}

// ---------------------------------------------------------------------------
// PCG start.  Reduced system on the free rows: A_ff x_f = b_f - A_fd g_d with
// b = M u0 + load.  Two starting guesses are evaluated in one pass:
//   (A) x0 = u0 (the previous field plus deposits)  -> rA = b - A x0
//   (B) x0 = 0                                      -> rB = b - A_fd g_d
// (Dirichlet nodes hold g in both.)  rB is also the right-hand side of the
// reduced system, so ||rB|| is the norm the stopping test is relative to.
// k_impose keeps whichever guess has the smaller residual: u0 is excellent in
// quasi-steady state and poor right after large point deposits, where the
// huge A*u0 would cap the attainable accuracy.
// ---------------------------------------------------------------------------
template <int TENSOR>
__global__ void __launch_bounds__(BX *BY)
k_init(LevelDev L, DirData dd, const double *__restrict__ u, double *__restrict__ rA,
       double *__restrict__ rB, double rs_l, double rs_r, CGScalars *sc, double *partials,
       unsigned *counter, double *out_rr0, double *out_b2)
{
    const int j = blockIdx.x * BX + threadIdx.x, i = blockIdx.y * BY + threadIdx.y;
    double v[2] = {0.0, 0.0};
    if (i >= L.own0 && i < L.own1 && j < L.nx) {
        const size_t g = (size_t)i * L.nx + j;
        double resA = 0.0, resB = 0.0;
        if (!is_dirichlet(L, i, j)) {
            const double b = load_row(L, i, j, u, rs_l, rs_r);
            double c[NBAND];
            stencil_row<TENSOR>(L, i, j, c);
            const bool hasW = j > 0, hasE = j < L.nx - 1, hasS = i > 0, hasN = i < L.ny - 1;
            double ax = c[B_C] * __ldg(u + g), ag = 0.0;
            auto acc = [&](int ii, int jj, double ck) {
                if (is_dirichlet(L, ii, jj)) {
                    const double t = ck * dir_value(L, dd, ii, jj);
                    ax += t; ag += t;
                } else ax += ck * __ldg(u + (size_t)ii * L.nx + jj);
            };
            if (hasE) acc(i, j + 1, c[B_E]);
            if (hasW) acc(i, j - 1, c[B_W]);
            if (hasN) acc(i + 1, j, c[B_N]);
            if (hasS) acc(i - 1, j, c[B_S]);
            if (hasN && hasE) acc(i + 1, j + 1, c[B_NE]);
            if (hasS && hasW) acc(i - 1, j - 1, c[B_SW]);
            resA = b - ax;
            resB = b - ag;
            v[0] = resA * resA;
            v[1] = resB * resB;
        }
        rA[g] = resA;
        rB[g] = resB;
    }
    double tot[2];
    if (grid_reduce<2>(v, partials, counter, tot)) {
        *out_rr0 = tot[0];
        *out_b2 = tot[1];
    }
}

// k_init_tile: 64x64 tile, one column per thread, 64*64/INIT_THREADS rows each.  1024 threads (4 rows):
// the four stage-and-walk phases are exposed to load latency, and 32 warps hide it better than 16 with 8 rows
// (122 registers, one CTA per SM either way).
#define INIT_THREADS 1024
constexpr int INIT_SMEM = 5 * 66 * 66 * (int)sizeof(double);   // five 64x64 tiles with a pad ring: 174 KB

// A x at a free node for the per-node (irregular frame) path of k_init_tile; Dirichlet neighbours
// contribute their boundary value (ag collects those terms: A_fd g_d).
__device__ __forceinline__ void frame_apply(const LevelDev &L, const DirData &dd, const double c[NBAND],
                                            const double *__restrict__ x, int i, int j, double &ax, double &ag)
{
    const bool hasW = j > 0, hasE = j < L.nx - 1, hasS = i > 0, hasN = i < L.ny - 1;
    ax = c[B_C] * __ldg(x + (size_t)i * L.nx + j);
    ag = 0.0;
    auto acc = [&](int ii, int jj, double ck) {
        if (is_dirichlet(L, ii, jj)) {
            const double t = ck * dir_value(L, dd, ii, jj);
            ax += t; ag += t;
        } else ax += ck * __ldg(x + (size_t)ii * L.nx + jj);
    };
    if (hasE) acc(i, j + 1, c[B_E]);
    if (hasW) acc(i, j - 1, c[B_W]);
    if (hasN) acc(i + 1, j, c[B_N]);
    if (hasS) acc(i - 1, j, c[B_S]);
    if (hasN && hasE) acc(i + 1, j + 1, c[B_NE]);
    if (hasS && hasW) acc(i - 1, j - 1, c[B_SW]);
}

// Tile version of k_init for the isotropic operator on one GPU, with warm starts.  Every field goes
// through a shared 64x64 tile (1-node halo, 62x62 owned), so that a 7-point row costs one global load per
// node.  Candidate starting guesses, all evaluated in this one pass (k_impose keeps the best):
//   (1) g1 = h0, the previous step's solution (nh >= 1), or u0 itself, the field as given (nh == 0);
//   (B) zero;
//   (D) 2 h0 - h1, the linear extrapolation of the last two solutions (nh >= 2);
//   (E) 3 h0 - 3 h1 + h2, the quadratic extrapolation of the last three (nh == 3).
// The previous solution does NOT contain this step's point deposits, whose huge A*d makes u0 a terrible
// guess (1500x the residual of zero on the bench colony); in quasi-steady state its residual is the change of
// the sources, and the extrapolations remove the linear / quadratic drift as well (measured on the bench
// colony after 30 steps: 2e-2, 7e-4 and 2e-5 of the zero guess's residual).
// Outputs: r1 = b - A g1, rB = b - A_fd g_d, d1 = A h0 - A h1, d2 = A h1 - A h2 (so that the residuals of
// (D) and (E) are r1 - d1 and r1 - 2 d1 + d2), and the four squared norms.
// Tiles that touch the irregular frame (wall rows/columns, their Dirichlet-adjacent neighbours, the
// narrower last cell) take the general per-node path.
// NTH threads (1024: 4 rows per thread, one CTA per SM; 512: 8 rows, two CTAs per SM when the history is short: MAXH <= 1
// keeps the register count at 64) and MAXH, the compile-time bound of nh.
template <int NTH, int MAXH>
__global__ void __launch_bounds__(NTH, NTH == 512 ? 2 : 1)
k_init_tile(LevelDev L, DirData dd, const double *__restrict__ u, const double *__restrict__ h0,
            const double *__restrict__ h1, const double *__restrict__ h2, const double *__restrict__ h3, int nh,
            double *__restrict__ r1, double *__restrict__ rB, double *__restrict__ d1o, double *__restrict__ d2o,
            double *__restrict__ d3o, double rs_l, double rs_r, double *partials, unsigned *counter, CGScalars *sc,
            int slab, const double *dk_prev, double *dk_next, int d4ok)
{
    // Row slabs: L is the rank's local view (halo rows included); only owned rows [own0, own1) are written
    // and summed, and the sums are rank-local partials (slab != 0) for the host to all-reduce.
    constexpr int TSI = 64, TPI = TSI + 2, TOI = TSI - 2, TROWS = TSI * TSI / NTH;
    extern __shared__ double sm_init[];   // five tiles: u0, h0, h1, h2, h3
    const int ox = blockIdx.x * TOI - 1, oy = blockIdx.y * TOI - 1;
    const int lx = threadIdx.x & (TSI - 1), ly0 = (threadIdx.x >> 6) * TROWS;
    const int gj = ox + lx;
    // owned nodes [ox+1, ox+62] x [oy+1, oy+62] all regular and at least two nodes away from every wall
    const bool deep = ox + 1 >= 2 && ox + TOI <= min(L.jreg_hi, L.nx - 3) && oy + 1 >= 2 &&
                      oy + TOI <= min(L.ireg_hi, L.ny - 3);
    // dk_next (nh >= 4): this step's d3 = A (h2 - h3), kept for the next step, whose d4 = A (h3 - h4) it is;
    // dk_prev (d4ok): the one the previous step kept -- the quartic candidate costs no fifth operator walk
    double v[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    if (deep) {
        // all input tiles in flight at once: one cp.async commit group per field (an empty group where the
        // history is shorter), waited for one by one below -- no registers held, no load phase per stage
        auto stage = [&](const double *__restrict__ src, int q, bool on) {
            if (on) {
                const double *g = src + (size_t)(oy + ly0) * L.nx + gj;
                double *t = sm_init + q * (TPI * TPI) + (ly0 + 1) * TPI + lx + 1;
#pragma unroll
                for (int k = 0; k < TROWS; ++k) cp_async8(t + k * TPI, g + (size_t)k * L.nx);
            }
            cp_async_commit();
        };
        stage(u, 0, true);
        stage(h0, 1, nh >= 1);
        stage(h1, 2, MAXH >= 2 && nh >= 2);
        stage(h2, 3, MAXH >= 3 && nh >= 3);
        stage(h3, 4, MAXH >= 4 && nh >= 4);
        // 7-point walk up the thread's column over tile q: mass row or operator row
        auto walk = [&](int q, bool mass, double (&out)[TROWS]) {
            const double *sp = sm_init + q * (TPI * TPI);
            const double cC = mass ? L.mC : L.cC, cEW = mass ? L.mO : L.cEW, cNS = mass ? L.mO : L.cNS;
            const double cD = mass ? L.mO : L.cD;
            int c = (ly0 + 1) * TPI + lx + 1;
            double sw = 0.0, s0 = 0.0;
            if (ly0 > 0) { sw = sp[c - TPI - 1]; s0 = sp[c - TPI]; }
            double ww = sp[c - 1], cc = sp[c], ee = sp[c + 1];
#pragma unroll
            for (int k = 0; k < TROWS; ++k, c += TPI) {
                double nw = 0.0, nc = 0.0, ne = 0.0;
                if (ly0 + k < TSI - 1) { nw = sp[c + TPI - 1]; nc = sp[c + TPI]; ne = sp[c + TPI + 1]; }
                out[k] = cC * cc + cEW * (ee + ww) + cNS * (nc + s0) + cD * (ne + sw);
                sw = ww; s0 = cc; ww = nw; cc = nc; ee = ne;
            }
        };
        const bool col = lx >= 1 && lx < TSI - 1;
        double b[TROWS], c1[TROWS], d1[TROWS], d2[TROWS], d3[TROWS], a[TROWS];
        cp_async_wait<4>();
        __syncthreads();
        if (col) walk(0, true, b);                    // b = M u0
        cp_async_wait<3>();
        __syncthreads();
        if (col) {
            walk(nh >= 1 ? 1 : 0, false, a);          // A g1 (g1 = u0 when there is no history)
#pragma unroll
            for (int k = 0; k < TROWS; ++k) {
                c1[k] = b[k] - a[k]; d1[k] = a[k]; d2[k] = 0.0; d3[k] = 0.0;
                // the right-hand side is complete: store it now, so that b's registers are free for the
                // remaining walks (64 registers per thread at 1024 threads)
                const int ly = ly0 + k;
                if (ly >= 1 && ly < TSI - 1 && oy + ly >= L.own0 && oy + ly < L.own1) {
                    rB[(size_t)(oy + ly) * L.nx + gj] = b[k];
                    v[1] += b[k] * b[k];
                }
            }
        }
        if (MAXH >= 2 && nh >= 2) {
            cp_async_wait<2>();
            __syncthreads();
            if (col) {
                walk(2, false, a);
#pragma unroll
                for (int k = 0; k < TROWS; ++k) { d1[k] -= a[k]; d2[k] = a[k]; }
            }
        }
        if (MAXH >= 3 && nh >= 3) {
            cp_async_wait<1>();
            __syncthreads();
            if (col) {
                walk(3, false, a);
#pragma unroll
                for (int k = 0; k < TROWS; ++k) { d2[k] -= a[k]; d3[k] = a[k]; }
            }
        }
        if (MAXH >= 4 && nh >= 4) {
            cp_async_wait<0>();
            __syncthreads();
            if (col) {
                walk(4, false, a);
#pragma unroll
                for (int k = 0; k < TROWS; ++k) d3[k] -= a[k];   // A h2 - A h3
            }
        }
        if (col) {
#pragma unroll
            for (int k = 0; k < TROWS; ++k) {
                const int ly = ly0 + k;
                if (ly < 1 || ly >= TSI - 1 || oy + ly < L.own0 || oy + ly >= L.own1) continue;
                const size_t g = (size_t)(oy + ly) * L.nx + gj;
                r1[g] = c1[k];
                v[0] += c1[k] * c1[k];
                if (MAXH >= 2 && nh >= 2) {
                    const double rd = c1[k] - d1[k];
                    d1o[g] = d1[k];
                    v[2] += rd * rd;
                }
                if (MAXH >= 3 && nh >= 3) {
                    const double re = c1[k] - 2.0 * d1[k] + d2[k];
                    d2o[g] = d2[k];
                    v[3] += re * re;
                }
                if (MAXH >= 4 && nh >= 4) {   // cubic: b - A (4 h0 - 6 h1 + 4 h2 - h3) = r1 - 3 d1 + 3 d2 - d3
                    const double rf = c1[k] - 3.0 * (d1[k] - d2[k]) - d3[k];
                    d3o[g] = d3[k];
                    v[4] += rf * rf;
                    if (dk_next) dk_next[g] = d3[k];
                    if (d4ok) {   // quartic: r1 - 4 d1 + 6 d2 - 4 d3 + d4
                        const double rg = c1[k] - 4.0 * (d1[k] + d3[k]) + 6.0 * d2[k] + dk_prev[g];
                        v[5] += rg * rg;
                    }
                }
            }
        }
    } else if (lx >= 1 && lx < TSI - 1 && gj < L.nx) {
        const double *g1 = nh >= 1 ? h0 : u;
        for (int k = 0; k < TROWS; ++k) {
            const int ly = ly0 + k, i = oy + ly, j = gj;
            if (ly < 1 || ly >= TSI - 1 || i < L.own0 || i >= L.own1) continue;
            const size_t g = (size_t)i * L.nx + j;
            double res1 = 0.0, resB = 0.0, e1 = 0.0, e2 = 0.0, e3 = 0.0;
            if (!is_dirichlet(L, i, j)) {
                const double b = load_row(L, i, j, u, rs_l, rs_r);
                double c[NBAND];
                stencil_iso(L, i, j, c);
                double ax, ag;
                frame_apply(L, dd, c, g1, i, j, ax, ag);
                res1 = b - ax;
                resB = b - ag;
                v[0] += res1 * res1;
                v[1] += resB * resB;
                if (MAXH >= 2 && nh >= 2) {
                    double ax1, ag1;
                    frame_apply(L, dd, c, h1, i, j, ax1, ag1);
                    e1 = ax - ax1;
                    v[2] += (res1 - e1) * (res1 - e1);
                    if (nh >= 3) {
                        double ax2, ag2;
                        frame_apply(L, dd, c, h2, i, j, ax2, ag2);
                        e2 = ax1 - ax2;
                        const double re = res1 - 2.0 * e1 + e2;
                        v[3] += re * re;
                        if (nh >= 4) {
                            double ax3, ag3;
                            frame_apply(L, dd, c, h3, i, j, ax3, ag3);
                            e3 = ax2 - ax3;
                            const double rf = res1 - 3.0 * (e1 - e2) - e3;
                            v[4] += rf * rf;
                            if (d4ok) {
                                const double rg = res1 - 4.0 * (e1 + e3) + 6.0 * e2 + dk_prev[g];
                                v[5] += rg * rg;
                            }
                        }
                    }
                }
            }
            r1[g] = res1;
            rB[g] = resB;
            if (MAXH >= 2 && nh >= 2) d1o[g] = e1;
            if (MAXH >= 3 && nh >= 3) d2o[g] = e2;
            if (MAXH >= 4 && nh >= 4) {
                d3o[g] = e3;
                if (dk_next) dk_next[g] = e3;   // zero on Dirichlet rows, like every d
            }
        }
    }
    double tot[6];
    if (grid_reduce<6>(v, partials, counter, tot)) {
        if (slab) {   // rank-local sums: one all-reduce of five doubles follows, k_slab_unpack hands them out
            sc->red_src[0] = tot[0]; sc->red_src[1] = tot[1]; sc->red_src[2] = tot[2]; sc->red_src[3] = tot[3];
            sc->red_src[4] = tot[4];
        } else {
            sc->rr0 = tot[0]; sc->bnorm2 = tot[1]; sc->rrD = tot[2]; sc->rrE = tot[3]; sc->rrF = tot[4];
            sc->rrG = d4ok ? tot[5] : 1.0e300;
        }
    }
}

// Per-node version of k_init_tile for the variable-tensor operator (one GPU): the same candidates -- previous
// solution (or the field as given without history), zero, linear and quadratic extrapolation -- with the row
// of the operator evaluated once per node (stencil_row) and applied to every history vector.
template <int TENSOR>
__global__ void __launch_bounds__(BX *BY)
k_init_hist(LevelDev L, DirData dd, const double *__restrict__ u, const double *__restrict__ h0,
            const double *__restrict__ h1, const double *__restrict__ h2, int nh, double *__restrict__ r1,
            double *__restrict__ rB, double *__restrict__ d1o, double *__restrict__ d2o, double rs_l, double rs_r,
            double *partials, unsigned *counter, CGScalars *sc)
{
    const int j = blockIdx.x * BX + threadIdx.x, i = blockIdx.y * BY + threadIdx.y;
    double v[4] = {0.0, 0.0, 0.0, 0.0};
    if (i >= L.own0 && i < L.own1 && j < L.nx) {
        const size_t g = (size_t)i * L.nx + j;
        const double *g1 = nh >= 1 ? h0 : u;
        double res1 = 0.0, resB = 0.0, e1 = 0.0, e2 = 0.0;
        if (!is_dirichlet(L, i, j)) {
            const double b = load_row(L, i, j, u, rs_l, rs_r);
            double c[NBAND];
            stencil_row<TENSOR>(L, i, j, c);
            double ax, ag;
            frame_apply(L, dd, c, g1, i, j, ax, ag);
            res1 = b - ax;
            resB = b - ag;
            v[0] = res1 * res1;
            v[1] = resB * resB;
            if (nh >= 2) {
                double ax1, ag1;
                frame_apply(L, dd, c, h1, i, j, ax1, ag1);
                e1 = ax - ax1;
                v[2] = (res1 - e1) * (res1 - e1);
                if (nh >= 3) {
                    double ax2, ag2;
                    frame_apply(L, dd, c, h2, i, j, ax2, ag2);
                    e2 = ax1 - ax2;
                    const double re = res1 - 2.0 * e1 + e2;
                    v[3] = re * re;
                }
            }
        }
        r1[g] = res1;
        rB[g] = resB;
        if (nh >= 2) d1o[g] = e1;
        if (nh >= 3) d2o[g] = e2;
    }
    double tot[4];
    if (grid_reduce<4>(v, partials, counter, tot)) {
        sc->rr0 = tot[0]; sc->bnorm2 = tot[1]; sc->rrD = tot[2]; sc->rrE = tot[3];
    }
}

// Least-squares starting guess (warm mode 4, single GPU).  k_init_tile leaves r1 = rB - A h0, the reduced
// right-hand side rB and the images d1 = A (h0 - h1), d2 = A (h1 - h2) of the history differences (free rows;
// zero on Dirichlet rows).  k_ls_gram sums the Gram matrix of a0 = A h0 = rB - r1, a1 = d1, a2 = d2 and their
// products with rB in one flat pass; its last block solves min || rB - c0 a0 - c1 a1 - c2 a2 || and leaves the
// coefficients in sc->lsc.  The guess c0 h0 + c1 (h0 - h1) + c2 (h1 - h2) is the best one in the span of the
// last three solutions, which contains the previous solution and both extrapolations (measured with the
// oracle on a 320^2 cut of the bench colony: 3e-8 of the zero guess's residual after 30 steps where the
// quadratic extrapolation has 5e-5).
__host__ __device__ __noinline__ void ls_solve3(const double G[6], const double f[3], double bb, double c[3], double &pred)
{
    // G = {a0.a0, a0.a1, a0.a2, a1.a1, a1.a2, a2.a2}.  Columns are scaled to unit norm; a vanishing column
    // (no history that deep, or two identical solutions) is dropped; a small ridge keeps the nearly
    // dependent differences of successive solutions solvable in fp64.
    const double gd[3] = {G[0], G[3], G[5]};
    double d[3], M[3][4];
    bool on[3];
    for (int i = 0; i < 3; ++i) {
        on[i] = gd[i] > 0.0 && gd[i] > 1e-30 * gd[0];
        d[i] = on[i] ? sqrt(gd[i]) : 1.0;
    }
    const double g[3][3] = {{G[0], G[1], G[2]}, {G[1], G[3], G[4]}, {G[2], G[4], G[5]}};
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) M[i][j] = (on[i] && on[j]) ? g[i][j] / (d[i] * d[j]) : 0.0;
        M[i][i] = on[i] ? M[i][i] + 1e-13 : 1.0;
        M[i][3] = on[i] ? f[i] / d[i] : 0.0;
    }
    bool ok = true;
    for (int k = 0; k < 3; ++k) {   // Gaussian elimination with partial pivoting
        int piv = k;
        for (int i = k + 1; i < 3; ++i) if (fabs(M[i][k]) > fabs(M[piv][k])) piv = i;
        if (!(fabs(M[piv][k]) > 1e-300)) { ok = false; break; }
        if (piv != k) for (int j = 0; j < 4; ++j) { const double t = M[k][j]; M[k][j] = M[piv][j]; M[piv][j] = t; }
        for (int i = k + 1; i < 3; ++i) {
            const double m = M[i][k] / M[k][k];
            for (int j = k; j < 4; ++j) M[i][j] -= m * M[k][j];
        }
    }
    double y[3] = {0.0, 0.0, 0.0};
    if (ok) {
        for (int i = 2; i >= 0; --i) {
            double t = M[i][3];
            for (int j = i + 1; j < 3; ++j) t -= M[i][j] * y[j];
            y[i] = t / M[i][i];
        }
    }
    for (int i = 0; i < 3; ++i) c[i] = (ok && on[i]) ? y[i] / d[i] : 0.0;
    double q = bb;
    for (int i = 0; i < 3; ++i) {
        q -= 2.0 * c[i] * f[i];
        for (int j = 0; j < 3; ++j) q += c[i] * g[i][j] * c[j];
    }
    const bool finite = ok && isfinite(c[0]) && isfinite(c[1]) && isfinite(c[2]) && isfinite(q);
    if (!finite) { c[0] = 1.0; c[1] = 0.0; c[2] = 0.0; }
    // below ~1e-16 bb the quadratic form is rounding noise: k_impose sums the true residual of the guess it forms
    pred = finite ? fmax(q, 0.0) : 1.0e300;
}

// host-side entry for the CPU tests of the 3x3 solve (eqgpu_ls_solve3)
void solver_ls_solve3(const double G[6], const double f[3], double bb, double c[3], double *pred)
{
    double p = 0.0;
    ls_solve3(G, f, bb, c, p);
    *pred = p;
}

// form 1 (default): the combination is a CORRECTION to the previous solution, u = h0 + c0 h0 + c1 (h0-h1) +
// c2 (h0-2h1+h2), fitted to r1 = rB - A h0 on the images {A h0, d1, d1-d2}.  Right-hand side and unknowns are
// then small (no c0 ~ 1 whose rounding error multiplies ||A h0|| ~ ||b||), and the second difference replaces
// the nearly parallel pair d1, d2; measured with the oracle this reaches 2e-12 of the zero guess's residual
// where form 0 (u = c0 h0 + c1 (h0-h1) + c2 (h1-h2) fitted to rB) stalls at 1e-10.
__global__ void __launch_bounds__(256)
k_ls_gram(size_t n, const double *__restrict__ r1, const double *__restrict__ rB, const double *__restrict__ d1,
          const double *__restrict__ d2, int nh, int form, double *partials, unsigned *counter, CGScalars *sc)
{
    double v[9];
#pragma unroll
    for (int q = 0; q < 9; ++q) v[q] = 0.0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < n; g += stride) {
        const double b = __ldg(rB + g), rv = __ldg(r1 + g), a0 = b - rv, e1 = __ldg(d1 + g);
        const double e2 = nh >= 3 ? __ldg(d2 + g) : 0.0;
        const double q2 = form ? (nh >= 3 ? e1 - e2 : 0.0) : e2;
        const double rhs = form ? rv : b;
        v[0] += a0 * a0; v[1] += a0 * e1; v[2] += a0 * q2;
        v[3] += e1 * e1; v[4] += e1 * q2; v[5] += q2 * q2;
        v[6] += a0 * rhs; v[7] += e1 * rhs; v[8] += q2 * rhs;
    }
    double tot[9];
    if (grid_reduce<9>(v, partials, counter, tot)) {
        double c[3], pred;
        ls_solve3(tot, tot + 6, form ? sc->rr0 : sc->bnorm2, c, pred);
        // d1 and d2 are stored as differences of stencil sums of size |A h|, so each carries an absolute error
        // of a few ulp of |A h0|; a coefficient c multiplies that error into the gap between the residual
        // vector k_impose forms and the true residual of its guess, a gap the PCG recurrence can never close.
        // Normal coefficients are O(1); when the history has stalled (steady state) and the right-hand side
        // then jumps, the fit reaches for c ~ 1e10 on pure rounding noise: such a combination is refused and
        // the fixed candidates stand (50 keeps the gap below 1e-13 ||b||).
        if (form && fabs(c[1]) + 2.0 * fabs(c[2]) > 50.0) pred = 1.0e300;
        sc->lsc[0] = c[0]; sc->lsc[1] = c[1]; sc->lsc[2] = c[2];
        sc->rrL = pred;
    }
}

// Pick the starting guess, impose u_d = g_d, set up the PCG scalars.
// Guess codes (sc->guess): 0 the field as given, 1 zero, 2 previous solution, 3 linear, 4 quadratic
// extrapolation, 5 least-squares combination, 6 cubic, 7 quartic extrapolation.  Without history (nh == 0: k_init, or k_init_tile's first step)
// h*/d* are unused.  ls != 0 (k_ls_gram ran, nh >= 2): the least-squares combination joins the candidates, and
// the squared residual of the guess actually formed is summed here, so that the PCG scalars start from a
// measured norm, not from the rounding-limited prediction.
__global__ void __launch_bounds__(BX *BY)
k_impose(LevelDev L, DirData dd, double *__restrict__ u, double *__restrict__ r,
         const double *__restrict__ rB, const double *__restrict__ d1, const double *__restrict__ d2,
         const double *__restrict__ h0, const double *__restrict__ h1, const double *__restrict__ h2, int nh,
         CGScalars *sc, double rtol, int max_iters, int ls, double *partials, unsigned *counter,
         const double *__restrict__ h3, const double *__restrict__ d3, const double *__restrict__ h4,
         const double *__restrict__ d4)
{
    const int j = blockIdx.x * BX + threadIdx.x, i = blockIdx.y * BY + threadIdx.y;
    const double rr1 = sc->rr0, rrB = sc->bnorm2;
    const double rrD = nh >= 2 ? sc->rrD : 1.0e300, rrE = nh >= 3 ? sc->rrE : 1.0e300;
    const double rrF = nh >= 4 ? sc->rrF : 1.0e300;
    const double rrL = ls ? sc->rrL : 1.0e300;
    int pick = nh >= 1 ? 2 : 0;
    double best = rr1;
    if (rrB < best) { best = rrB; pick = 1; }
    if (rrD < best) { best = rrD; pick = 3; }
    if (rrE < best) { best = rrE; pick = 4; }
    if (rrF < best) { best = rrF; pick = 6; }
    if (d4 && nh >= 5 && sc->rrG < best) { best = sc->rrG; pick = 7; }
    if (rrL <= best) { best = rrL; pick = 5; }
    double v[1] = {0.0};
    if (i >= L.own0 && i < L.own1 && j < L.nx) {
        const size_t g = (size_t)i * L.nx + j;
        const bool dir = is_dirichlet(L, i, j);
        if (dir) u[g] = dir_value(L, dd, i, j);
        else if (pick == 1) { u[g] = 0.0; r[g] = rB[g]; }
        else if (pick == 2) u[g] = h0[g];
        else if (pick == 3) { u[g] = 2.0 * h0[g] - h1[g]; r[g] -= d1[g]; }
        else if (pick == 4) { u[g] = 3.0 * (h0[g] - h1[g]) + h2[g]; r[g] += d2[g] - 2.0 * d1[g]; }
        else if (pick == 7) {   // quartic extrapolation of the last five solutions
            u[g] = 5.0 * (h0[g] - h3[g]) + 10.0 * (h2[g] - h1[g]) + h4[g];
            r[g] += 6.0 * d2[g] - 4.0 * (d1[g] + d3[g]) + d4[g];
        }
        else if (pick == 6) {   // cubic extrapolation of the last four solutions
            u[g] = 4.0 * (h0[g] + h2[g]) - 6.0 * h1[g] - h3[g];
            r[g] += 3.0 * (d2[g] - d1[g]) - d3[g];
        }
        else if (pick == 5) {
            const double c0 = sc->lsc[0], c1 = sc->lsc[1], c2 = sc->lsc[2];
            const double x0 = h0[g], x1 = h1[g], b = rB[g], rv = r[g];   // r holds r1 = rB - A h0 on entry
            double xn, rn;
            if (ls == 2) {   // form 1: correction to h0 fitted to r1
                xn = x0 + c0 * x0 + c1 * (x0 - x1);
                rn = rv - c0 * (b - rv) - c1 * d1[g];
                if (nh >= 3) { xn += c2 * (x0 - 2.0 * x1 + h2[g]); rn -= c2 * (d1[g] - d2[g]); }
            } else {         // form 0
                xn = c0 * x0 + c1 * (x0 - x1);
                rn = b - c0 * (b - rv) - c1 * d1[g];
                if (nh >= 3) { xn += c2 * (x1 - h2[g]); rn -= c2 * d2[g]; }
            }
            u[g] = xn;
            r[g] = rn;
        }
        if (ls && !dir) v[0] = r[g] * r[g];
    }
    if (ls) {   // a kernel argument: every block takes part in the reduction
        double tot[1];
        if (grid_reduce<1>(v, partials, counter, tot)) {
            const double stop2 = rtol * rtol * sc->bnorm2;
            sc->rr = tot[0];
            sc->rr_init = tot[0];
            sc->stop2 = stop2;
            sc->iters = 0;
            sc->max_iters = max_iters;
            sc->done = (tot[0] <= stop2) ? 1 : 0;
            sc->rz_old = 1.0;
            sc->rz_new = 0.0;
            sc->x_stamp = 0;
            sc->x_applied = 0;
            sc->guess = pick;
        }
        return;
    }
    __syncthreads();
    if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0 && threadIdx.y == 0) {
        const double stop2 = rtol * rtol * sc->bnorm2;
        // every block has read bnorm2/rr0/rrD/rrE before this block can be the last to
        // finish only if we do not overwrite them: keep them, write the rest.
        sc->rr = best;
        sc->rr_init = best;
        sc->stop2 = stop2;
        sc->iters = 0;
        sc->max_iters = max_iters;
        sc->done = (best <= stop2) ? 1 : 0;
        sc->rz_old = 1.0;
        sc->rz_new = 0.0;
        sc->x_stamp = 0;
        sc->x_applied = 0;
        sc->guess = pick;
    }
}

// k_impose for the short history (nh <= 1, no least-squares fit; one GPU, even row pitch): the candidates are the field as
// given (nh == 0) or the previous solution, and zero; 128-bit accesses over the flat field, Dirichlet nodes by index.
__global__ void __launch_bounds__(256)
k_impose_prev(LevelDev L, DirData dd, double *__restrict__ u, double *__restrict__ r, const double *__restrict__ rB,
              const double *__restrict__ h0, int nh, CGScalars *sc, double rtol, int max_iters)
{
    const double rr1 = sc->rr0, rrB = sc->bnorm2;
    int pick = nh >= 1 ? 2 : 0;
    double best = rr1;
    if (rrB < best) { best = rrB; pick = 1; }
    const size_t npairs = (size_t)L.nx * L.ny / 2;
    const unsigned m = L.dirmask;
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < npairs; q += (size_t)gridDim.x * blockDim.x) {
        const size_t g = 2 * q;
        const int i = (int)(g / L.nx), j = (int)(g - (size_t)i * L.nx);
        double2 uv;
        if (pick == 2) uv = *reinterpret_cast<const double2 *>(h0 + g);
        else if (pick == 1) { uv = make_double2(0.0, 0.0); *reinterpret_cast<double2 *>(r + g) = *reinterpret_cast<const double2 *>(rB + g); }
        else uv = *reinterpret_cast<const double2 *>(u + g);
        const bool rowd = ((m & 8u) && i == 0) || ((m & 4u) && i == L.ny - 1);
        if (rowd || ((m & 1u) && j == 0)) uv.x = dir_value(L, dd, i, j);
        if (rowd || ((m & 2u) && j + 1 == L.nx - 1)) uv.y = dir_value(L, dd, i, j + 1);
        *reinterpret_cast<double2 *>(u + g) = uv;
    }
    __syncthreads();
    if (blockIdx.x == 0 && threadIdx.x == 0) {   // (the other blocks only read rr0 / bnorm2, which stay as they are)
        const double stop2 = rtol * rtol * sc->bnorm2;
        sc->rr = best;
        sc->rr_init = best;
        sc->stop2 = stop2;
        sc->iters = 0;
        sc->max_iters = max_iters;
        sc->done = (best <= stop2) ? 1 : 0;
        sc->rz_old = 1.0;
        sc->rz_new = 0.0;
        sc->x_stamp = 0;
        sc->x_applied = 0;
        sc->guess = pick;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Warm mode 7 (opt-in; written after round 1's GPU budget was spent, first run in its last 20 seconds:
// profiles/r01_ring_first_run.json, 1.46 iterations per step against 2.66 on the bench, same field): image ring.
// The solver keeps the last K <= 7 solutions h_i AND their images a_i = A_ff h_i on the free rows (the newest image is
// b~ - r_final of its own step: no operator walk).  In the backward-difference basis nabla^j h0 (the Newton form, in
// which the fixed extrapolation through K solutions is the all-ones combination) the guess is
//     u = sum_j (1 + c_j) nabla^j h0 ,      c = argmin || rho - sum_j c_j nabla^j a0 || ,  rho = b~ - sum_j nabla^j a0
// i.e. a least-squares CORRECTION to the fixed extrapolation, fitted to its residual through the K x K normal equations.
// CPU study (profiles/r01_guess_study.md, a model of this solver that reproduces mode 6's 2.66 iterations per step):
// 1.5 iterations per step with K = 7, 1.96 with K = 5; the same fit of b~ itself, or in the first-difference basis, is
// lost to the conditioning of the Gram matrix (1e16 against 1e12 here, and the right-hand side is 1e-10 ||b|| here).
// ---------------------------------------------------------------------------------------------------------------
constexpr int RING_MAX = 7;
struct RingPtrs { const double *p[RING_MAX]; };

// out[j] = nabla^j v_0 = sum_i (-1)^i C(j,i) v_i, j < K (t is consumed)
template <int K>
__device__ __forceinline__ void backward_differences(double (&t)[K], double (&out)[K])
{
    out[0] = t[0];
#pragma unroll
    for (int j = 1; j < K; ++j) {
#pragma unroll
        for (int i = 0; i < K - j; ++i) t[i] = t[i] - t[i + 1];
        out[j] = t[0];
    }
}

// K x K normal equations G c = f, G packed upper triangle row-major (K(K+1)/2 entries).  Columns scaled to unit norm,
// vanishing columns dropped, ridge 1e-13, Gaussian elimination with partial pivoting (ls_solve3's recipe for any K).
__host__ __device__ __noinline__ void ring_solve(int K, const double *G, const double *f, double *c)
{
    double M[RING_MAX][RING_MAX + 1], d[RING_MAX], y[RING_MAX];
    bool on[RING_MAX];
    auto at = [&](int i, int j) { if (i > j) { const int t = i; i = j; j = t; } return G[i * K - i * (i - 1) / 2 + (j - i)]; };
    double gmax = 0.0;
    for (int i = 0; i < K; ++i) gmax = fmax(gmax, at(i, i));
    for (int i = 0; i < K; ++i) {
        const double g = at(i, i);
        on[i] = g > 0.0 && g > 1e-30 * gmax;
        d[i] = on[i] ? sqrt(g) : 1.0;
    }
    for (int i = 0; i < K; ++i) {
        for (int j = 0; j < K; ++j) M[i][j] = (on[i] && on[j]) ? at(i, j) / (d[i] * d[j]) : 0.0;
        M[i][i] = on[i] ? M[i][i] + 1e-13 : 1.0;
        M[i][K] = on[i] ? f[i] / d[i] : 0.0;
    }
    bool ok = true;
    for (int k = 0; k < K && ok; ++k) {
        int piv = k;
        for (int i = k + 1; i < K; ++i) if (fabs(M[i][k]) > fabs(M[piv][k])) piv = i;
        if (!(fabs(M[piv][k]) > 1e-300)) { ok = false; break; }
        if (piv != k) for (int j = 0; j <= K; ++j) { const double t = M[k][j]; M[k][j] = M[piv][j]; M[piv][j] = t; }
        for (int i = k + 1; i < K; ++i) {
            const double m = M[i][k] / M[k][k];
            for (int j = k; j <= K; ++j) M[i][j] -= m * M[k][j];
        }
    }
    for (int i = 0; i < K; ++i) y[i] = 0.0;
    if (ok)
        for (int i = K - 1; i >= 0; --i) {
            double t = M[i][K];
            for (int j = i + 1; j < K; ++j) t -= M[i][j] * y[j];
            y[i] = t / M[i][i];
        }
    bool finite = ok;
    for (int i = 0; i < K; ++i) { c[i] = (ok && on[i]) ? y[i] / d[i] : 0.0; finite = finite && isfinite(c[i]); }
    if (!finite) for (int i = 0; i < K; ++i) c[i] = 0.0;   // no correction: the fixed extrapolation stands
}

void solver_ring_solve(int K, const double *G, const double *f, double *c)
{
    if (K >= 1 && K <= RING_MAX) ring_solve(K, G, f, c);
}

// Pass 1: one flat pass over b~ and the K images.  Sums the Gram matrix of nabla^j a0, its products with rho and
// ||rho||^2; the last block solves for the correction and leaves the weights 1 + c_j and the predicted squared
// residual in sc.  Dirichlet rows hold zeros in b~ and in every image, so they drop out of the sums.
template <int K>
__global__ void __launch_bounds__(256)
k_ring_gram(size_t n, const double *__restrict__ rB, RingPtrs a, double *partials, unsigned *counter, CGScalars *sc)
{
    constexpr int NG = K * (K + 1) / 2, NS = NG + K + 1;
    double v[NS];
#pragma unroll
    for (int q = 0; q < NS; ++q) v[q] = 0.0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < n; g += stride) {
        double t[K], da[K];
#pragma unroll
        for (int i = 0; i < K; ++i) t[i] = __ldg(a.p[i] + g);
        backward_differences<K>(t, da);
        double rho = __ldg(rB + g);
#pragma unroll
        for (int j = 0; j < K; ++j) rho -= da[j];
        int q = 0;
#pragma unroll
        for (int i = 0; i < K; ++i)
#pragma unroll
            for (int j = i; j < K; ++j) v[q++] += da[i] * da[j];
#pragma unroll
        for (int i = 0; i < K; ++i) v[NG + i] += da[i] * rho;
        v[NG + K] += rho * rho;
    }
    double tot[NS];
    if (grid_reduce<NS>(v, partials, counter, tot)) {
        double c[RING_MAX];
        ring_solve(K, tot, tot + NG, c);
        // predicted squared residual of the corrected guess; rounding noise below ~1e-16 of the sums, so the true one is
        // summed again by k_ring_impose.  A coefficient multiplies the rounding error of its difference (2^j ulp of
        // |A h0|) into the gap between the residual vector formed and the true residual: refuse wild fits.
        double q = tot[NG + K], amp = 0.0;
        for (int i = 0; i < K; ++i) {
            q -= 2.0 * c[i] * tot[NG + i];
            amp += fabs(c[i]) * (double)(1 << i);
            for (int j = 0; j < K; ++j) {
                const int lo = i < j ? i : j, hi = i < j ? j : i;
                q += c[i] * tot[lo * K - lo * (lo - 1) / 2 + (hi - lo)] * c[j];
            }
        }
        if (!(amp <= 1.0e4) || !isfinite(q)) { for (int i = 0; i < K; ++i) c[i] = 0.0; q = tot[NG + K]; }
        for (int i = 0; i < RING_MAX + 1; ++i) sc->ringw[i] = i < K ? 1.0 + c[i] : 0.0;
        sc->rrR = fmax(q, 0.0);
        sc->ring_k = K;
    }
}

// Pass 2: pick between the field as given (0), zero (1) and the ring guess (8) by squared residual, form u and r,
// impose u_d = g_d, sum the TRUE squared residual of what was formed and set up the PCG scalars (k_impose's job).
// On entry r = b~ - A u0 (k_init_tile without history) and rB = b~.
template <int K>
__global__ void __launch_bounds__(BX *BY)
k_ring_impose(LevelDev L, DirData dd, double *__restrict__ u, double *__restrict__ r, const double *__restrict__ rB,
              RingPtrs h, RingPtrs a, CGScalars *sc, double rtol, int max_iters, double *partials, unsigned *counter)
{
    const int j = blockIdx.x * BX + threadIdx.x, i = blockIdx.y * BY + threadIdx.y;
    int pick = 0;
    double best = sc->rr0;
    if (sc->bnorm2 < best) { best = sc->bnorm2; pick = 1; }
    if (sc->rrR <= best) { best = sc->rrR; pick = 8; }
    double v[1] = {0.0};
    if (i >= L.own0 && i < L.own1 && j < L.nx) {
        const size_t g = (size_t)i * L.nx + j;
        const bool dir = is_dirichlet(L, i, j);
        if (dir) u[g] = dir_value(L, dd, i, j);
        else if (pick == 1) { u[g] = 0.0; r[g] = rB[g]; }
        else if (pick == 8) {
            double t[K], dv[K];
#pragma unroll
            for (int k = 0; k < K; ++k) t[k] = __ldg(h.p[k] + g);
            backward_differences<K>(t, dv);
            double un = 0.0;
#pragma unroll
            for (int k = K - 1; k >= 0; --k) un += sc->ringw[k] * dv[k];      // smallest terms first
#pragma unroll
            for (int k = 0; k < K; ++k) t[k] = __ldg(a.p[k] + g);
            backward_differences<K>(t, dv);
            double corr = 0.0;
#pragma unroll
            for (int k = K - 1; k >= 0; --k) corr += sc->ringw[k] * dv[k];
            u[g] = un;
            r[g] = rB[g] - corr;
        }
        if (!dir) v[0] = r[g] * r[g];
    }
    double tot[1];
    if (grid_reduce<1>(v, partials, counter, tot)) {
        const double stop2 = rtol * rtol * sc->bnorm2;
        sc->rr = tot[0];
        sc->rr_init = tot[0];
        sc->stop2 = stop2;
        sc->iters = 0;
        sc->max_iters = max_iters;
        sc->done = (tot[0] <= stop2) ? 1 : 0;
        sc->rz_old = 1.0;
        sc->rz_new = 0.0;
        sc->x_stamp = 0;
        sc->x_applied = 0;
        sc->guess = pick;
    }
}

// End of a converged step: the image of the solution just found, a = A_ff u_f on the free rows (zero on Dirichlet
// rows, columns to Dirichlet nodes dropped, so the images survive setBoundaryValues).  It is evaluated with one
// operator walk over the solution, NOT taken as b~ - r_final: the PCG recurrence residual starts from a vector built
// out of the earlier images, so an image taken from it would carry their error forward through extrapolation weights
// whose absolute sum is 127 at K = 7 -- a drift that the stopping test (measured on that same recurrence) cannot see
// (1e-8 after 300 steps in a CPU model of the recursion).  With exact images the residual k_ring_impose forms is the
// true residual of its guess up to rounding, every step anew.
__global__ void __launch_bounds__(BX *BY)
k_ring_image(LevelDev L, const double *__restrict__ u, double *__restrict__ a)
{
    const int j = blockIdx.x * BX + threadIdx.x, i = blockIdx.y * BY + threadIdx.y;
    if (i >= L.ny || j >= L.nx) return;
    const size_t g = (size_t)i * L.nx + j;
    double out = 0.0;
    if (!is_dirichlet(L, i, j)) {
        double c[NBAND];
        stencil_iso(L, i, j, c);
        const bool hasW = j > 0, hasE = j < L.nx - 1, hasS = i > 0, hasN = i < L.ny - 1;
        out = c[B_C] * __ldg(u + g);
        auto acc = [&](int ii, int jj, double ck) {
            if (!is_dirichlet(L, ii, jj)) out += ck * __ldg(u + (size_t)ii * L.nx + jj);
        };
        if (hasE) acc(i, j + 1, c[B_E]);
        if (hasW) acc(i, j - 1, c[B_W]);
        if (hasN) acc(i + 1, j, c[B_N]);
        if (hasS) acc(i - 1, j, c[B_S]);
        if (hasN && hasE) acc(i + 1, j + 1, c[B_NE]);
        if (hasS && hasW) acc(i - 1, j - 1, c[B_SW]);
    }
    a[g] = out;
}

template <int K>
static void launch_ring(eqgpu_solver *s, const LevelDev &L, const DirData &dd, const RingPtrs &h, const RingPtrs &a, int nb1,
                        dim3 g0, dim3 blk, double rtol, int max_iters)
{
    k_ring_gram<K><<<nb1, 256, 0, s->stream>>>(s->N, s->z, a, s->ring_partials, s->counters + 6, s->sc);
    k_ring_impose<K><<<g0, blk, 0, s->stream>>>(L, dd, s->u, s->r, s->z, h, a, s->sc, rtol, max_iters, s->partials,
                                                 s->counters + 7);
}

// rz_new = r . z
__global__ void __launch_bounds__(256)
k_dot(size_t n, const double *__restrict__ a, const double *__restrict__ b, CGScalars *sc,
      double *partials, unsigned *counter, double *out)
{
    if (sc->done) return;
    double v[1] = {0.0};
    for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < n;
         g += (size_t)gridDim.x * blockDim.x)
        v[0] += __ldg(a + g) * __ldg(b + g);
    double tot[1];
    if (grid_reduce<1>(v, partials, counter, tot)) *out = tot[0];
}

// p = z + beta p
__global__ void __launch_bounds__(256)
k_update_p(size_t n, const double *__restrict__ z, double *__restrict__ p, const CGScalars *sc)
{
    if (sc->done) return;
    const double beta = sc->iters == 0 ? 0.0 : sc->rz_new / sc->rz_old;
    for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < n;
         g += (size_t)gridDim.x * blockDim.x)
        p[g] = __ldg(z + g) + beta * p[g];
}

// x += alpha p ; r -= alpha Ap ; rr = r.r ; bookkeeping in the last block.
// 128-bit accesses, two independent double2 per thread per trip.
__global__ void __launch_bounds__(256)
k_update_xr(size_t n, double *__restrict__ x, double *__restrict__ r, const double *__restrict__ p,
            const double *__restrict__ Ap, CGScalars *sc, double *partials, unsigned *counter, int book,
            double *out_rr)
{
    pdl_trigger();
    pdl_wait();
    if (sc->done) return;
    const double alpha = sc->rz_new / sc->pAp;
    double v[1] = {0.0};
    const bool aligned = ((((size_t)x) | ((size_t)r) | ((size_t)p) | ((size_t)Ap)) & 15) == 0;
    const size_t n2 = aligned ? n >> 1 : 0, stride = (size_t)gridDim.x * blockDim.x;
    if (!aligned) {  // a slab whose first owned row starts at an odd element: scalar accesses
        for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += stride) {
            x[t] += alpha * __ldg(p + t);
            const double rn = r[t] - alpha * __ldg(Ap + t);
            r[t] = rn;
            v[0] += rn * rn;
        }
    }
    double2 *x2 = reinterpret_cast<double2 *>(x), *r2 = reinterpret_cast<double2 *>(r);
    const double2 *p2 = reinterpret_cast<const double2 *>(p), *A2 = reinterpret_cast<const double2 *>(Ap);
    size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; g + stride < n2; g += 2 * stride) {
        const size_t h = g + stride;
        double2 xa = x2[g], xb = x2[h], ra = r2[g], rb = r2[h];
        const double2 pa = __ldg(p2 + g), pb = __ldg(p2 + h), aa = __ldg(A2 + g), ab = __ldg(A2 + h);
        xa.x += alpha * pa.x; xa.y += alpha * pa.y; xb.x += alpha * pb.x; xb.y += alpha * pb.y;
        ra.x -= alpha * aa.x; ra.y -= alpha * aa.y; rb.x -= alpha * ab.x; rb.y -= alpha * ab.y;
        x2[g] = xa; x2[h] = xb; r2[g] = ra; r2[h] = rb;
        v[0] += ra.x * ra.x + ra.y * ra.y + rb.x * rb.x + rb.y * rb.y;
    }
    for (; g < n2; g += stride) {
        double2 xa = x2[g], ra = r2[g];
        const double2 pa = __ldg(p2 + g), aa = __ldg(A2 + g);
        xa.x += alpha * pa.x; xa.y += alpha * pa.y;
        ra.x -= alpha * aa.x; ra.y -= alpha * aa.y;
        x2[g] = xa; r2[g] = ra;
        v[0] += ra.x * ra.x + ra.y * ra.y;
    }
    if (aligned && (n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
        const size_t t = n - 1;
        x[t] += alpha * p[t];
        const double rn = r[t] - alpha * Ap[t];
        r[t] = rn;
        v[0] += rn * rn;
    }
    double tot[1];
    if (grid_reduce<1>(v, partials, counter, tot)) {
        *out_rr = tot[0];
        if (book) {
            sc->rz_old = sc->rz_new;
            sc->iters += 1;
            if (tot[0] <= sc->stop2 || sc->iters >= sc->max_iters) sc->done = 1;
        }
    }
}

// Split form of k_update_xr for the single-GPU fused path.  Only r is on the critical path of the next
// iteration (the V-cycle smooths r), so k_update_r updates r alone (24 B/DOF) and leaves the step length for
// k_update_x, which adds alpha*p to x while the next iteration works through its latency-bound coarse
// levels -- that window leaves most of the HBM bandwidth idle.
__global__ void __launch_bounds__(256)
k_update_r(size_t n, double *__restrict__ r, const double *__restrict__ Ap, CGScalars *sc, double *partials,
           unsigned *counter, double *out_part = nullptr)
{
    pdl_trigger();
    pdl_wait();
    if (sc->done) return;
    const double alpha = sc->rz_new / sc->pAp;
    double v[1] = {0.0};
    const size_t n2 = n >> 1, stride = (size_t)gridDim.x * blockDim.x;
    double2 *r2 = reinterpret_cast<double2 *>(r);
    const double2 *A2 = reinterpret_cast<const double2 *>(Ap);
    size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; g + stride < n2; g += 2 * stride) {
        const size_t h = g + stride;
        double2 ra = r2[g], rb = r2[h];
        const double2 aa = __ldg(A2 + g), ab = __ldg(A2 + h);
        ra.x -= alpha * aa.x; ra.y -= alpha * aa.y; rb.x -= alpha * ab.x; rb.y -= alpha * ab.y;
        r2[g] = ra; r2[h] = rb;
        v[0] += ra.x * ra.x + ra.y * ra.y + rb.x * rb.x + rb.y * rb.y;
    }
    for (; g < n2; g += stride) {
        double2 ra = r2[g];
        const double2 aa = __ldg(A2 + g);
        ra.x -= alpha * aa.x; ra.y -= alpha * aa.y;
        r2[g] = ra;
        v[0] += ra.x * ra.x + ra.y * ra.y;
    }
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
        const size_t t = n - 1;
        const double rn = r[t] - alpha * Ap[t];
        r[t] = rn;
        v[0] += rn * rn;
    }
    double tot[1];
    if (grid_reduce<1>(v, partials, counter, tot)) {
        if (out_part) {   // row slabs: this rank's share; summed over the ranks, then k_book_x does the bookkeeping
            *out_part = tot[0];
            return;
        }
        sc->rr = tot[0];
        sc->rz_old = sc->rz_new;
        sc->iters += 1;
        sc->alpha_x = alpha;
        sc->x_stamp = sc->iters;
        if (tot[0] <= sc->stop2 || sc->iters >= sc->max_iters) sc->done = 1;
    }
}

// x += alpha_x * p for the pending iteration, p being the search direction of that iteration: odd iterations
// leave it in p_odd, even ones in p_even (the p ping-pong).  Does nothing when no update is pending; the
// pending mark is cleared by the next kernel on the main path (k_apply_p or k_mark_x), after this one completed.
__global__ void __launch_bounds__(256)
k_update_x(size_t n, double *__restrict__ x, const double *__restrict__ p_odd, const double *__restrict__ p_even,
           const CGScalars *sc)
{
    if (sc->x_applied == sc->x_stamp) return;
    const double alpha = sc->alpha_x;
    const double *p = (sc->x_stamp & 1) ? p_odd : p_even;
    const size_t n2 = n >> 1, stride = (size_t)gridDim.x * blockDim.x;
    double2 *x2 = reinterpret_cast<double2 *>(x);
    const double2 *p2 = reinterpret_cast<const double2 *>(p);
    size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; g + stride < n2; g += 2 * stride) {
        const size_t h = g + stride;
        double2 xa = x2[g], xb = x2[h];
        const double2 pa = __ldg(p2 + g), pb = __ldg(p2 + h);
        xa.x += alpha * pa.x; xa.y += alpha * pa.y; xb.x += alpha * pb.x; xb.y += alpha * pb.y;
        x2[g] = xa; x2[h] = xb;
    }
    for (; g < n2; g += stride) {
        double2 xa = x2[g];
        const double2 pa = __ldg(p2 + g);
        xa.x += alpha * pa.x; xa.y += alpha * pa.y;
        x2[g] = xa;
    }
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) x[n - 1] += alpha * p[n - 1];
}

// End of a step: the pending x update, if any, and (hist_out != null) a copy of the solution for the next
// step's warm start, in one pass.
__global__ void __launch_bounds__(256)
k_finish_x(size_t n, double *__restrict__ x, const double *__restrict__ p_odd, const double *__restrict__ p_even,
           const CGScalars *sc, double *__restrict__ hist_out)
{
    const bool pending = sc->x_applied != sc->x_stamp;
    if (!pending && !hist_out) return;
    const double alpha = pending ? sc->alpha_x : 0.0;
    const double *p = (sc->x_stamp & 1) ? p_odd : p_even;
    const size_t n2 = n >> 1, stride = (size_t)gridDim.x * blockDim.x;
    double2 *x2 = reinterpret_cast<double2 *>(x), *h2 = reinterpret_cast<double2 *>(hist_out);
    const double2 *p2 = reinterpret_cast<const double2 *>(p);
    for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < n2; g += stride) {
        double2 xa = x2[g];
        if (pending) {
            const double2 pa = __ldg(p2 + g);
            xa.x += alpha * pa.x; xa.y += alpha * pa.y;
            x2[g] = xa;
        }
        if (hist_out) h2[g] = xa;
    }
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
        double xa = x[n - 1];
        if (pending) { xa += alpha * p[n - 1]; x[n - 1] = xa; }
        if (hist_out) hist_out[n - 1] = xa;
    }
}

__global__ void k_mark_x(CGScalars *sc) { sc->x_applied = sc->x_stamp; }

// slab mode: the rank-summed start-of-step norms (one fused all-reduce) to where k_impose reads them
__global__ void k_slab_unpack(CGScalars *sc)
{
    sc->rr0 = sc->red_dst[0]; sc->bnorm2 = sc->red_dst[1]; sc->rrD = sc->red_dst[2]; sc->rrE = sc->red_dst[3];
    sc->rrF = sc->red_dst[4];
}

// iteration bookkeeping when ||r||^2 had to be summed over ranks first (slab mode)
__global__ void k_book(CGScalars *sc)
{
    if (sc->done) return;
    sc->rz_old = sc->rz_new;
    sc->iters += 1;
    if (sc->rr <= sc->stop2 || sc->iters >= sc->max_iters) sc->done = 1;
}

// ... and with the deferred x update (k_update_r / k_update_x) on slabs: the step length and its stamp as k_update_r leaves them
__global__ void k_book_x(CGScalars *sc)
{
    if (sc->done) return;
    sc->alpha_x = sc->rz_new / sc->pAp;
    sc->rz_old = sc->rz_new;
    sc->iters += 1;
    sc->x_stamp = sc->iters;
    if (sc->rr <= sc->stop2 || sc->iters >= sc->max_iters) sc->done = 1;
}

// ---------------------------------------------------------------------------
// multigrid pieces
// ---------------------------------------------------------------------------
// x = omega * D^-1 b   (first sweep from a zero guess)
template <int TENSOR>
__global__ void __launch_bounds__(BX *BY)
k_jacobi0(LevelDev L, const double *__restrict__ b, double *__restrict__ x, double omega,
          const CGScalars *sc)
{
    if (sc->done) return;
    const int j = blockIdx.x * BX + threadIdx.x, i = blockIdx.y * BY + threadIdx.y;
    if (i < L.own0 || i >= L.own1 || j >= L.nx) return;
    const size_t g = (size_t)i * L.nx + j;
    double out = 0.0;
    if (!is_dirichlet(L, i, j)) {
        double c[NBAND];
        stencil_row<TENSOR>(L, i, j, c);
        out = omega * __ldg(b + g) / c[B_C];
    }
    x[g] = out;
}

// xout = xin + omega * D^-1 (b - A xin);  RESID: xout = b - A xin instead
template <int TENSOR, bool RESID>
__global__ void __launch_bounds__(BX *BY)
k_jacobi(LevelDev L, const double *__restrict__ b, const double *__restrict__ xin,
         double *__restrict__ xout, double omega, const CGScalars *sc)
{
    if (sc->done) return;
    const int j = blockIdx.x * BX + threadIdx.x, i = blockIdx.y * BY + threadIdx.y;
    if (i < L.own0 || i >= L.own1 || j >= L.nx) return;
    const size_t g = (size_t)i * L.nx + j;
    double out = 0.0;
    if (!is_dirichlet(L, i, j)) {
        double c[NBAND];
        stencil_row<TENSOR>(L, i, j, c);
        const double res = __ldg(b + g) - stencil_dot(L, i, j, c, xin);
        out = RESID ? res : __ldg(xin + g) + omega * res / c[B_C];
    }
    xout[g] = out;
}


// bc = P^T rf  (P = P1 interpolation on the "right" mesh: midpoints of the
// E-W, N-S and SW-NE edges take half of each end).  Row logic uses global rows
// (row0/gny) so the same kernel serves a row slab.
__global__ void __launch_bounds__(BX *BY)
k_restrict(LevelDev F, LevelDev Cc, const double *__restrict__ rf, double *__restrict__ bc,
           const CGScalars *sc)
{
    if (sc->done) return;
    const int J = blockIdx.x * BX + threadIdx.x, I = blockIdx.y * BY + threadIdx.y;
    if (I < Cc.own0 || I >= Cc.own1 || J >= Cc.nx) return;
    double out = 0.0;
    if (!is_dirichlet(Cc, I, J)) {
        const int gfi = fine_of(I + Cc.row0, F.gny), fj = fine_of(J, F.nx);
        const int fi = gfi - F.row0;
        const size_t g = (size_t)fi * F.nx + fj;
        const bool e = fj + 1 < F.nx && is_mid(fj + 1, F.nx);
        const bool w = fj - 1 >= 0 && is_mid(fj - 1, F.nx);
        const bool n = gfi + 1 < F.gny && is_mid(gfi + 1, F.gny);
        const bool s = gfi - 1 >= 0 && is_mid(gfi - 1, F.gny);
        out = __ldg(rf + g);
        double h = 0.0;
        if (e) h += __ldg(rf + g + 1);
        if (w) h += __ldg(rf + g - 1);
        if (n) h += __ldg(rf + g + F.nx);
        if (s) h += __ldg(rf + g - F.nx);
        if (n && e) h += __ldg(rf + g + F.nx + 1);
        if (s && w) h += __ldg(rf + g - F.nx - 1);
        out += 0.5 * h;
    }
    bc[(size_t)I * Cc.nx + J] = out;
}

// xf += P xc
__global__ void __launch_bounds__(BX *BY)
k_prolong_add(LevelDev F, LevelDev Cc, const double *__restrict__ xc, double *__restrict__ xf,
              const CGScalars *sc)
{
    if (sc->done) return;
    const int j = blockIdx.x * BX + threadIdx.x, i = blockIdx.y * BY + threadIdx.y;
    if (i < F.own0 || i >= F.own1 || j >= F.nx) return;
    if (is_dirichlet(F, i, j)) return;
    const int gi = i + F.row0;
    const bool mi = is_mid(gi, F.gny), mj = is_mid(j, F.nx);
    const double *c0 = xc + (size_t)(coarse_lo(gi, F.gny, Cc.gny) - Cc.row0) * Cc.nx + coarse_lo(j, F.nx, Cc.nx);
    double add;
    if (!mi && !mj) add = __ldg(c0);
    else if (!mi && mj) add = 0.5 * (__ldg(c0) + __ldg(c0 + 1));
    else if (mi && !mj) add = 0.5 * (__ldg(c0) + __ldg(c0 + Cc.nx));
    else add = 0.5 * (__ldg(c0) + __ldg(c0 + Cc.nx + 1));
    xf[(size_t)i * F.nx + j] += add;
}

// assembled rows of the variable-tensor operator, packed by symmetry (LevelDev::kC ...)
__global__ void __launch_bounds__(BX *BY)
k_coef_assemble(LevelDev L, double *__restrict__ kC, double *__restrict__ kE, double *__restrict__ kN,
                double *__restrict__ kNE)
{
    const int j = blockIdx.x * BX + threadIdx.x, i = blockIdx.y * BY + threadIdx.y;
    if (i >= L.ny || j >= L.nx) return;
    double c[NBAND];
    stencil_tensor(L, i, j, c);
    const size_t g = (size_t)i * L.nx + j;
    kC[g] = c[B_C]; kE[g] = c[B_E]; kN[g] = c[B_N]; kNE[g] = c[B_NE];
}

// nodal injection of a tensor component to the coarse grid
__global__ void k_inject(LevelDev F, LevelDev Cc, const double *__restrict__ f, double *__restrict__ c)
{
    const int J = blockIdx.x * BX + threadIdx.x, I = blockIdx.y * BY + threadIdx.y;
    if (I >= Cc.ny || J >= Cc.nx) return;
    c[(size_t)I * Cc.nx + J] = f[(size_t)(fine_of(I + Cc.row0, F.gny) - F.row0) * F.nx + fine_of(J, F.nx)];
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
static dim3 grid2d(const LevelDev &L) { return dim3((L.nx + BX - 1) / BX, (L.ny + BY - 1) / BY); }

static std::vector<double> coarsen_cells(const std::vector<double> &h)
{
    const int n = (int)h.size() + 1;  // nodes
    const int nc = n / 2 + 1;
    std::vector<double> hc(nc - 1);
    for (int J = 0; J < nc - 1; ++J) {
        const int f0 = std::min(2 * J, n - 1), f1 = std::min(2 * J + 2, n - 1);
        double w = 0.0;
        for (int k = f0; k < f1; ++k) w += h[k];
        hc[J] = w;
    }
    return hc;
}

// Padded cell sizes of nodes [first, first+count): entry k = size of the cell below/left of node first+k.
static int upload_padded(eqgpu_solver *s, const std::vector<double> &h, double **d_h, double **d_ih,
                         int first = 0, int count = -1)
{
    const int n = (int)h.size() + 1;
    if (count < 0) count = n;
    std::vector<double> pad(n + 1, 0.0), ipad(n + 1, 0.0);
    for (int k = 0; k < n - 1; ++k) { pad[k + 1] = h[k]; ipad[k + 1] = 1.0 / h[k]; }
    EQ_CUDA(cudaMalloc(d_h, sizeof(double) * (count + 1)));
    EQ_CUDA(cudaMalloc(d_ih, sizeof(double) * (count + 1)));
    EQ_CUDA(cudaMemcpy(*d_h, pad.data() + first, sizeof(double) * (count + 1), cudaMemcpyHostToDevice));
    EQ_CUDA(cudaMemcpy(*d_ih, ipad.data() + first, sizeof(double) * (count + 1), cudaMemcpyHostToDevice));
    return 0;
}

static int leading_regular(const std::vector<double> &h)
{
    int k = 0;
    while (k < (int)h.size() && h[k] == h[0]) ++k;
    return k;  // cells 0..k-1 are regular
}

static void fill_level_consts(eqgpu_solver *s, Level &lv)
{
    LevelDev &L = lv.dev;
    const eqgpu_params &p = s->p;
    L.tau = p.dt * p.D;
    L.rob_l = p.bc_type[EQGPU_LEFT] == EQGPU_BC_ROBIN ? p.dt * p.bc_value[EQGPU_LEFT] : 0.0;
    L.rob_r = p.bc_type[EQGPU_RIGHT] == EQGPU_BC_ROBIN ? p.dt * p.bc_value[EQGPU_RIGHT] : 0.0;
    unsigned m = 0;
    for (int w = 0; w < 4; ++w)
        if (p.bc_type[w] == EQGPU_BC_DIRICHLET || p.bc_type[w] == EQGPU_BC_DIRICHLET_CHANNEL)
            m |= 1u << w;
    const unsigned mglob = m;
    // a slab sees the top/bottom walls only if its local first/last row IS the global wall row
    if (L.row0 > 0) m &= ~8u;
    if (L.row0 + L.ny < L.gny) m &= ~4u;
    L.dirmask = m;
    const double a = lv.hx_host[0], b = lv.hy_host[0];
    // node j (1 <= j) is regular when cells j-1 and j are: j <= leading_regular-1
    L.jreg_hi = std::min(leading_regular(lv.hx_host) - 1, L.nx - 2);
    L.ireg_hi = std::min(leading_regular(lv.hy_host) - 1, L.gny - 2) - L.row0;  // local index
    L.lumped = p.discretisation == EQGPU_DISC_FD ? 1 : 0;
    if (L.lumped) {   // 5-point row times h^2: (1 + 4F) and -F of diffuclass.cpp:857-860
        L.mC = a * b; L.mO = 0.0;
        L.cC = 2.0 * L.tau * (b / a + a / b) + a * b;
        L.cEW = -L.tau * b / a;
        L.cNS = -L.tau * a / b;
        L.cD = 0.0;
    } else {
        L.mC = a * b * 0.5; L.mO = a * b / 12.0;
        L.cC = 2.0 * L.tau * (b / a + a / b) + a * b * 0.5;
        L.cEW = a * b / 12.0 - L.tau * b / a;
        L.cNS = a * b / 12.0 - L.tau * a / b;
        L.cD = a * b / 12.0;
    }
    L.icC = 1.0 / L.cC;
    L.hxr = lv.hx_host.front(); L.hxl = lv.hx_host.back(); L.hyr = lv.hy_host.front(); L.hyl = lv.hy_host.back();
    L.hx = lv.d_hx; L.ihx = lv.d_ihx; L.hy = lv.d_hy; L.ihy = lv.d_ihy;
    L.d11 = lv.t11; L.d22 = lv.t22; L.d12 = lv.t12;
    L.kC = nullptr; L.kE = nullptr; L.kN = nullptr; L.kNE = nullptr;   // set by solver_refresh_levels once assembled
    // global view for the tile kernels: whole-grid index logic, windows saying which rows are stored / owned
    LevelDev &G = lv.gdev;
    G = L;
    if (s->slab_fused) {
        G.ny = L.gny;
        G.row0 = 0;
        G.own0 = 0; G.own1 = L.gny;
        G.dirmask = mglob;
        G.ireg_hi = L.ireg_hi + L.row0;
        G.hy = lv.d_hy_g; G.ihy = lv.d_ihy_g;
        G.slo = L.row0; G.shi = L.row0 + L.ny;
        G.wlo = lv.g0; G.whi = lv.g1;
        G.tbase = lv.g0 & ~1;
    }
}

static CTailDesc make_ctail_desc(eqgpu_solver *s, int first, int ncta);
static CoarseW coarse_weights(eqgpu_solver *s);
static bool make_field_map(CUtensorMap *map, double *base, int nx, int ny);

int solver_setup(eqgpu_solver *s)
{
    const eqgpu_params &p = s->p;
    s->N = (size_t)p.nW * p.nH;  // replaced by the local size once the slab window is known
    const double hx0 = p.hx, hy0 = p.hy > 0 ? p.hy : p.hx;
    EQ_CUDA(cudaFuncSetAttribute((k_init_tile<INIT_THREADS, 4>), cudaFuncAttributeMaxDynamicSharedMemorySize, INIT_SMEM));
    EQ_CUDA(cudaFuncSetAttribute((k_init_tile<512, 1>), cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * INIT_SMEM / 5));
    s->nu = p.smooth_sweeps > 0 ? p.smooth_sweeps : 3;
    // one more sweep on the coarser levels: they are latency-bound, so it is nearly free, and it widens the
    // margin at the 8th iteration (relres 1.6e-13 vs 7.2e-13 at 2048^2; measured 488 vs 479 steps/s)
    s->nuc = p.smooth_sweeps > 0 ? s->nu : 4;
    if (const char *e = getenv("EQGPU_NU0")) s->nu = std::max(1, std::min(atoi(e), 4));   // tuning knobs
    if (const char *e = getenv("EQGPU_NUC")) s->nuc = std::max(1, std::min(atoi(e), 4));
    s->nu1 = s->nuc;
    if (const char *e = getenv("EQGPU_NU1")) s->nu1 = std::max(1, std::min(atoi(e), 4));
    // ---- hierarchy -------------------------------------------------------
    Level l0;
    l0.dev.nx = p.nW; l0.dev.gny = p.nH;
    l0.hx_host.assign(p.nW - 1, hx0);
    l0.hy_host.assign(p.nH - 1, hy0);
    l0.g0 = 0; l0.g1 = p.nH;
    if (s->slab) {  // contiguous row slabs with even boundaries
        auto cut = [&](int r) { return r >= s->slab_world ? p.nH : (int)(((long long)p.nH * r / s->slab_world) & ~1LL); };
        l0.g0 = cut(s->slab_rank); l0.g1 = cut(s->slab_rank + 1);
        if (l0.g1 - l0.g0 < 16) { s->set_error("slab too thin: fewer than 16 rows per rank"); return EQGPU_EINVAL; }
        s->slab_fused = getenv("EQGPU_SLAB_UNFUSED") == nullptr;
        s->halo = s->slab_fused ? 6 : 1;
        if (!s->slab_fused) s->fused = false;
    }
    s->levels.clear();
    s->levels.push_back(l0);
    int maxl = p.max_levels > 0 ? p.max_levels : 12;
    if (const char *e = getenv("EQGPU_MAX_LEVELS")) maxl = std::max(1, atoi(e));   // tuning knob
    const double tau = p.dt * p.D;
    while ((int)s->levels.size() < maxl) {
        const Level &f = s->levels.back();
        if (std::min(f.dev.nx, f.dev.gny) < 5) break;
        if (s->slab && (f.dev.gny / 2) / s->slab_world < (s->slab_fused ? 12 : 4)) break;  // keep enough rows per rank
        // stop once the mass term dominates: Jacobi alone converges fast there
        if (tau / (f.hx_host[0] * f.hy_host[0]) < 0.6) break;
        Level c;
        c.hx_host = coarsen_cells(f.hx_host);
        c.hy_host = coarsen_cells(f.hy_host);
        c.dev.nx = (int)c.hx_host.size() + 1;
        c.dev.gny = (int)c.hy_host.size() + 1;
        // coarse row I belongs to whoever owns its coincident fine row min(2I, nf-1)
        c.g0 = c.dev.gny; c.g1 = 0;
        for (int I = 0; I < c.dev.gny; ++I) {
            const int fi = std::min(2 * I, f.dev.gny - 1);
            if (fi >= f.g0 && fi < f.g1) { c.g0 = std::min(c.g0, I); c.g1 = std::max(c.g1, I + 1); }
        }
        if (c.g1 <= c.g0) break;
        s->levels.push_back(c);
    }
    for (auto &lv : s->levels) {  // local window: owned rows plus one halo row towards each neighbour
        LevelDev &L = lv.dev;
        const int hb = lv.g0 > 0 ? s->halo : 0, ht = lv.g1 < L.gny ? s->halo : 0;
        if (s->slab && lv.g1 - lv.g0 < s->halo) { s->set_error("slab level thinner than its halo"); return EQGPU_EINVAL; }
        L.row0 = lv.g0 - hb;
        L.ny = (lv.g1 + ht) - L.row0;
        L.own0 = hb;
        L.own1 = L.ny - ht;
        L.slo = 0; L.shi = L.ny; L.wlo = 0; L.whi = L.ny; L.tbase = 0;
    }
    for (size_t l = 0; l < s->levels.size(); ++l) {
        Level &lv = s->levels[l];
        if (upload_padded(s, lv.hx_host, &lv.d_hx, &lv.d_ihx)) return EQGPU_ECUDA;
        if (upload_padded(s, lv.hy_host, &lv.d_hy, &lv.d_ihy, lv.dev.row0, lv.dev.ny)) return EQGPU_ECUDA;
        const size_t bytes = sizeof(double) * lv.n();
        EQ_CUDA(cudaMalloc(&lv.t, bytes));
        if (l > 0) {
            EQ_CUDA(cudaMalloc(&lv.x, bytes));
            EQ_CUDA(cudaMalloc(&lv.b, bytes));
        }
        if (s->slab_fused && upload_padded(s, lv.hy_host, &lv.d_hy_g, &lv.d_ihy_g)) return EQGPU_ECUDA;
        fill_level_consts(s, lv);
    }
    // ---- fine vectors ----------------------------------------------------
    s->N = s->levels[0].n();  // local nodes (owned rows + halo rows); == nW*nH on a single GPU
    const size_t bytes = sizeof(double) * s->N;
    EQ_CUDA(cudaMalloc(&s->u, bytes));
    EQ_CUDA(cudaMalloc(&s->r, bytes));
    EQ_CUDA(cudaMalloc(&s->pv, bytes));
    EQ_CUDA(cudaMalloc(&s->pv2, bytes));
    EQ_CUDA(cudaMemset(s->pv2, 0, bytes));
    EQ_CUDA(cudaMalloc(&s->Ap, bytes));
    EQ_CUDA(cudaMalloc(&s->z, bytes));
    EQ_CUDA(cudaMemset(s->u, 0, bytes));
    EQ_CUDA(cudaMemset(s->pv, 0, bytes));
    EQ_CUDA(cudaMemset(s->r, 0, bytes));
    EQ_CUDA(cudaMemset(s->Ap, 0, bytes));
    EQ_CUDA(cudaMemset(s->z, 0, bytes));
    for (auto &lv : s->levels) {
        EQ_CUDA(cudaMemset(lv.t, 0, sizeof(double) * lv.n()));
        if (&lv != &s->levels[0]) { EQ_CUDA(cudaMemset(lv.x, 0, sizeof(double) * lv.n())); EQ_CUDA(cudaMemset(lv.b, 0, sizeof(double) * lv.n())); }
    }
    s->levels[0].x = s->z;
    s->levels[0].b = s->r;
    // streaming smoothers: single GPU, isotropic operator; TMA descriptors where the pitch allows them
    // (opt-in, EQGPU_STREAM=1: measured slower than the tile kernels on the B200 -- 49-93 us against 40 us for the level-0
    // pre-smoother, profiles/r02_stream_smoothers.md -- and kept for the record of that experiment)
    s->stream_smooth = false;
    if (const char *e = getenv("EQGPU_STREAM")) s->stream_smooth = atoi(e) != 0 && !s->slab;
    s->stream_min_nodes = 0;
    if (const char *e = getenv("EQGPU_STREAM_MIN")) s->stream_min_nodes = atoi(e);
    s->stream_pipe = getenv("EQGPU_STREAM_PIPE") == nullptr || atoi(getenv("EQGPU_STREAM_PIPE")) != 0;
    s->stream_uni = getenv("EQGPU_STREAM_UNI") == nullptr || atoi(getenv("EQGPU_STREAM_UNI")) != 0;
    s->stream_apply = s->stream_smooth;
    if (const char *e = getenv("EQGPU_STREAM_APPLY")) s->stream_apply = atoi(e) != 0 && s->stream_smooth;
    s->map_pv_ptr = s->pv;
    s->tma_p = s->stream_apply && getenv("EQGPU_NO_TMA") == nullptr &&
               make_field_map(&s->map_z, s->z, s->levels[0].dev.nx, s->levels[0].dev.ny) &&
               make_field_map(&s->map_pv, s->pv, s->levels[0].dev.nx, s->levels[0].dev.ny) &&
               make_field_map(&s->map_pv2, s->pv2, s->levels[0].dev.nx, s->levels[0].dev.ny);
    for (size_t l = 0; l + 1 < s->levels.size(); ++l) {
        Level &lv = s->levels[l];
        lv.tma = s->stream_smooth && getenv("EQGPU_NO_TMA") == nullptr &&
                 make_field_map(&lv.map_b, lv.b, lv.dev.nx, lv.dev.ny) && make_field_map(&lv.map_t, lv.t, lv.dev.nx, lv.dev.ny);
    }
    // ---- reductions ------------------------------------------------------
    dim3 g0 = grid2d(s->levels[0].dev);
    s->max_blocks = std::max<int>(g0.x * g0.y, 8 * s->num_sms);
    EQ_CUDA(cudaMalloc(&s->partials, sizeof(double) * 16 * s->max_blocks));   // up to 9 sums per block (k_ls_gram)
    EQ_CUDA(cudaMalloc(&s->counters, sizeof(unsigned) * 16));
    EQ_CUDA(cudaMemset(s->counters, 0, sizeof(unsigned) * 16));
    EQ_CUDA(cudaMalloc(&s->sc, sizeof(CGScalars)));
    EQ_CUDA(cudaMemset(s->sc, 0, sizeof(CGScalars)));
    EQ_CUDA(cudaMallocHost(&s->sc_host, sizeof(CGScalars)));
    EQ_CUDA(cudaMalloc(&s->chan_top, sizeof(double) * p.nW));
    EQ_CUDA(cudaMalloc(&s->chan_bot, sizeof(double) * p.nW));
    EQ_CUDA(cudaMalloc(&s->flux_top, sizeof(double) * p.nW));
    EQ_CUDA(cudaMalloc(&s->flux_bot, sizeof(double) * p.nW));
    EQ_CUDA(cudaMemset(s->chan_top, 0, sizeof(double) * p.nW));
    EQ_CUDA(cudaMemset(s->chan_bot, 0, sizeof(double) * p.nW));
    EQ_CUDA(cudaMemset(s->flux_top, 0, sizeof(double) * p.nW));
    EQ_CUDA(cudaMemset(s->flux_bot, 0, sizeof(double) * p.nW));
    EQ_CUDA(cudaMalloc(&s->flux_dev, sizeof(double)));
    EQ_CUDA(cudaMallocHost(&s->flux_host, sizeof(double)));
    s->st.levels = (int)s->levels.size();
    for (int w = 0; w < 4; ++w) s->dir_val[w] = p.bc_value[w];
    // ---- single-CTA tail: the deepest levels that fit one CTA's shared memory
    EQ_CUDA(cudaMalloc(&s->d_levels, sizeof(LevelDev) * MAX_LEVELS));
    {
        const size_t budget = 200 * 1024;
        const int nl = (int)s->levels.size();
        int first = nl - 1;
        auto need = [&](int l) {
            const Level &lv = s->levels[l];
            return sizeof(double) * (3 * (size_t)(lv.dev.nx + 2) * (lv.dev.ny + 2) + 2 * (lv.dev.nx + 1) +
                                     2 * (lv.dev.ny + 1));
        };
        size_t used = need(nl - 1);
        while (first > 0 && used + need(first - 1) <= budget) {
            --first;
            used += need(first);
        }
        s->tail_first = first;
        s->tail_smem = used;
        // the tile kernel solves the coarsest level whatever its size when the coarse Chebyshev degree is
        // <= 8; otherwise the coarsest level has to fit the single-CTA tail, else the unfused path runs
        const int cn = coarse_weights(s).n;
        s->tail_fits = used <= budget;
        if (!(coarsest_tileable(s, cn) && nl >= 2) && !s->tail_fits) s->fused = false;
        if (s->slab_fused) {  // the tail kernels see one rank's rows only: slabs need the tiled coarsest solve
            if (coarsest_tileable(s, cn) && nl >= 2) s->tile_coarsest = true;
            else { s->slab_fused = false; s->fused = false; }
        }
        if (!s->tail_fits) { s->tail_smem = 0; s->tile_coarsest = true; }
    }
    if (s->fused) {
        // Cluster tail.  Measured at 2048^2 (profiles/r01_cluster_tail.md): the ~1.5 us cluster barrier per
        // sweep makes it lose against the tile kernels for levels >= 129^2, and win (388 vs 381 steps/s)
        // when it takes exactly the levels the single-CTA tail would take.  EQGPU_CTAIL_FIRST overrides.
        const int nl = (int)s->levels.size();
        int cf = s->tail_first;
        if (const char *e = getenv("EQGPU_CTAIL_FIRST")) {
            cf = std::max(0, std::min(atoi(e), nl - 1));
            while (cf < nl - 1 && (size_t)make_ctail_desc(s, cf, CT_MAX_CTAS).total * sizeof(double) > 200 * 1024) ++cf;
        }
        const CTailDesc td = make_ctail_desc(s, cf, CT_MAX_CTAS);
        const size_t csm = (size_t)td.total * sizeof(double);
        s->use_cluster = false;
        if (!s->slab && s->tail_fits && csm <= 200 * 1024 &&
            cudaFuncSetAttribute(k_ctail, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess &&
            cudaFuncSetAttribute(k_ctail, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)csm) == cudaSuccess) {
            cudaLaunchConfig_t cfg{};
            cfg.gridDim = dim3(CT_MAX_CTAS); cfg.blockDim = dim3(CT_THREADS); cfg.dynamicSmemBytes = csm;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = CT_MAX_CTAS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            int nclusters = 0;
            if (cudaOccupancyMaxActiveClusters(&nclusters, k_ctail, &cfg) == cudaSuccess && nclusters >= 1) {
                s->use_cluster = true;
                s->ctail_first = cf; s->ctail_ncta = CT_MAX_CTAS; s->ctail_smem = csm;
            }
        }
        cudaGetLastError();  // a refused attribute is not an error of the solver
        if (s->tail_fits)
            EQ_CUDA(cudaFuncSetAttribute(k_tail, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s->tail_smem));
        const int nu = s->nu;
        if (nu > 4) { s->set_error("smooth_sweeps must be <= 4"); return EQGPU_EINVAL; }
        // 64-node tiles need more than the default 48 KB of dynamic shared memory (32-node tiles: 27.7 KB)
        const int tsm = (int)SM64_2, tsm3 = (int)SM64_3;
#define SET_SMEM(NU)                                                                                         \
    case NU:                                                                                                 \
        EQ_CUDA(cudaFuncSetAttribute((T64::k_presmooth<NU, 8>), cudaFuncAttributeMaxDynamicSharedMemorySize, tsm));       \
        EQ_CUDA(cudaFuncSetAttribute((T64::k_presmooth<NU, 4>), cudaFuncAttributeMaxDynamicSharedMemorySize, tsm3));       \
        EQ_CUDA(cudaFuncSetAttribute((T64::k_postsmooth<NU, true, 8>), cudaFuncAttributeMaxDynamicSharedMemorySize, tsm)); \
        EQ_CUDA(cudaFuncSetAttribute((T64::k_postsmooth<NU, false, 8>), cudaFuncAttributeMaxDynamicSharedMemorySize, tsm)); \
        EQ_CUDA(cudaFuncSetAttribute((T64::k_postsmooth<NU, true, 4>), cudaFuncAttributeMaxDynamicSharedMemorySize, tsm3)); \
        EQ_CUDA(cudaFuncSetAttribute((T64::k_postsmooth<NU, false, 4>), cudaFuncAttributeMaxDynamicSharedMemorySize, tsm3)); \
        break;
        for (int q = 1; q <= 4; ++q) switch (q) { SET_SMEM(1) SET_SMEM(2) SET_SMEM(3) SET_SMEM(4) }
        // tile-list instances beside the register-tile kernels (even halo: one extra node where the sweep count asks for an odd one)
        EQ_CUDA(cudaFuncSetAttribute((T64::k_presmooth<4, 8, 1>), cudaFuncAttributeMaxDynamicSharedMemorySize, tsm));
        EQ_CUDA(cudaFuncSetAttribute((T64::k_postsmooth<3, true, 8, 1>), cudaFuncAttributeMaxDynamicSharedMemorySize, tsm));
        EQ_CUDA(cudaFuncSetAttribute((T64::k_postsmooth<3, false, 8, 1>), cudaFuncAttributeMaxDynamicSharedMemorySize, tsm));
        EQ_CUDA(cudaFuncSetAttribute((T64::k_coarsest<4>), cudaFuncAttributeMaxDynamicSharedMemorySize, tsm3));
        if (const char *e = getenv("EQGPU_TILE_COARSEST")) s->tile_coarsest = atoi(e) != 0 || !s->tail_fits || s->slab;
        // Levels whose 64-node tiling gives fewer than t32_below CTAs are latency-bound (one big tile per
        // SM, most SMs idle): they run on 32-node tiles, 256 threads, several CTAs per SM.
        if (const char *e = getenv("EQGPU_T32_BELOW")) s->t32_below = atoi(e);
        if (const char *e = getenv("EQGPU_T32_WIDE")) s->t32_wide = atoi(e) != 0;
        if (const char *e = getenv("EQGPU_INIT_TILE")) s->init_tile = atoi(e) != 0;
        // (row slabs: the same split, on the owned rows; EQGPU_SLAB_DEFER_X=0 falls back to the combined k_update_xr)
        s->defer_x = !s->slab || s->slab_fused;
        if (const char *e = getenv("EQGPU_DEFER_X")) s->defer_x = atoi(e) != 0 && !s->slab;
        if (s->slab) {
            if (const char *e = getenv("EQGPU_SLAB_DEFER_X")) s->defer_x = atoi(e) != 0 && s->slab_fused;
            // (k_update_r / k_update_x use 128-bit accesses from the first owned node on)
            const LevelDev &L0 = s->levels[0].dev;
            if (((size_t)L0.own0 * L0.nx) & 1) s->defer_x = false;
        }
        // Least-squares combination on top of the extrapolations: measured (profiles/r01_ls_guess.md) 3.1 vs 5.3
        // iterations per step at 257^2 and 4.5 vs 5.1 at 512^2 over 40 steps of the bench colony, but 4.9 vs 4.5
        // at 2048^2, where the field is far from steady and PCG converges more slowly from the residual-optimal
        // start: on by default up to 512^2 nodes
        // ... and the cubic extrapolation of the last four solutions above
        s->warm = (size_t)p.nW * p.nH <= (size_t)512 * 512 ? 4 : 6;
        if (const char *e = getenv("EQGPU_WARM")) s->warm = std::max(0, std::min(atoi(e), 7));
        if (const char *e = getenv("EQGPU_LS_FORM")) s->ls_form = atoi(e) != 0 ? 1 : 0;   // tuning knob
        if (const char *e = getenv("EQGPU_WARM_ADAPTIVE")) s->warm_adaptive = atoi(e) != 0;
        if ((s->defer_x || s->slab) && s->init_tile && s->warm > 0) {
            for (int k = 0; k < 5; ++k) {
                EQ_CUDA(cudaMalloc(&s->uh[k], sizeof(double) * s->N));
                EQ_CUDA(cudaMemset(s->uh[k], 0, sizeof(double) * s->N));
            }
            if (!s->slab)
                for (int k = 0; k < 2; ++k) {
                    EQ_CUDA(cudaMalloc(&s->dk[k], sizeof(double) * s->N));
                    EQ_CUDA(cudaMemset(s->dk[k], 0, sizeof(double) * s->N));
                }
            s->dk_valid = false;
            s->hist = 0;
        }
        if (s->defer_x) {
            int prio_lo = 0, prio_hi = 0;   // lowest priority: the critical path's CTAs are scheduled first
            cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
            EQ_CUDA(cudaStreamCreateWithPriority(&s->side_stream, cudaStreamNonBlocking, prio_lo));
            s->xupd_blocks = std::max(1, s->num_sms / 2);   // measured: 74 CTAs 522 steps/s, 148: 518, 296: 512, 592: 511
            if (const char *e = getenv("EQGPU_XUPD_BLOCKS")) s->xupd_blocks = std::max(1, atoi(e));
            s->join_pdl = true;   // k_apply_p keeps its programmatic edge from k_postsmooth beside the full edge from k_update_x
            if (const char *e = getenv("EQGPU_JOIN_PDL")) s->join_pdl = atoi(e) != 0;
            EQ_CUDA(cudaEventCreateWithFlags(&s->ev_fork, cudaEventDisableTiming));
            EQ_CUDA(cudaEventCreateWithFlags(&s->ev_join, cudaEventDisableTiming));
        }
        s->pdl = !s->slab;   // slab mode has NCCL calls between the kernels
        if (const char *e = getenv("EQGPU_PDL")) s->pdl = atoi(e) != 0 && !s->slab;
        {
            int rc = rt_setup(s);
            if (rc) return rc;
        }

#undef SET_SMEM
    }
    return 0;
}

void solver_teardown(eqgpu_solver *s)
{
    rt_teardown(s);
    for (auto &lv : s->levels) {
        cudaFree(lv.d_hx); cudaFree(lv.d_ihx); cudaFree(lv.d_hy); cudaFree(lv.d_ihy);
        cudaFree(lv.d_hy_g); cudaFree(lv.d_ihy_g);
        cudaFree(lv.t);
        if (&lv != &s->levels[0]) { cudaFree(lv.x); cudaFree(lv.b); }
        if (&lv != &s->levels[0]) { cudaFree(lv.t11); cudaFree(lv.t22); cudaFree(lv.t12); }
        cudaFree(lv.kC); cudaFree(lv.kE); cudaFree(lv.kN); cudaFree(lv.kNE);
    }
    s->levels.clear();
    if (s->graph_exec) { cudaGraphExecDestroy(s->graph_exec); s->graph_exec = nullptr; }
    if (s->graph_exec2) { cudaGraphExecDestroy(s->graph_exec2); s->graph_exec2 = nullptr; }
    for (int k = 0; k < 5; ++k) { cudaFree(s->uh[k]); s->uh[k] = nullptr; }
    for (int k = 0; k < 2; ++k) { cudaFree(s->dk[k]); s->dk[k] = nullptr; }
    for (int k = 0; k < 7; ++k) { cudaFree(s->ring_h[k]); s->ring_h[k] = nullptr; cudaFree(s->ring_a[k]); s->ring_a[k] = nullptr; }
    cudaFree(s->ring_b); s->ring_b = nullptr;
    cudaFree(s->ring_partials); s->ring_partials = nullptr;
    s->ring_n = 0;
    s->hist = 0;
    if (s->ev_fork) { cudaEventDestroy(s->ev_fork); s->ev_fork = nullptr; }
    if (s->ev_join) { cudaEventDestroy(s->ev_join); s->ev_join = nullptr; }
    if (s->side_stream) { cudaStreamDestroy(s->side_stream); s->side_stream = nullptr; }
    cudaFree(s->d_levels); cudaFree(s->pv2);
    cudaFree(s->u); cudaFree(s->r); cudaFree(s->pv); cudaFree(s->Ap); cudaFree(s->z);
    cudaFree(s->d11); cudaFree(s->d22); cudaFree(s->d12);
    cudaFree(s->partials); cudaFree(s->counters); cudaFree(s->sc);
    cudaFree(s->chan_top); cudaFree(s->chan_bot); cudaFree(s->flux_top); cudaFree(s->flux_bot);
    cudaFree(s->chan_coef); cudaFree(s->flux_dev);
    if (s->sc_host) cudaFreeHost(s->sc_host);
    if (s->flux_host) cudaFreeHost(s->flux_host);
}

// (re)build the coarse tensor fields after eqgpu_set_tensor
int solver_refresh_levels(eqgpu_solver *s)
{
    Level &l0 = s->levels[0];
    l0.t11 = s->d11; l0.t22 = s->d22; l0.t12 = s->d12;
    for (size_t l = 0; l < s->levels.size(); ++l) {
        Level &lv = s->levels[l];
        if (l > 0) {
            if (s->tensor) {
                const size_t bytes = sizeof(double) * lv.n();
                if (!lv.t11) {
                    EQ_CUDA(cudaMalloc(&lv.t11, bytes));
                    EQ_CUDA(cudaMalloc(&lv.t22, bytes));
                    EQ_CUDA(cudaMalloc(&lv.t12, bytes));
                }
                Level &f = s->levels[l - 1];
                fill_level_consts(s, lv);
                k_inject<<<grid2d(lv.dev), dim3(BX, BY), 0, s->stream>>>(f.dev, lv.dev, f.t11, lv.t11);
                k_inject<<<grid2d(lv.dev), dim3(BX, BY), 0, s->stream>>>(f.dev, lv.dev, f.t22, lv.t22);
                k_inject<<<grid2d(lv.dev), dim3(BX, BY), 0, s->stream>>>(f.dev, lv.dev, f.t12, lv.t12);
                s->launches += 3;
            }
        }
        fill_level_consts(s, lv);
        // one GPU, consistent mass: assemble the rows once per tensor update; every kernel of the variable-tensor path then
        // reads four coefficients per node (EQGPU_TENSOR_ASSEMBLE=0: evaluate them on the fly, as round 1 did)
        static const bool assemble = getenv("EQGPU_TENSOR_ASSEMBLE") == nullptr || atoi(getenv("EQGPU_TENSOR_ASSEMBLE")) != 0;
        if (s->tensor && !s->slab && assemble && lv.t11) {
            const size_t bytes = sizeof(double) * lv.n();
            if (!lv.kC) {
                EQ_CUDA(cudaMalloc(&lv.kC, bytes));
                EQ_CUDA(cudaMalloc(&lv.kE, bytes));
                EQ_CUDA(cudaMalloc(&lv.kN, bytes));
                EQ_CUDA(cudaMalloc(&lv.kNE, bytes));
            }
            k_coef_assemble<<<grid2d(lv.dev), dim3(BX, BY), 0, s->stream>>>(lv.dev, lv.kC, lv.kE, lv.kN, lv.kNE);
            s->launches++;
            lv.dev.kC = lv.kC; lv.dev.kE = lv.kE; lv.dev.kN = lv.kN; lv.dev.kNE = lv.kNE;
            lv.gdev.kC = lv.kC; lv.gdev.kE = lv.kE; lv.gdev.kN = lv.kN; lv.gdev.kNE = lv.kNE;
        }
    }
    {
        std::vector<LevelDev> host(MAX_LEVELS);
        for (size_t l = 0; l < s->levels.size() && l < MAX_LEVELS; ++l) host[l] = s->levels[l].dev;
        EQ_CUDA(cudaMemcpyAsync(s->d_levels, host.data(), sizeof(LevelDev) * MAX_LEVELS, cudaMemcpyHostToDevice,
                                s->stream));
        EQ_CUDA(cudaStreamSynchronize(s->stream));
    }
    EQ_CUDA(cudaGetLastError());
    return 0;
}

static DirData make_dirdata(eqgpu_solver *s)
{
    DirData d;
    for (int w = 0; w < 4; ++w) d.val[w] = s->dir_val[w];
    d.top = s->p.bc_type[EQGPU_TOP] == EQGPU_BC_DIRICHLET_CHANNEL ? s->chan_top : nullptr;
    d.bot = s->p.bc_type[EQGPU_BOTTOM] == EQGPU_BC_DIRICHLET_CHANNEL ? s->chan_bot : nullptr;
    return d;
}

static SmoothW smooth_weights(eqgpu_solver *s);
static CoarseW coarse_weights(eqgpu_solver *s);

// Unfused V-cycle: one kernel per sweep / transfer.  Serves the variable-tensor operator and the
// row-slab mode, where a one-row halo exchange (slab_exchange) precedes every kernel that reads
// neighbours.  Same Chebyshev-weighted smoother and coarse solve as the fused cycle.
template <bool T>
static void vcycle(eqgpu_solver *s)
{
    const dim3 blk(BX, BY);
    cudaStream_t st = s->stream;
    const int nl = (int)s->levels.size();
    const SmoothW sw = smooth_weights(s);
    const CoarseW cw = coarse_weights(s);
    auto xch = [&](Level &lv, double *v) { if (s->slab) slab_exchange(s, lv.dev, v); };
    // sweeps from a zero guess, result left in lv.x (uses lv.t as ping-pong)
    auto smooth_from_zero = [&](Level &lv, int sweeps, const double *w) {
        const dim3 g = grid2d(lv.dev);
        double *cur = (sweeps & 1) ? lv.x : lv.t;  // so that the last write lands in lv.x
        if (T && lv.dev.kC) k_jacobi0<2><<<g, blk, 0, st>>>(lv.dev, lv.b, cur, w[0], s->sc);
        else k_jacobi0<T><<<g, blk, 0, st>>>(lv.dev, lv.b, cur, w[0], s->sc);
        s->launches++;
        for (int k = 1; k < sweeps; ++k) {
            double *nxt = (cur == lv.x) ? lv.t : lv.x;
            xch(lv, cur);
            if (T && lv.dev.kC) k_jacobi<2, false><<<g, blk, 0, st>>>(lv.dev, lv.b, cur, nxt, w[k], s->sc);
            else k_jacobi<T, false><<<g, blk, 0, st>>>(lv.dev, lv.b, cur, nxt, w[k], s->sc);
            s->launches++;
            cur = nxt;
        }
    };
    auto smooth = [&](Level &lv, int sweeps, const double *w) {  // in: lv.x, out: lv.x
        const dim3 g = grid2d(lv.dev);
        double *cur = lv.x;
        if (sweeps & 1) {  // odd count: start from a copy in lv.t so the last write lands in lv.x
            cudaMemcpyAsync(lv.t, lv.x, sizeof(double) * lv.n(), cudaMemcpyDeviceToDevice, st);
            cur = lv.t;
        }
        for (int k = 0; k < sweeps; ++k) {
            double *nxt = (cur == lv.x) ? lv.t : lv.x;
            xch(lv, cur);
            if (T && lv.dev.kC) k_jacobi<2, false><<<g, blk, 0, st>>>(lv.dev, lv.b, cur, nxt, w[k], s->sc);
            else k_jacobi<T, false><<<g, blk, 0, st>>>(lv.dev, lv.b, cur, nxt, w[k], s->sc);
            s->launches++;
            cur = nxt;
        }
    };
    for (int l = 0; l < nl - 1; ++l) {
        Level &lv = s->levels[l], &cv = s->levels[l + 1];
        smooth_from_zero(lv, s->nu, sw.w);
        xch(lv, lv.x);
        if (T && lv.dev.kC) k_jacobi<2, true><<<grid2d(lv.dev), blk, 0, st>>>(lv.dev, lv.b, lv.x, lv.t, 0.0, s->sc);
        else k_jacobi<T, true><<<grid2d(lv.dev), blk, 0, st>>>(lv.dev, lv.b, lv.x, lv.t, 0.0, s->sc);
        xch(lv, lv.t);
        k_restrict<<<grid2d(cv.dev), blk, 0, st>>>(lv.dev, cv.dev, lv.t, cv.b, s->sc);
        s->launches += 2;
    }
    smooth_from_zero(s->levels[nl - 1], cw.n, cw.w);
    for (int l = nl - 2; l >= 0; --l) {
        Level &lv = s->levels[l], &cv = s->levels[l + 1];
        xch(cv, cv.x);
        k_prolong_add<<<grid2d(lv.dev), blk, 0, st>>>(lv.dev, cv.dev, cv.x, lv.x, s->sc);
        s->launches++;
        smooth(lv, s->nu, sw.w);
    }
}

static CTailDesc make_ctail_desc(eqgpu_solver *s, int first, int ncta)
{
    CTailDesc td{};
    const int nl = (int)s->levels.size();
    td.first = first; td.last = nl - 1; td.ncta = ncta;
    int off = 0, maxnx = 0;
    for (int l = first; l < nl; ++l) {
        const LevelDev &L = s->levels[l].dev;
        td.rp[l] = (L.ny + ncta - 1) / ncta;
        td.off[l] = off;
        off += 3 * td.rp[l] * (L.nx + 2);
        maxnx = std::max(maxnx, L.nx);
    }
    for (int l = first; l < nl; ++l) {
        const LevelDev &L = s->levels[l].dev;
        td.soff[l] = off;
        off += 2 * (L.nx + 1) + 2 * (L.ny + 1);
    }
    td.zoff = off;
    off += maxnx + 2;
    td.total = off;
    return td;
}

// Chebyshev-root Jacobi weights: n sweeps x <- x + w_k D^-1 (b - A x) whose
// error polynomial is the scaled Chebyshev polynomial on [lo, hi] (eigenvalue
// range of D^-1 A to damp).  Roots are taken alternately from both ends.
static void cheb_weights(int n, double lo, double hi, double *w)
{
    const double theta = 0.5 * (hi + lo), delta = 0.5 * (hi - lo);
    std::vector<double> r(n);
    for (int k = 0; k < n; ++k) r[k] = 1.0 / (theta - delta * std::cos(M_PI * (2 * k + 1) / (2.0 * n)));
    for (int k = 0, a = 0, b = n - 1; k < n; ++k) w[k] = (k & 1) ? r[b--] : r[a++];
}

static SmoothW smooth_weights(eqgpu_solver *s)
{
    SmoothW sw{};
    // high-frequency range of D^-1 A on this stencil is [0.5, 2] (DESIGN.md "smoother")
    cheb_weights(s->nu, 0.5, 2.0, sw.w);
    return sw;
}

static CoarseW coarse_weights(eqgpu_solver *s)
{
    CoarseW cw{};
    const Level &c = s->levels.back();
    const double F = s->p.dt * s->p.D / (c.hx_host[0] * c.hy_host[0]);
    // smallest eigenvalue of D^-1 A >= (mass row sum)/(diagonal) = 1/(0.5 + 4F) on square cells
    // (1/(1 + 4F) with the lumped mass of the finite-difference discretisation)
    const double lo = 0.8 / ((s->p.discretisation == EQGPU_DISC_FD ? 1.0 : 0.5) + 4.0 * F), hi = 2.0;
    const double sigma = (hi + lo) / (hi - lo);
    double target = 50.0;   // worst-case error reduction of the coarsest solve over [lo, hi]
    if (const char *e = getenv("EQGPU_COARSE_TARGET")) target = std::max(2.0, atof(e));   // tuning knob
    int n = (int)std::ceil(std::acosh(target) / std::acosh(sigma));
    n = std::max(2, std::min(n, MAX_CHEB));
    cw.n = n;
    cheb_weights(n, lo, hi, cw.w);
    return cw;
}

static std::vector<cudaEvent_t> *g_trace = nullptr;  // debugging aid (EQGPU_TRACE): event after each launch
static std::vector<const char *> g_trace_labels;
void solver_trace_mark(cudaStream_t st, const char *label)
{
    if (!g_trace) return;
    cudaEvent_t e;
    cudaEventCreate(&e);
    cudaEventRecord(e, st);
    g_trace->push_back(e);
    g_trace_labels.push_back(label);
}
static void trace_mark(cudaStream_t st, const char *label = "kernel") { solver_trace_mark(st, label); }

static SmoothW smooth_weights_n(int n)
{
    SmoothW sw{};
    double lo = 0.25, hi = 2.0;  // smoothing interval of D^-1 A: [0.5,2] is the high-frequency range; 0.25 measured best
    if (const char *e = getenv("EQGPU_CHEB_LO")) lo = atof(e);  // tuning knobs
    if (const char *e = getenv("EQGPU_CHEB_HI")) hi = atof(e);
    cheb_weights(n, lo, hi, sw.w);
    return sw;
}

// The LevelDev the tile kernels see (global row indices) and the matching pre-offset pointers.
static inline const LevelDev &TV(const eqgpu_solver *s, const Level &lv) { return s->slab_fused ? lv.gdev : lv.dev; }
template <class P>
static inline P *VP(const eqgpu_solver *s, const Level &lv, P *ptr)
{
    return s->slab_fused ? ptr - (ptrdiff_t)lv.dev.row0 * lv.dev.nx : ptr;
}
static inline dim3 tile_grid(const LevelDev &L, int to)
{
    return dim3((L.nx + to - 1) / to, (L.whi - L.tbase + to - 1) / to);
}
static inline void xch(eqgpu_solver *s, Level &lv, double *v, int depth)
{
    if (s->slab) slab_exchange(s, lv.dev, v, depth);
}

// Launch on `st`; with s->pdl and PDL_OK the launch carries the programmatic-serialization attribute, so
// the kernel's prologue overlaps the tail of its predecessor (every such kernel calls pdl_wait() before it
// touches data).  The first kernel of an iteration is launched plainly: it has no predecessor in the graph.
#define LAUNCH_K(PDL_OK, KERN, G, B, SM, ST, ...)                                                         \
    do {                                                                                                  \
        cudaLaunchConfig_t cfg_{};                                                                        \
        cfg_.gridDim = (G); cfg_.blockDim = (B); cfg_.dynamicSmemBytes = (SM); cfg_.stream = (ST);        \
        cudaLaunchAttribute at_[1];                                                                       \
        if (s->pdl && (PDL_OK) && !s->pdl_block) {                                                        \
            at_[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;                               \
            at_[0].val.programmaticStreamSerializationAllowed = 1;                                        \
            cfg_.attrs = at_; cfg_.numAttrs = 1;                                                          \
        }                                                                                                 \
        s->pdl_block = false;   /* (set by a launch that ends in a stream join) */                        \
        cudaLaunchKernelEx(&cfg_, KERN, __VA_ARGS__);                                                     \
    } while (0)

// Tile edge for a level: 32-node tiles (256 threads, ~28 KB) when the 64-node tiling would leave the GPU
// mostly idle -- those levels are bound by the latency of one tile, not by bandwidth.
static inline bool use_t32(const eqgpu_solver *s, const LevelDev &L, int to64)
{
    const dim3 g = tile_grid(L, to64);
    return (int)(g.x * g.y) < s->t32_below;
}

// ---- streaming smoothers (mg_stream.cuh): host side -------------------------------------------------------------------
typedef CUresult (*PFN_tensorMapEncodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                             const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                             CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// TMA descriptor of an nx x ny fp64 field (row-major, pitch nx) with a box of RB rows x 64 columns; false when the
// driver entry point is missing or the pitch is not a multiple of 16 bytes (odd nx): the caller stages with cp.async.
static bool make_field_map(CUtensorMap *map, double *base, int nx, int ny)
{
    static PFN_tensorMapEncodeTiled enc = nullptr;
    static bool looked = false;
    if (!looked) {
        looked = true;
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            enc = (PFN_tensorMapEncodeTiled)fn;
        (void)cudaGetLastError();
    }
    if (!enc || (nx & 1) || (((size_t)base) & 15)) return false;
    const cuuint64_t gdim[2] = {(cuuint64_t)nx, (cuuint64_t)ny};
    const cuuint64_t gstride[1] = {(cuuint64_t)nx * sizeof(double)};
    const cuuint32_t box[2] = {(cuuint32_t)STRM::SWID, (cuuint32_t)STRM::RB};
    const cuuint32_t estr[2] = {1, 1};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, base, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static bool use_stream(const eqgpu_solver *s, const Level &lv)
{
    // (the only irregular rows the kernels tabulate are 0, ny-2 and ny-1: true for the uniform meshes and their
    // coarsenings this library builds)
    return s->stream_smooth && !s->slab && (long long)lv.dev.nx * lv.dev.ny >= s->stream_min_nodes && lv.dev.nx >= 8 &&
           lv.dev.ny >= 8 && lv.dev.ireg_hi >= lv.dev.ny - 3;
}

// Every in-grid column of the level is regular or Dirichlet: the UNI kernel instances apply (coefficients as constant-bank
// operands).  Natural-boundary wall columns and the narrower last cell of a coarse grid need per-lane sets.
static bool uniform_columns(const LevelDev &F)
{
    return (F.dirmask & 1u) && (F.dirmask & 2u) && F.jreg_hi >= F.nx - 2 && F.ireg_hi >= 1;
}

template <int NU>
static STRM::UniCoef<NU> uni_coef(const LevelDev &F, const SmoothW &sw)
{
    STRM::UniCoef<NU> U;
    U.nC = -F.cC; U.nE = -F.cEW; U.nW = -F.cEW; U.nN = -F.cNS; U.nS = -F.cNS; U.nNE = -F.cD; U.nSW = -F.cD;
    for (int k = 0; k < NU; ++k) U.wic[k] = sw.w[k] * F.icC;
    return U;
}

// Resident CTAs of a streaming kernel on the whole device (registers decide: 64-thread CTAs of 100-250 registers).
template <class K>
static int stream_slots(const eqgpu_solver *s, K kernel, size_t smem, int threads = 32 * STRM::WPC)
{
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel, threads, smem) != cudaSuccess || nb < 1) nb = 4;
    (void)cudaGetLastError();
    return nb * s->num_sms;
}

// Task shape: strips of 64 columns, chunks of hs owned rows.  `lag` = rows the last sweep trails the entering row plus
// what the output stage needs (pre: 2NU + 1, post: 2NU, apply: 2).  hs is chosen so that the grid is a whole number of
// waves of resident CTAs (a 1.03-wave grid takes two task durations) at the smallest walked-rows x waves product.
static STRM::StreamGeom stream_geom(const eqgpu_solver *s, const LevelDev &F, int halo, int lag, int slots,
                                    int wpc = STRM::WPC, int fill = 14)
{
    STRM::StreamGeom G;
    G.halo = halo;
    const int so = STRM::SWID - 2 * G.halo;
    G.nstrips = (F.nx + so - 1) / so;
    const int cx = (G.nstrips + wpc - 1) / wpc;
    long long best = -1;
    G.hs = 16;
    for (int hs = 12; hs <= 192; hs += 2) {
        const int nch = (F.ny + hs - 1) / hs;
        const long long waves = ((long long)cx * nch + slots - 1) / slots;
        const int steps = ((hs + halo + lag + STRM::RB - 1) / STRM::RB) * STRM::RB;
        const long long cost = waves * (steps + fill);   // + set-up, pipeline fill and drain of a task, in walk steps
        if (best < 0 || cost <= best) { best = cost; G.hs = hs; }
    }
    if (const char *e = getenv("EQGPU_STREAM_HS")) G.hs = std::max(2, atoi(e) & ~1);   // tuning knob
    G.nchunks = (F.ny + G.hs - 1) / G.hs;
    G.nblk = (G.hs + G.halo + lag + STRM::RB - 1) / STRM::RB;
    return G;
}

template <int NU>
static void launch_pre_stream(eqgpu_solver *s, cudaStream_t st, int l, bool pdl_ok)
{
    Level &lv = s->levels[l], &cv = s->levels[l + 1];
    const SmoothW sw = smooth_weights_n(NU);
    const size_t smem = STRM::WPC * STRM::NSLOT * STRM::BLK_BYTES + STRM::WPC * STRM::NSLOT * sizeof(unsigned long long);
    const CGScalars *scc = s->sc;
    const STRM::UniCoef<NU> U = uni_coef<NU>(lv.dev, sw);
    const int halo = (NU + 2) & ~1;   // even, >= NU + 1: the restriction reads one residual further
#define SPRE(TMA, UNI)                                                                                               \
    do {                                                                                                             \
        const STRM::StreamGeom G = stream_geom(s, lv.dev, halo, 2 * NU + 1,                                          \
                                               stream_slots(s, STRM::ks_presmooth<NU, TMA, UNI>, smem));             \
        const dim3 grid((G.nstrips + STRM::WPC - 1) / STRM::WPC, G.nchunks);                                         \
        LAUNCH_K(pdl_ok, (STRM::ks_presmooth<NU, TMA, UNI>), grid, dim3(32 * STRM::WPC), smem, st, lv.dev, cv.dev,   \
                 lv.map_b, (const double *)lv.b, lv.t, cv.b, sw, U, G, scc);                                         \
    } while (0)
    const bool uni = s->stream_uni && uniform_columns(lv.dev);
    if (lv.tma) { if (uni) SPRE(true, true); else SPRE(true, false); }
    else { if (uni) SPRE(false, true); else SPRE(false, false); }
#undef SPRE
}

template <int NU>
static void launch_post_stream(eqgpu_solver *s, cudaStream_t st, int l)
{
    Level &lv = s->levels[l], &cv = s->levels[l + 1];
    const SmoothW sw = smooth_weights_n(NU);
    const size_t smem = 2 * STRM::WPC * STRM::NSLOT * STRM::BLK_BYTES + STRM::WPC * STRM::NSLOT * STRM::CBLK_BYTES +
                        STRM::WPC * STRM::NSLOT * sizeof(unsigned long long);
    double *out_dot = &s->sc->rz_new;
    const STRM::UniCoef<NU> U = uni_coef<NU>(lv.dev, sw);
    const int halo = (NU + 1) & ~1;   // even, >= NU
#define SPOST(DOT, TMA, UNI)                                                                                          \
    do {                                                                                                              \
        const STRM::StreamGeom G = stream_geom(s, lv.dev, halo, 2 * NU,                                               \
                                               stream_slots(s, STRM::ks_postsmooth<NU, DOT, TMA, UNI>, smem));        \
        const dim3 grid((G.nstrips + STRM::WPC - 1) / STRM::WPC, G.nchunks);                                          \
        LAUNCH_K(true, (STRM::ks_postsmooth<NU, DOT, TMA, UNI>), grid, dim3(32 * STRM::WPC), smem, st, lv.dev, cv.dev, \
                 lv.map_b, lv.map_t, (const double *)lv.b, (const double *)lv.t, lv.x, (const double *)cv.x, sw, U, G, \
                 s->sc, s->partials, s->counters + 1, out_dot);                                                       \
    } while (0)
    const bool uni = s->stream_uni && uniform_columns(lv.dev);
    if (l == 0) {
        if (lv.tma) { if (uni) SPOST(true, true, true); else SPOST(true, true, false); }
        else { if (uni) SPOST(true, false, true); else SPOST(true, false, false); }
    } else {
        if (lv.tma) { if (uni) SPOST(false, true, true); else SPOST(false, true, false); }
        else { if (uni) SPOST(false, false, true); else SPOST(false, false, false); }
    }
#undef SPOST
}

// p' = z + beta p, Ap', p'.Ap' on level 0 (pin -> pout)
static void launch_apply_stream(eqgpu_solver *s, cudaStream_t st, bool pdl_ok, double *pin, double *pout)
{
    Level &l0 = s->levels[0];
    const LevelDev &F = l0.dev;
    const size_t smem = 2 * STRM::WPC * STRM::NSLOT * STRM::BLK_BYTES + STRM::WPC * STRM::NSLOT * sizeof(unsigned long long);
    const CUtensorMap &mp = pin == s->map_pv_ptr ? s->map_pv : s->map_pv2;   // the descriptors belong to buffers, and pv / pv2 swap
    SmoothW sw1{};
    const STRM::UniCoef<1> U = uni_coef<1>(F, sw1);
#define SAPPLY(TMA, UNI)                                                                                              \
    do {                                                                                                              \
        const STRM::StreamGeom G = stream_geom(s, F, 2, 2, stream_slots(s, STRM::ks_apply_p<TMA, UNI>, smem));        \
        const dim3 grid((G.nstrips + STRM::WPC - 1) / STRM::WPC, G.nchunks);                                          \
        LAUNCH_K(pdl_ok, (STRM::ks_apply_p<TMA, UNI>), grid, dim3(32 * STRM::WPC), smem, st, F, s->map_z, mp,         \
                 (const double *)s->z, (const double *)pin, pout, s->Ap, U, G, s->sc, s->partials, s->counters + 2,   \
                 &s->sc->pAp);                                                                                        \
    } while (0)
    const bool uni = s->stream_uni && uniform_columns(F);
    if (s->tma_p) { if (uni) SAPPLY(true, true); else SAPPLY(true, false); }
    else { if (uni) SAPPLY(false, true); else SAPPLY(false, false); }
#undef SAPPLY
}

// warp-specialised sweep pipelines (PIPE::kp_*): one CTA of NU warps per task, TMA-staged levels only
template <int NU>
static void launch_pre_pipe(eqgpu_solver *s, cudaStream_t st, int l, bool pdl_ok)
{
    Level &lv = s->levels[l], &cv = s->levels[l + 1];
    const SmoothW sw = smooth_weights_n(NU);
    const size_t smem = PIPE::Layout<NU, false>::BYTES;
    static bool attr = false;   // per process and device-independent here: raising the limit is idempotent
    if (!attr || true) cudaFuncSetAttribute(PIPE::kp_presmooth<NU>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr = true;
    const CGScalars *scc = s->sc;
    const STRM::StreamGeom G = stream_geom(s, lv.dev, (NU + 2) & ~1, 2 * NU + 1,
                                           stream_slots(s, PIPE::kp_presmooth<NU>, smem, 32 * NU), 1, 12 + 6 * NU);
    LAUNCH_K(pdl_ok, (PIPE::kp_presmooth<NU>), dim3(G.nstrips, G.nchunks), dim3(32 * NU), smem, st, lv.dev, cv.dev, lv.map_b,
             lv.t, cv.b, sw, G, scc);
}

template <int NU>
static void launch_post_pipe(eqgpu_solver *s, cudaStream_t st, int l)
{
    Level &lv = s->levels[l], &cv = s->levels[l + 1];
    const SmoothW sw = smooth_weights_n(NU);
    const size_t smem = PIPE::Layout<NU, true>::BYTES;
    double *out_dot = &s->sc->rz_new;
#define PPOST(DOT)                                                                                                  \
    do {                                                                                                            \
        cudaFuncSetAttribute(PIPE::kp_postsmooth<NU, DOT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        const STRM::StreamGeom G = stream_geom(s, lv.dev, (NU + 1) & ~1, 2 * NU,                                    \
                                               stream_slots(s, PIPE::kp_postsmooth<NU, DOT>, smem, 32 * NU), 1,     \
                                               12 + 6 * NU);                                                        \
        LAUNCH_K(true, (PIPE::kp_postsmooth<NU, DOT>), dim3(G.nstrips, G.nchunks), dim3(32 * NU), smem, st, lv.dev, \
                 cv.dev, lv.map_b, lv.map_t, lv.x, (const double *)cv.x, sw, G, s->sc, s->partials,                 \
                 s->counters + 1, out_dot);                                                                         \
    } while (0)
    if (l == 0) PPOST(true); else PPOST(false);
#undef PPOST
}

template <int NU>
static void launch_pre(eqgpu_solver *s, cudaStream_t st, int l)
{
    Level &lv = s->levels[l], &cv = s->levels[l + 1];
    constexpr int H2 = 2 * (NU + 1);
    const SmoothW sw = smooth_weights_n(NU);
    const LevelDev &F = TV(s, lv), &Cc = TV(s, cv);
    const CGScalars *scc = s->sc;
    const bool pdl_ok = l > 0;   // level 0 opens the iteration
    if (use_stream(s, lv)) {
        if (s->stream_pipe && lv.tma && NU >= 3) launch_pre_pipe<(NU >= 3 ? NU : 3)>(s, st, l, pdl_ok);
        else launch_pre_stream<NU>(s, st, l, pdl_ok);
        s->launches++;
        trace_mark(st, "pre");
        return;
    }
    if constexpr (NU == 3 || NU == 4) {
        if (s->rt_smooth && lv.rt_pre.on) {
            // interior tiles: register-tile kernel (smooth_rt.cu) on `st`; perimeter tiles: the general tile kernel on a
            // tile list, beside it on a second stream (a parallel branch of the iteration graph)
            constexpr int HX = (NU + 1) & 1;   // the register tiles need an even halo
            const bool fork = lv.rt_pre.nperim > 0;
            if (fork) {
                cudaEventRecord(s->ev_rt_fork, st);
                cudaStreamWaitEvent(s->rt_stream, s->ev_rt_fork, 0);
                LevelDev Fp = F;
                Fp.tlist = lv.rt_pre.d_tlist; Fp.tl_gx = lv.rt_pre.gx; Fp.tl_gy = lv.rt_pre.gy;
                LAUNCH_K(false, (T64::k_presmooth<NU, 8, HX>), dim3(lv.rt_pre.nperim), dim3(512), SM64_2,
                         s->rt_stream, Fp, Cc, VP(s, lv, lv.b), VP(s, lv, lv.t), VP(s, cv, cv.b), sw, scc);
                s->launches++;
                cudaEventRecord(s->ev_rt_join, s->rt_stream);
            }
            rt_launch_pre(s, st, l, NU, sw, pdl_ok);
            if (fork) {
                cudaStreamWaitEvent(st, s->ev_rt_join, 0);
                s->pdl_block = true;
            }
            s->launches++;
            trace_mark(st, "pre");
            return;
        }
    }
    xch(s, lv, lv.b, NU + 1);
    if (use_t32(s, F, 64 - H2)) {
        const size_t tsm = SM32_3;
        const dim3 g32 = tile_grid(F, 32 - H2);
        // at most one tile per SM: the kernel lasts as long as one tile does -- one node per thread (1024 threads)
        if (s->t32_wide && (int)(g32.x * g32.y) <= s->num_sms)
            LAUNCH_K(pdl_ok, (T32::k_presmooth<NU, 1>), g32, dim3(1024), tsm, st, F, Cc, VP(s, lv, lv.b), VP(s, lv, lv.t),
                     VP(s, cv, cv.b), sw, scc);
        else
            LAUNCH_K(pdl_ok, (T32::k_presmooth<NU, 4>), g32, dim3(256), tsm, st, F, Cc, VP(s, lv, lv.b),
                     VP(s, lv, lv.t), VP(s, cv, cv.b), sw, scc);
    } else {
        const size_t tsm = SM64_2, tsm3 = SM64_3;
        const dim3 g = tile_grid(F, 64 - H2);
        if ((int)(g.x * g.y) >= 2 * s->num_sms)
            LAUNCH_K(pdl_ok, (T64::k_presmooth<NU, 8>), g, dim3(512), tsm, st, F, Cc, VP(s, lv, lv.b), VP(s, lv, lv.t),
                     VP(s, cv, cv.b), sw, scc);
        else
            LAUNCH_K(pdl_ok, (T64::k_presmooth<NU, 4>), g, dim3(1024), tsm3, st, F, Cc, VP(s, lv, lv.b),
                     VP(s, lv, lv.t), VP(s, cv, cv.b), sw, scc);
    }
    s->launches++;
    trace_mark(st, "pre");
}

template <int NU>
static void launch_post(eqgpu_solver *s, cudaStream_t st, int l)
{
    Level &lv = s->levels[l], &cv = s->levels[l + 1];
    const SmoothW sw = smooth_weights_n(NU);
    const LevelDev &F = TV(s, lv), &Cc = TV(s, cv);
    if (use_stream(s, lv)) {
        if (s->stream_pipe && lv.tma && NU >= 3) launch_post_pipe<(NU >= 3 ? NU : 3)>(s, st, l);
        else launch_post_stream<NU>(s, st, l);
        s->launches++;
        trace_mark(st, "post");
        return;
    }
    double *out_dot = s->slab ? &s->sc->part_rz : &s->sc->rz_new;
    if constexpr (NU == 3 || NU == 4) {
        if (s->rt_smooth && lv.rt_post.on) {   // as in launch_pre; the r.z sum of level 0 is shared by the two kernels
            constexpr int HX = NU & 1;
            const bool fork = lv.rt_post.nperim > 0;
            if (fork) {
                cudaEventRecord(s->ev_rt_fork, st);
                cudaStreamWaitEvent(s->rt_stream, s->ev_rt_fork, 0);
                LevelDev Fp = F;
                Fp.tlist = lv.rt_post.d_tlist; Fp.tl_gx = lv.rt_post.gx; Fp.tl_gy = lv.rt_post.gy;
#define PERIM(DOT)                                                                                                       \
    LAUNCH_K(false, (T64::k_postsmooth<NU, DOT, 8, HX>), dim3(lv.rt_post.nperim), dim3(512), SM64_2,    \
             s->rt_stream, Fp, Cc, (const double *)VP(s, lv, lv.b), (const double *)VP(s, lv, lv.t), VP(s, lv, lv.x),     \
             (const double *)VP(s, cv, cv.x), sw, s->sc, s->partials, s->counters + 1, out_dot)
                if (l == 0) PERIM(true); else PERIM(false);
#undef PERIM
                s->launches++;
                cudaEventRecord(s->ev_rt_join, s->rt_stream);
            }
            rt_launch_post(s, st, l, NU, sw, l == 0, out_dot);
            if (fork) {
                cudaStreamWaitEvent(st, s->ev_rt_join, 0);
                s->pdl_block = true;
            }
            s->launches++;
            trace_mark(st, "post");
            return;
        }
    }
    if (s->slab) slab_group_begin(s);   // both exchanges in one NCCL group
    xch(s, cv, cv.x, NU + 1);   // coarse correction rows reached by the prolongation of my halo
    xch(s, lv, lv.t, NU);       // pre-smoothed iterate; lv.b halos are still valid from the pre-smoothing exchange
    if (s->slab) slab_group_end(s);
#define POST(NS, DOT, R, NT, G, SM)                                                                               \
    LAUNCH_K(true, (NS::k_postsmooth<NU, DOT, R>), G, dim3(NT), SM, st, F, Cc, (const double *)VP(s, lv, lv.b),   \
             (const double *)VP(s, lv, lv.t), VP(s, lv, lv.x), (const double *)VP(s, cv, cv.x), sw, s->sc,        \
             s->partials, s->counters + 1, out_dot)
    if (use_t32(s, F, 64 - 2 * NU)) {
        const size_t tsm = SM32_3;
        const dim3 g = tile_grid(F, 32 - 2 * NU);
        const bool wide = s->t32_wide && (int)(g.x * g.y) <= s->num_sms;
        if (l == 0) POST(T32, true, 4, 256, g, tsm);
        else if (wide) POST(T32, false, 1, 1024, g, tsm);
        else POST(T32, false, 4, 256, g, tsm);
    } else {
        const size_t tsm = SM64_2, tsm3 = SM64_3;
        const dim3 g = tile_grid(F, 64 - 2 * NU);
        const bool big = (int)(g.x * g.y) >= 2 * s->num_sms;
        if (l == 0) { if (big) POST(T64, true, 8, 512, g, tsm); else POST(T64, true, 4, 1024, g, tsm3); }
        else { if (big) POST(T64, false, 8, 512, g, tsm); else POST(T64, false, 4, 1024, g, tsm3); }
    }
#undef POST
    s->launches++;
    trace_mark(st, "post");
}

// Coarsest level: cw.n Chebyshev-Jacobi sweeps in one tile pass (halo cw.n-1).  32-node tiles while they
// keep an owned region of >= 8 nodes and the level is small (latency-bound), else 64-node tiles.
static inline bool coarsest_t32(const eqgpu_solver *s, const LevelDev &F, int n)
{
    const int to32 = 32 - 2 * (n - 1), to64 = 64 - 2 * (n - 1);
    if (to32 < 8) return false;
    if (to64 < 8) return true;
    const dim3 g32 = tile_grid(F, to32);
    return use_t32(s, F, to64) && (int)(g32.x * g32.y) <= 4 * s->num_sms;
}

static void launch_coarsest(eqgpu_solver *s, cudaStream_t st, const CoarseW &cw)
{
    Level &lv = s->levels.back();
    const LevelDev &F = TV(s, lv);
    const CGScalars *scc = s->sc;
    const int H2 = 2 * (cw.n - 1);
    xch(s, lv, lv.b, cw.n - 1);
    if (coarsest_t32(s, F, cw.n))
    {
        const dim3 gc = tile_grid(F, 32 - H2);
        if (s->t32_wide && (int)(gc.x * gc.y) <= s->num_sms)
            LAUNCH_K(true, (T32::k_coarsest<1>), tile_grid(F, 32 - H2), dim3(1024), SM32_3, st, F,
                     (const double *)VP(s, lv, lv.b), VP(s, lv, lv.x), cw, scc);
        else
            LAUNCH_K(true, (T32::k_coarsest<4>), tile_grid(F, 32 - H2), dim3(256), SM32_3, st, F,
                     (const double *)VP(s, lv, lv.b), VP(s, lv, lv.x), cw, scc);
    }
    else
        LAUNCH_K(true, (T64::k_coarsest<4>), tile_grid(F, 64 - H2), dim3(1024), SM64_3, st, F,
                 (const double *)VP(s, lv, lv.b), VP(s, lv, lv.x), cw, scc);
    s->launches++;
    trace_mark(st, "coarsest");
}

static int nu_of(const eqgpu_solver *s, int l) { return l == 0 ? s->nu : l == 1 ? s->nu1 : s->nuc; }

// Fused V-cycle (isotropic): two kernels per large level + one tail kernel.
// Leaves z = B r in levels[0].x and (when the fine level is tiled) r.z in sc->rz_new.
static void vcycle_fused(eqgpu_solver *s, cudaStream_t st)
{
    const int nl = (int)s->levels.size();
    const CoarseW cw = coarse_weights(s);
    const bool tiled_coarsest = s->tile_coarsest && coarsest_tileable(s, cw.n) && nl >= 2;
    const int lt = tiled_coarsest ? nl - 1 : (s->use_cluster ? s->ctail_first : s->tail_first);
    const SmoothW sw = smooth_weights_n(s->nuc);
    for (int l = 0; l < lt; ++l) {
        switch (nu_of(s, l)) {
        case 1: launch_pre<1>(s, st, l); break;
        case 2: launch_pre<2>(s, st, l); break;
        case 3: launch_pre<3>(s, st, l); break;
        default: launch_pre<4>(s, st, l); break;
        }
        if (l == 0 && s->defer_x) {
            // side branch: the previous iteration's x += alpha p, beside the latency-bound coarse levels
            // (its p is this iteration's input direction s->pv, whatever the parity); joined before k_apply_p
            const Level &l0 = s->levels[0];
            // few CTAs: it must not crowd the coarse-level kernels out of the SMs, and has ~70 us to finish
            cudaEventRecord(s->ev_fork, st);
            cudaStreamWaitEvent(s->side_stream, s->ev_fork, 0);
            const size_t xoff = s->slab ? (size_t)l0.dev.own0 * l0.dev.nx : 0;   // slabs: owned rows only (halo rows of p are stale)
            const size_t xn = s->slab ? (size_t)(l0.dev.own1 - l0.dev.own0) * l0.dev.nx : l0.n();
            k_update_x<<<s->xupd_blocks, 256, 0, s->side_stream>>>(xn, s->u + xoff, s->pv + xoff, s->pv + xoff, s->sc);
            cudaEventRecord(s->ev_join, s->side_stream);
            s->x_forked = true;
            s->launches++;
        }
    }
    if (tiled_coarsest) {
        launch_coarsest(s, st, cw);
    } else if (s->use_cluster) {
        const CTailDesc ctd = make_ctail_desc(s, lt, s->ctail_ncta);
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(s->ctail_ncta); cfg.blockDim = dim3(CT_THREADS); cfg.dynamicSmemBytes = s->ctail_smem;
        cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = s->ctail_ncta; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        const LevelDev *dl = s->d_levels;
        const double *bin = s->levels[lt].b;
        double *xout = s->levels[lt].x;
        int nu = s->nuc;
        const CGScalars *scp = s->sc;
        cudaLaunchKernelEx(&cfg, k_ctail, dl, ctd, bin, xout, nu, sw, cw, scp);
        s->launches++;
    } else {
        TailDesc td;
        td.first = lt; td.last = nl - 1;
        int off = 0;
        for (int l = lt; l < nl; ++l) {
            td.off[l] = off;
            off += 3 * (s->levels[l].dev.nx + 2) * (s->levels[l].dev.ny + 2);
        }
        for (int l = lt; l < nl; ++l) {
            td.soff[l] = off;
            off += 2 * (s->levels[l].dev.nx + 1) + 2 * (s->levels[l].dev.ny + 1);
        }
        td.total = off;
        k_tail<<<1, TAIL_THREADS, s->tail_smem, st>>>(s->d_levels, td, s->levels[lt].b, s->levels[lt].x, s->nuc, sw,
                                                      cw, s->sc);
        s->launches++;
    }
    for (int l = lt - 1; l >= 0; --l) {
        switch (nu_of(s, l)) {
        case 1: launch_post<1>(s, st, l); break;
        case 2: launch_post<2>(s, st, l); break;
        case 3: launch_post<3>(s, st, l); break;
        default: launch_post<4>(s, st, l); break;
        }
    }
}

// One PCG iteration of the fused (isotropic) path, enqueued on `st`.
static void enqueue_fused_iteration(eqgpu_solver *s, cudaStream_t st)
{
    Level &l0 = s->levels[0];
    const LevelDev &L = TV(s, l0), &Ll = l0.dev;
    CGScalars *sc = s->sc;
    const bool sl = s->slab;
    const int nb1 = std::min<int>(s->max_blocks, 4 * s->num_sms);
    const size_t ooff = (size_t)Ll.own0 * Ll.nx, on = (size_t)(Ll.own1 - Ll.own0) * Ll.nx;
    s->x_forked = false;
    vcycle_fused(s, st);
    if (s->levels.size() < 2 || (!s->tile_coarsest && (s->use_cluster ? s->ctail_first : s->tail_first) == 0)) {
        k_dot<<<nb1, 256, 0, st>>>(on, s->r + ooff, s->z + ooff, sc, s->partials, s->counters + 1,
                                   sl ? &sc->part_rz : &sc->rz_new);
        s->launches++;
    }
    if (sl) {
        slab_allreduce(s, &sc->part_rz, &sc->rz_new, 1);
        slab_exchange2(s, Ll, s->z, s->pv, 1);   // one NCCL group for both one-row halos
    }
    const dim3 tg = tile_grid(L, 64 - 2);
    if (s->defer_x) {
        if (s->x_forked) cudaStreamWaitEvent(st, s->ev_join, 0);
        else {   // no tiled level ran (tiny grid): nothing to hide behind, update x in line
            k_update_x<<<nb1, 256, 0, st>>>(on, s->u + ooff, s->pv + ooff, s->pv + ooff, sc);
            s->launches++;
        }
    }
    if (use_stream(s, l0) && s->stream_apply)
        launch_apply_stream(s, st, !s->defer_x || s->join_pdl, s->pv, s->pv2);
    else
        LAUNCH_K((!sl || s->peer_ok) && (!s->defer_x || s->join_pdl), T64::k_apply_p, tg, dim3(256), 0, st, L, (const double *)VP(s, l0, s->z),
                 (const double *)VP(s, l0, s->pv), VP(s, l0, s->pv2), VP(s, l0, s->Ap), sc, s->partials, s->counters + 2,
                 sl ? &sc->part_pAp : &sc->pAp);
    trace_mark(st, "apply_p");
    std::swap(s->pv, s->pv2);
    if (sl) slab_allreduce(s, &sc->part_pAp, &sc->pAp, 1);
    if (s->defer_x)
        LAUNCH_K(true, k_update_r, dim3(nb1), dim3(256), 0, st, on, s->r + ooff, (const double *)(s->Ap + ooff), sc,
                 s->partials, s->counters + 3, sl ? &sc->part_rr : (double *)nullptr);
    else
        LAUNCH_K(!sl || s->peer_ok, k_update_xr, dim3(nb1), dim3(256), 0, st, on, s->u + ooff, s->r + ooff,
                 (const double *)(s->pv + ooff), (const double *)(s->Ap + ooff), sc, s->partials, s->counters + 3,
                 sl ? 0 : 1, sl ? &sc->part_rr : &sc->rr);
    trace_mark(st, "update_xr");
    if (sl) {
        slab_allreduce(s, &sc->part_rr, &sc->rr, 1);
        if (s->defer_x) k_book_x<<<1, 1, 0, st>>>(sc);
        else k_book<<<1, 1, 0, st>>>(sc);
    }
    s->launches += 2;
}

// Two iterations (so the p ping-pong returns to its starting buffers) captured
// once into a CUDA graph; every kernel re-reads its scalars from device memory
// and exits at once when the converged flag is set, so the graph is replayed
// without any host decision in between.
static int build_iteration_graph(eqgpu_solver *s)
{
    if (s->graph_exec) return 0;
    cudaStream_t cap;
    EQ_CUDA(cudaStreamCreateWithFlags(&cap, cudaStreamNonBlocking));
    // graph 0: p in pv -> p' in pv2 ; graph 1: the other way round.  Replayed alternately.
    for (int k = 0; k < 2; ++k) {
        const int64_t l0 = s->launches;
        EQ_CUDA(cudaStreamBeginCapture(cap, cudaStreamCaptureModeThreadLocal));
        enqueue_fused_iteration(s, cap);
        cudaGraph_t graph = nullptr;
        cudaError_t e = cudaStreamEndCapture(cap, &graph);
        s->graph_launches = (int)(s->launches - l0);
        s->launches = l0;
        if (e != cudaSuccess) { s->set_error(std::string("graph capture: ") + cudaGetErrorString(e)); return EQGPU_ECUDA; }
        EQ_CUDA(cudaGraphInstantiate(k == 0 ? &s->graph_exec : &s->graph_exec2, graph, 0));
        cudaGraphDestroy(graph);
    }
    cudaStreamDestroy(cap);
    s->graph_phase = 0;
    return 0;
}

template <bool T>
static int pcg(eqgpu_solver *s)
{
    const eqgpu_params &p = s->p;
    cudaStream_t st = s->stream;
    Level &l0 = s->levels[0];
    const LevelDev &L = l0.dev;
    const dim3 blk(BX, BY), g0 = grid2d(L);
    const double rtol = p.rtol > 0 ? p.rtol : 1e-12;
    const int max_iters = p.max_iters > 0 ? p.max_iters : 200;
    const double rs_l = L.rob_l * p.robin_s[0], rs_r = L.rob_r * p.robin_s[1];
    DirData dd = make_dirdata(s);
    const int nb1 = std::min<int>(s->max_blocks, 4 * s->num_sms);
    const bool fused = !T && s->fused;
    // EQGPU_TENSOR_PRECOND=tensor: the unfused V-cycle of the tensor operator itself (round 1's preconditioner)
    static const bool iso_env = getenv("EQGPU_TENSOR_PRECOND") == nullptr || std::string(getenv("EQGPU_TENSOR_PRECOND")) != "tensor";
    const bool iso_precond = T && iso_env && s->fused && !s->slab && s->levels.size() >= 2;
    if (fused && !s->slab) {
        int rc = build_iteration_graph(s);
        if (rc) return rc;
    }

    // owned rows are contiguous in memory: flat kernels run on [own0*nx, own1*nx)
    const size_t ooff = (size_t)L.own0 * L.nx, on = (size_t)(L.own1 - L.own0) * L.nx;
    if (s->slab) slab_exchange(s, L, s->u);
    CGScalars *sc = s->sc;
    const bool sl = s->slab;
    // warm start: history only on the path that maintains it (k_init_tile + the deferred-x step tail)
    // (single GPU: the deferred-x step tail stores the history; slabs: a device copy + halo exchange)
    // (variable tensor, one GPU: k_init_hist evaluates the candidates per node, a device copy stores the history)
    const bool hist_tensor = T && !sl && s->warm > 0 && s->uh[0] != nullptr && getenv("EQGPU_TENSOR_COLD") == nullptr;
    // warm mode 7: image ring (single GPU, isotropic fused path with the deferred-x step tail); it replaces the
    // uh[]/dk[] history below.  Elsewhere mode 7 behaves as mode 3.
    bool ring = !T && !sl && s->warm == 7 && s->init_tile && fused && s->defer_x;
    if (ring && !s->ring_partials) {   // 14 more fine-level vectors: if they do not fit, mode 6 is what runs
        bool ok = true;
        for (int k = 0; k < RING_MAX && ok; ++k)
            ok = cudaMalloc(&s->ring_h[k], sizeof(double) * s->N) == cudaSuccess &&
                 cudaMalloc(&s->ring_a[k], sizeof(double) * s->N) == cudaSuccess;
        ok = ok && cudaMalloc(&s->ring_partials, sizeof(double) * 40 * s->max_blocks) == cudaSuccess;
        if (!ok) {
            (void)cudaGetLastError();
            for (int k = 0; k < RING_MAX; ++k) {
                cudaFree(s->ring_h[k]); s->ring_h[k] = nullptr;
                cudaFree(s->ring_a[k]); s->ring_a[k] = nullptr;
            }
            cudaFree(s->ring_partials); s->ring_partials = nullptr;
            cudaFree(s->ring_b); s->ring_b = nullptr;
            s->warm = 6;
            ring = false;
        }
        s->ring_n = 0;
        s->ring_head = 0;
    }
    const bool keep_hist = !ring && (hist_tensor || (!T && s->init_tile && s->fused && (sl || s->defer_x) && s->warm > 0 && s->uh[0]));
    // history depth: modes 1-3 use that many solutions, 4 three (+ least squares), 5 four (+ cubic; one GPU,
    // isotropic path)
    const int nh_max = (s->warm >= 5 && !sl && !T) ? (s->warm >= 6 ? 5 : 4) : std::min(s->warm, 3);
    const int nh = keep_hist ? std::min(s->hist, std::min(nh_max, s->warm_adaptive ? s->nh_cap : 5)) : 0;
    // least-squares combination of the history beside the fixed extrapolations (single GPU: its nine sums
    // are not rank-reduced)
    const bool ls = keep_hist && s->warm == 4 && !sl && nh >= 2;
    // quartic candidate: this step keeps its d3 when four solutions are in play; the next one may use it as d4
    const bool wrote_dk = keep_hist && !T && !sl && s->init_tile && nh >= 4 && s->dk[0] != nullptr;
    const bool d4ok = wrote_dk && nh >= 5 && s->dk_valid;
    // scratch for the extrapolation terms: Ap and pv2 are free until the first k_apply_p writes them
    if (!T && s->init_tile) {
        const dim3 gi((L.nx + 61) / 62, (L.ny + 61) / 62);
        // d3 scratch: the level-0 work vector t is free until the first pre-smoothing writes it
        // short history (a colony whose rods move; the first steps): two tiles, 512 threads, two CTAs per SM
        if (nh <= 1 && !wrote_dk)
            k_init_tile<512, 1><<<gi, 512, 2 * INIT_SMEM / 5, st>>>(L, dd, s->u, s->uh[0], s->uh[1], s->uh[2], s->uh[3], nh, s->r,
                                                                   s->z, s->Ap, s->pv2, l0.t, rs_l, rs_r, s->partials,
                                                                   s->counters + 0, sc, sl ? 1 : 0, s->dk[s->dk_cur], nullptr, 0);
        else
            k_init_tile<INIT_THREADS, 4><<<gi, INIT_THREADS, INIT_SMEM, st>>>(
                L, dd, s->u, s->uh[0], s->uh[1], s->uh[2], s->uh[3], nh, s->r, s->z, s->Ap, s->pv2, l0.t, rs_l, rs_r, s->partials,
                s->counters + 0, sc, sl ? 1 : 0, s->dk[s->dk_cur], wrote_dk ? s->dk[s->dk_cur ^ 1] : nullptr, d4ok ? 1 : 0);
        if (sl) {
            slab_allreduce(s, sc->red_src, sc->red_dst, 5);
            k_slab_unpack<<<1, 1, 0, st>>>(sc);
            s->launches++;
        }
    } else if (hist_tensor)
    {
        if (T && L.kC)
            k_init_hist<2><<<g0, blk, 0, st>>>(L, dd, s->u, s->uh[0], s->uh[1], s->uh[2], nh, s->r, s->z, s->Ap, s->pv2, rs_l,
                                               rs_r, s->partials, s->counters + 0, sc);
        else
            k_init_hist<T><<<g0, blk, 0, st>>>(L, dd, s->u, s->uh[0], s->uh[1], s->uh[2], nh, s->r, s->z, s->Ap, s->pv2, rs_l,
                                               rs_r, s->partials, s->counters + 0, sc);
    }
    else
    {
        if (T && L.kC)
            k_init<2><<<g0, blk, 0, st>>>(L, dd, s->u, s->r, s->z, rs_l, rs_r, sc, s->partials, s->counters + 0,
                                          sl ? &sc->part_rr0 : &sc->rr0, sl ? &sc->part_b2 : &sc->bnorm2);
        else
            k_init<T><<<g0, blk, 0, st>>>(L, dd, s->u, s->r, s->z, rs_l, rs_r, sc, s->partials, s->counters + 0,
                                          sl ? &sc->part_rr0 : &sc->rr0, sl ? &sc->part_b2 : &sc->bnorm2);
    }
    if (sl && !(!T && s->init_tile)) {  // per-node k_init: rank-sum the two start residuals
        slab_allreduce(s, &sc->part_rr0, &sc->rr0, 1);
        slab_allreduce(s, &sc->part_b2, &sc->bnorm2, 1);
    }
    if (ls) {
        k_ls_gram<<<nb1, 256, 0, st>>>(s->N, s->r, s->z, s->Ap, s->pv2, std::min(nh, 3), s->ls_form, s->partials,
                                       s->counters + 4, sc);
        s->launches++;
    }
    int ring_depth = RING_MAX;
    if (const char *e = getenv("EQGPU_RING_DEPTH")) ring_depth = std::max(2, std::min(atoi(e), RING_MAX));   // tuning knob
    const int ring_k = ring ? std::min(s->ring_n, ring_depth) : 0;
    if (ring_k >= 2) {
        RingPtrs rh, ra;
        for (int k = 0; k < RING_MAX; ++k) {
            rh.p[k] = s->ring_h[(s->ring_head + k) % RING_MAX];
            ra.p[k] = s->ring_a[(s->ring_head + k) % RING_MAX];
        }
        switch (ring_k) {
        case 2: launch_ring<2>(s, L, dd, rh, ra, nb1, g0, blk, rtol, max_iters); break;
        case 3: launch_ring<3>(s, L, dd, rh, ra, nb1, g0, blk, rtol, max_iters); break;
        case 4: launch_ring<4>(s, L, dd, rh, ra, nb1, g0, blk, rtol, max_iters); break;
        case 5: launch_ring<5>(s, L, dd, rh, ra, nb1, g0, blk, rtol, max_iters); break;
        case 6: launch_ring<6>(s, L, dd, rh, ra, nb1, g0, blk, rtol, max_iters); break;
        default: launch_ring<7>(s, L, dd, rh, ra, nb1, g0, blk, rtol, max_iters); break;
        }
        s->launches++;
    } else if (!T && s->init_tile && !sl && nh <= 1 && !ls && !(L.nx & 1))
        k_impose_prev<<<nb1, 256, 0, st>>>(L, dd, s->u, s->r, s->z, s->uh[0], nh, s->sc, rtol, max_iters);
    else
        k_impose<<<g0, blk, 0, st>>>(L, dd, s->u, s->r, s->z, s->Ap, s->pv2, s->uh[0], s->uh[1], s->uh[2], nh, s->sc,
                                     rtol, max_iters, ls ? 1 + s->ls_form : 0, s->partials, s->counters + 5, s->uh[3], l0.t,
                                     s->uh[4], d4ok ? s->dk[s->dk_cur] : nullptr);
    s->launches += 2;

    int issued = 0;
    int chunk = s->st.iterations > 0 ? std::max(1, s->st.iterations) : 4;
    // ring mode: the count flips between 1 and 2 (or reaches 0) late in a run; an iteration issued in vain costs a dozen
    // early-exit launches, one issued too late a host round trip and a second step tail, so predict the larger of the
    // last two counts
    if (ring && s->st.steps > 0) chunk = std::max(1, std::max(s->st.iterations, s->ring_prev_iters));
    s->graph_phase = 0;   // odd iterations leave their search direction in pv2, even ones in pv (k_update_x flush)
    double *const p_odd = s->pv2, *const p_even = s->pv;
    int tensor_swaps = 0;   // k_apply_p_asm ping-pongs the host's view of pv / pv2; restored after the loop (the isotropic iteration graphs hold the buffers by address)
#ifdef EQ_KTRACE
    if (s->st.steps == 6) {
        unsigned long long h[256];
        cudaStreamSynchronize(st);
        cudaMemcpyFromSymbol(h, g_ktrace, sizeof h);
        for (int base = 0; base < 160; base += 32) {
            fprintf(stderr, "ktrace base %d:", base);
            for (int k = 1; k <= 10; ++k) fprintf(stderr, " %lld", (long long)(h[base + k] - h[base]));
            fprintf(stderr, "\n");
        }
    }
#endif
    if (fused && getenv("EQGPU_TRACE") && s->st.steps == 5) {  // debugging aid: in-situ per-kernel times
        std::vector<cudaEvent_t> ev;
        g_trace = &ev;
        g_trace_labels.clear();
        trace_mark(st, "start");
        for (int k = 0; k < 2; ++k, ++issued) { enqueue_fused_iteration(s, st); }
        g_trace = nullptr;
        s->graph_phase = 0;
        cudaStreamSynchronize(st);
        for (size_t k = 1; k < ev.size(); ++k) {
            float ms = 0;
            cudaEventElapsedTime(&ms, ev[k - 1], ev[k]);
            fprintf(stderr, "trace rank %d %2zu %-12s %7.1f us\n", s->slab_rank, k, g_trace_labels[k], ms * 1e3);
        }
        for (auto e : ev) cudaEventDestroy(e);
    }
    while (true) {
        if (fused) {
            for (int k = 0; k < chunk && issued < max_iters; ++k, ++issued) {
                if (s->slab) { enqueue_fused_iteration(s, st); continue; }  // NCCL calls in between: no graph
                EQ_CUDA(cudaGraphLaunch(s->graph_phase == 0 ? s->graph_exec : s->graph_exec2, st));
                s->graph_phase ^= 1;
                s->launches += s->graph_launches;
            }
        } else {
            for (int k = 0; k < chunk && issued < max_iters; ++k, ++issued) {
                if (T && iso_precond) {
                    // variable tensor, one GPU: the preconditioner is the fused V-cycle of the ISOTROPIC operator with the
                    // same dt*D (spectrally equivalent to the tensor operator within the range of the tensor's
                    // eigenvalues: identity outside the rods, the shipped scalings inside them).  PCG is exact for any
                    // SPD preconditioner; the tensor operator itself is applied by k_apply below.
                    const bool dx = s->defer_x;
                    s->defer_x = false;   // (x is updated in line by k_update_xr on this path)
                    vcycle_fused(s, st);
                    s->defer_x = dx;
                } else {
                    vcycle<T>(s);
                }
                // (the fused cycle's level-0 post-smoother has left r.z in sc->rz_new, as in enqueue_fused_iteration)
                const bool have_rz = T && iso_precond &&
                                     !(!s->tile_coarsest && (s->use_cluster ? s->ctail_first : s->tail_first) == 0);
                if (!have_rz)
                    k_dot<<<nb1, 256, 0, st>>>(on, s->r + ooff, s->z + ooff, sc, s->partials, s->counters + 1,
                                               sl ? &sc->part_rz : &sc->rz_new);
                if (sl) slab_allreduce(s, &sc->part_rz, &sc->rz_new, 1);
                // opt-in (EQGPU_TENSOR_FUSE_P=1): measured slower on the B200 at 2048^2 -- 195.9 against 208.0 steps/s for the
                // unfused pair, bit-identical results; fourteen neighbour loads through L1 cost more than a second pass over p
                static const bool fuse_p = getenv("EQGPU_TENSOR_FUSE_P") != nullptr && atoi(getenv("EQGPU_TENSOR_FUSE_P")) != 0;
                const bool fused_p = T && L.kC && !sl && fuse_p;
                if (fused_p) {   // p' = z + beta p and A p' in one pass, p ping-pong
                    k_apply_p_asm<<<g0, blk, 0, st>>>(L, s->z, s->pv, s->pv2, s->Ap, sc, s->partials, s->counters + 2, &sc->pAp);
                    std::swap(s->pv, s->pv2);
                    ++tensor_swaps;
                } else
                k_update_p<<<nb1, 256, 0, st>>>(on, s->z + ooff, s->pv + ooff, sc);
                if (sl) slab_exchange(s, L, s->pv);
                if (fused_p) { /* done above */ }
                else if (T && L.kC)
                    k_apply<2, true><<<g0, blk, 0, st>>>(L, s->pv, s->Ap, sc, s->partials, s->counters + 2,
                                                         sl ? &sc->part_pAp : &sc->pAp);
                else
                    k_apply<T, true><<<g0, blk, 0, st>>>(L, s->pv, s->Ap, sc, s->partials, s->counters + 2,
                                                         sl ? &sc->part_pAp : &sc->pAp);
                if (sl) slab_allreduce(s, &sc->part_pAp, &sc->pAp, 1);
                k_update_xr<<<nb1, 256, 0, st>>>(on, s->u + ooff, s->r + ooff, s->pv + ooff, s->Ap + ooff, sc,
                                                 s->partials, s->counters + 3, sl ? 0 : 1,
                                                 sl ? &sc->part_rr : &sc->rr);
                if (sl) {
                    slab_allreduce(s, &sc->part_rr, &sc->rr, 1);
                    k_book<<<1, 1, 0, st>>>(sc);
                }
                s->launches += 4;
            }
        }
        EQ_CUDA(cudaMemcpyAsync(s->sc_host, s->sc, sizeof(CGScalars), cudaMemcpyDeviceToHost, st));
        // Speculative step tail, so that a converged step costs one host round trip instead of two: flush
        // the pending x update (the kernel checks the stamp on the device) and run the boundary-flux
        // functional before the host knows whether this was the last iteration.  If it was not, the loop
        // goes on (the in-graph k_update_x then finds nothing pending) and the tail is redone.  Not with
        // channels: their sub-steps advance state and must see the converged field only.
        const bool spec = fused && s->defer_x && !s->p.channels;
        if (fused && s->defer_x) {
            // ... and, for the next step's warm start, leave a copy of the solution in the older history slot
            // (ring mode: into the ring's oldest slot, which becomes the newest once the step has converged)
            double *const ring_slot = ring ? s->ring_h[(s->ring_head + RING_MAX - 1) % RING_MAX] : nullptr;
            if (sl)   // owned rows; the history copy of a slab is made below, once the step has converged (its halo rows are exchanged)
                k_finish_x<<<nb1, 256, 0, st>>>(on, s->u + ooff, p_odd + ooff, p_even + ooff, sc, (double *)nullptr);
            else
                k_finish_x<<<nb1, 256, 0, st>>>(l0.n(), s->u, p_odd, p_even, sc, keep_hist ? s->uh[4] : ring_slot);
            k_mark_x<<<1, 1, 0, st>>>(sc);
            s->launches += 2;
        }
        if (spec) {
            int rc = boundary_functional_enqueue(s);
            if (rc) return rc;
        }
        EQ_CUDA(cudaStreamSynchronize(st));
        if (s->slab) { int rc = slab_peer_check(s); if (rc) return rc; }
        s->functional_enqueued = spec;
        if (s->sc_host->done || issued >= max_iters) break;
        chunk = 1;
    }
    if (tensor_swaps & 1) std::swap(s->pv, s->pv2);
    s->ring_prev_iters = s->st.iterations;
    s->st.iterations = s->sc_host->iters;
    s->last_guess = s->sc_host->guess;
    if (getenv("EQGPU_LS_DEBUG")) {   // debugging aid: the candidates' residuals relative to the zero guess's
        const CGScalars &h = *s->sc_host;
        const double b2 = h.bnorm2 > 0 ? h.bnorm2 : 1.0;
        fprintf(stderr, "guess step %lld nh %d ls %d d4ok %d quartic %.2e: prev %.2e lin %.2e quad %.2e cubic %.2e ls(pred) %.2e picked %d "
                        "init(true) %.2e c = (%.6g, %.6g, %.6g) iters %d final %.2e\n",
                (long long)s->st.steps, nh, ls ? 1 + s->ls_form : 0, d4ok ? 1 : 0, sqrt(fabs(h.rrG) / b2), sqrt(h.rr0 / b2),
                sqrt(fabs(h.rrD) / b2),
                sqrt(fabs(h.rrE) / b2), sqrt(fabs(h.rrF) / b2), sqrt(fabs(h.rrL) / b2), h.guess,
                sqrt(fabs(h.rr_init) / b2), h.lsc[0], h.lsc[1], h.lsc[2], h.iters, sqrt(h.rr / b2));
    }
    // (row slabs too: the norms below are the rank-summed ones, bit-identical on every rank, so all ranks decide alike;
    // EQGPU_SLAB_ADAPTIVE=0 keeps the full history depth on slabs)
    static const bool slab_adaptive = getenv("EQGPU_SLAB_ADAPTIVE") == nullptr || atoi(getenv("EQGPU_SLAB_ADAPTIVE")) != 0;
    if (s->warm_adaptive && keep_hist && !T && (!sl || slab_adaptive)) {
        // Did the extrapolations pay?  gain = squared residual of the previous solution over the best higher candidate's.
        // Below 4 (a factor 2 in norm, a fifth of a PCG iteration) for three steps running, the operator walks over the
        // older solutions cost more than they save: walk the newest one only.  The history keeps rotating at full depth,
        // and every 64 steps the full depth is probed again for one step.
        const CGScalars &h = *s->sc_host;
        if (nh >= 2) {
            double hi = h.rrD;
            if (nh >= 3) hi = std::min(hi, h.rrE);
            if (nh >= 4) hi = std::min(hi, h.rrF);
            if (d4ok) hi = std::min(hi, h.rrG);
            if (ls) hi = std::min(hi, h.rrL);
            const bool low = !(hi * 4.0 < h.rr0);
            s->low_gain_steps = low ? s->low_gain_steps + 1 : 0;
            if (s->low_gain_steps >= 3 && s->hist >= std::min(nh_max, 3)) { s->nh_cap = 1; s->probe_countdown = 64; }
        } else if (s->nh_cap == 1 && s->hist >= 2 && --s->probe_countdown <= 0) {
            s->nh_cap = 5;
            s->low_gain_steps = 2;   // one probing step: back to the short history at once if it still does not pay
        }
    }
    if (ring && s->sc_host->rr <= s->sc_host->stop2) {   // solution copied by k_finish_x; its image by one operator walk
        const int slot = (s->ring_head + RING_MAX - 1) % RING_MAX;
        k_ring_image<<<g0, blk, 0, st>>>(L, s->u, s->ring_a[slot]);
        s->launches++;
        s->ring_head = slot;
        s->ring_n = std::min(s->ring_n + 1, RING_MAX);
    } else if (ring) {
        // not converged: the step tail has overwritten the oldest solution slot with this iterate while its image
        // slot still belongs to the old one -- a pair that no longer matches would make the next guess's residual
        // vector wrong, so the ring starts over
        s->ring_n = 0;
    }
    if (keep_hist && s->sc_host->rr <= s->sc_host->stop2) {   // the copy just written is now the newest solution
        if (sl) {   // slabs: copy now (owned rows are final), then bring the halo rows of the copy up to date
            EQ_CUDA(cudaMemcpyAsync(s->uh[4], s->u, sizeof(double) * s->N, cudaMemcpyDeviceToDevice, st));
            int rc = slab_exchange(s, L, s->uh[4]);
            if (rc) return rc;
        } else if (hist_tensor) {   // the unfused loop has no k_finish_x: plain device copy
            EQ_CUDA(cudaMemcpyAsync(s->uh[4], s->u, sizeof(double) * s->N, cudaMemcpyDeviceToDevice, st));
        }
        double *newest = s->uh[4];   // the oldest slot received the copy
        s->uh[4] = s->uh[3]; s->uh[3] = s->uh[2]; s->uh[2] = s->uh[1]; s->uh[1] = s->uh[0]; s->uh[0] = newest;
        s->hist = std::min(s->hist + 1, 5);
        // the d3 this step kept is the next step's d4 exactly when the history moved on by this one solution
        s->dk_valid = wrote_dk;
        if (wrote_dk) s->dk_cur ^= 1;
    } else {
        s->dk_valid = false;
    }
    const double ref = s->sc_host->bnorm2;
    s->st.relres = ref > 0 ? std::sqrt(s->sc_host->rr / ref) : 0.0;
    if (s->sc_host->rr > s->sc_host->stop2) {
        char buf[160];
        snprintf(buf, sizeof buf, "PCG did not converge: %d iterations, relres %.3e (rtol %.1e)",
                 s->sc_host->iters, s->st.relres, rtol);
        s->set_error(buf);
        s->unconverged++;
        if (s->noconv_policy == 0) return EQGPU_ENOCONV;
        // policy 1: the best iterate stands, the message stays readable through eqgpu_last_error, the run goes on
    }
    return 0;
}

int solver_step(eqgpu_solver *s)
{
    int rc = s->tensor ? pcg<true>(s) : pcg<false>(s);
    if (rc) return rc;
    if (s->p.channels) {
        rc = channels_step(s);
        if (rc) return rc;
    }
    if (s->functional_enqueued && !s->p.channels) boundary_functional_finish(s);   // already computed in pcg()'s tail
    else {
        rc = boundary_functional(s);
        if (rc) return rc;
    }
    s->functional_enqueued = false;
    s->st.steps++;
    s->st.kernel_launches = s->launches;
    return 0;
}

int solver_apply(eqgpu_solver *s, const double *dx, double *dy, bool constrained)
{
    const LevelDev &L = s->levels[0].dev;
    if (s->tensor)
        k_apply_check<true><<<grid2d(L), dim3(BX, BY), 0, s->stream>>>(L, dx, dy, constrained ? 1 : 0);
    else
        k_apply_check<false><<<grid2d(L), dim3(BX, BY), 0, s->stream>>>(L, dx, dy, constrained ? 1 : 0);
    s->launches++;
    EQ_CUDA(cudaGetLastError());
    return 0;
}

// Verification hook: one V-cycle on the solver's own r -> z, outside any PCG state.
int solver_precond(eqgpu_solver *s)
{
    if (s->tensor || s->slab || !s->fused || s->levels.size() < 2) {
        s->set_error("the preconditioner hook serves the fused isotropic single-GPU path");
        return EQGPU_ESTATE;
    }
    cudaStream_t st = s->stream;
    EQ_CUDA(cudaMemsetAsync(s->sc, 0, sizeof(CGScalars), st));
    const bool pdl = s->pdl;
    s->pdl = false;
    s->x_forked = false;
    vcycle_fused(s, st);
    if (s->x_forked) cudaStreamWaitEvent(st, s->ev_join, 0);
    s->pdl = pdl;
    EQ_CUDA(cudaStreamSynchronize(st));
    EQ_CUDA(cudaGetLastError());
    return 0;
}

int solver_rhs(eqgpu_solver *s, const double *du0, double *db)
{
    const LevelDev &L = s->levels[0].dev;
    k_rhs_check<<<grid2d(L), dim3(BX, BY), 0, s->stream>>>(L, du0, db, L.rob_l * s->p.robin_s[0],
                                                            L.rob_r * s->p.robin_s[1]);
    s->launches++;
    EQ_CUDA(cudaGetLastError());
    return 0;
}

// Times one named kernel in isolation (bench.py roofline line).  The vectors
// used are the solver's own work vectors; the field u is not touched.
int solver_bench(eqgpu_solver *s, const char *name, int reps, double *avg_ms, double *alg_bytes)
{
    Level &l0 = s->levels[0];
    const LevelDev &L = l0.dev;
    const dim3 blk(BX, BY), g0 = grid2d(L);
    cudaStream_t st = s->stream;
    const std::string nm(name);
    cudaEvent_t e0, e1;
    EQ_CUDA(cudaEventCreate(&e0));
    EQ_CUDA(cudaEventCreate(&e1));
    EQ_CUDA(cudaMemsetAsync(&s->sc->done, 0, sizeof(int), st));
    auto launch = [&](int k) -> bool {
        if (nm == "apply") {  // read p, write Ap (+ p.Ap): 16 B/DOF
            k_apply<false, true><<<g0, blk, 0, st>>>(L, s->pv, s->Ap, s->sc, s->partials, s->counters + 2, &s->sc->pAp);
            *alg_bytes = 16.0 * s->N;
        } else if (nm == "jacobi") {  // read x, b, write x': 24 B/DOF
            double *a = (k & 1) ? l0.t : s->z, *b = (k & 1) ? s->z : l0.t;
            k_jacobi<false, false><<<g0, blk, 0, st>>>(L, s->r, a, b, s->omega, s->sc);
            *alg_bytes = 24.0 * s->N;
        } else if (nm == "update_xr") {  // read x,r,p,Ap write x,r: 48 B/DOF
            const int nb1 = std::min<int>(s->max_blocks, 4 * s->num_sms);
            k_update_xr<<<nb1, 256, 0, st>>>(s->N, l0.t, s->z, s->pv, s->Ap, s->sc, s->partials, s->counters + 3, 1,
                                             &s->sc->rr);
            *alg_bytes = 48.0 * s->N;
        } else if (nm == "update_r") {  // read r,Ap write r: 24 B/DOF (split update, r on the critical path)
            const int nb1 = std::min<int>(s->max_blocks, 4 * s->num_sms);
            k_update_r<<<nb1, 256, 0, st>>>(s->N, s->z, s->Ap, s->sc, s->partials, s->counters + 3);
            *alg_bytes = 24.0 * s->N;
        } else if (nm == "apply_p") {  // read z,p write p',Ap: 32 B/DOF
            const dim3 tg((L.nx + 61) / 62, (L.ny + 61) / 62);
            if (use_stream(s, l0) && s->stream_apply) launch_apply_stream(s, st, false, s->pv, s->pv2);
            else T64::k_apply_p<<<tg, 256, 0, st>>>(L, s->z, s->pv, s->pv2, s->Ap, s->sc, s->partials, s->counters + 2, &s->sc->pAp);
            *alg_bytes = 32.0 * s->N;
        } else if (nm == "presmooth" || nm == "postsmooth") {
            // the level-0 smoother launches of the V-cycle as the solver issues them (streaming kernels by default)
            if (!s->fused || s->levels.size() < 2 || s->nu != 3 || s->slab) return false;
            if (nm == "presmooth") {   // read b, write x and b_coarse: 16 + 2 B/DOF
                launch_pre<3>(s, st, 0);
                *alg_bytes = 18.0 * s->N;
            } else {   // read x, b, x_coarse, write x: 24 + 2 B/DOF
                launch_post<3>(s, st, 0);
                *alg_bytes = 26.0 * s->N;
            }
        } else return false;
        return true;
    };
    if (!launch(0)) { s->set_error("unknown kernel name"); return EQGPU_EINVAL; }
    // Every timed launch starts from a cold L2: a 256 MB scratch (twice the 126 MB L2) is overwritten before it, outside
    // the event pair that brackets the launch -- replaying a kernel on the same 67-109 MB would time the L2, not HBM.
    const size_t flush_bytes = (size_t)256 << 20;
    void *flush = nullptr;
    EQ_CUDA(cudaMalloc(&flush, flush_bytes));
    double total = 0.0;
    for (int k = 0; k < reps; ++k) {
        EQ_CUDA(cudaMemsetAsync(&s->sc->done, 0, sizeof(int), st));
        EQ_CUDA(cudaMemsetAsync(flush, k & 0xff, flush_bytes, st));
        EQ_CUDA(cudaEventRecord(e0, st));
        launch(k);
        EQ_CUDA(cudaEventRecord(e1, st));
        EQ_CUDA(cudaEventSynchronize(e1));
        float ms = 0;
        EQ_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        total += ms;
    }
    cudaFree(flush);
    EQ_CUDA(cudaMemsetAsync(&s->sc->done, 0, sizeof(int), st));
    *avg_ms = total / reps;
    s->launches += reps + 1;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    EQ_CUDA(cudaGetLastError());
    return 0;
}
