// Flow channels and boundary flux: the tail of fenicsInterface::stepDiffusion
// (src/fHSL.cpp:110-160).
//   computeBoundaryFlux (:54-96)       -> k_channel_flux
//   numIterations CN sub-steps (:117-143) of fenics/AdvectionDiffusion.ufl:62-64
//                                      -> k_channel_substeps (one CTA per channel)
//   flux functional (:156-160, fenics/boundary.ufl:9-12) -> k_boundary_functional
#include "eqgpu_internal.cuh"
#include <cmath>
#include <vector>

#define CH_THREADS 1024

// One-sided FD flux into the channel node above/below each column, scaled to
// the channel volume element (src/fHSL.cpp:61-94).
__global__ void k_channel_flux(int nW, int nH, const double *__restrict__ u, double h, double dt,
                               double D, double well, double *__restrict__ fb, double *__restrict__ ft)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nW) return;
    const double ds = 1.0 * h;
    double gradc = (u[(size_t)1 * nW + j] - u[j]) / h;
    fb[j] = ds * (dt * D * gradc) / well;
    gradc = (u[(size_t)(nH - 2) * nW + j] - u[(size_t)(nH - 1) * nW + j]) / h;
    ft[j] = ds * (dt * D * gradc) / well;
}

struct Affine { double a, b; };  // x -> a*x + b
__device__ __forceinline__ Affine compose(const Affine &first, const Affine &then)
{
    Affine r; r.a = then.a * first.a; r.b = then.a * first.b + then.b; return r;
}

// Solves the linear recurrence y[k] = A(k) * y[k-1] + B(k), y[-1] = 0, for
// k = 0..n-1 in logical order (REVERSE walks memory backwards), writing y into
// out.  Block-wide scan of affine maps: per-thread serial composition, warp
// shuffle scan, one shared-memory pass across warps.
template <bool REVERSE, class FA, class FB>
__device__ void recurrence_scan(int n, FA Acoef, FB Bcoef, double *out, Affine *wsum)
{
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int per = (n + CH_THREADS - 1) / CH_THREADS;
    const int k0 = t * per, k1 = min(n, k0 + per);
    Affine f; f.a = 1.0; f.b = 0.0;
    for (int k = k0; k < k1; ++k) {
        const int m = REVERSE ? n - 1 - k : k;
        Affine e; e.a = Acoef(m); e.b = Bcoef(m);
        f = compose(f, e);
    }
    // inclusive scan across the warp
    Affine inc = f;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        Affine prev; prev.a = __shfl_up_sync(0xffffffffu, inc.a, o);
        prev.b = __shfl_up_sync(0xffffffffu, inc.b, o);
        if (lane >= o) inc = compose(prev, inc);
    }
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        Affine w = wsum[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            Affine prev; prev.a = __shfl_up_sync(0xffffffffu, w.a, o);
            prev.b = __shfl_up_sync(0xffffffffu, w.b, o);
            if (lane >= o) w = compose(prev, w);
        }
        wsum[lane] = w;
    }
    __syncthreads();
    // exclusive prefix of this thread = (warp prefix) o (lanes before me)
    Affine ex; ex.a = __shfl_up_sync(0xffffffffu, inc.a, 1); ex.b = __shfl_up_sync(0xffffffffu, inc.b, 1);
    if (lane == 0) { ex.a = 1.0; ex.b = 0.0; }
    if (warp > 0) ex = compose(wsum[warp - 1], ex);
    double y = ex.b;  // y[-1] = 0 -> value entering this thread's segment
    for (int k = k0; k < k1; ++k) {
        const int m = REVERSE ? n - 1 - k : k;
        y = Acoef(m) * y + Bcoef(m);
        out[m] = y;
    }
    __syncthreads();
}

// coef layout (host-built): lo[n], im[n] (1/pivot), cp[n] (Thomas c')
__global__ void __launch_bounds__(CH_THREADS)
k_channel_substeps(int n, int num_iter, double h, double dtx, double D, double v, double r1,
                   double s1, double r2, double s2, const double *__restrict__ coef,
                   const double *__restrict__ flux_top, const double *__restrict__ flux_bot,
                   double *__restrict__ u_top, double *__restrict__ u_bot, double *scratch)
{
    extern __shared__ double sm[];
    __shared__ Affine wsum[32];
    const bool use_smem = scratch == nullptr;
    double *u = use_smem ? sm : scratch + (size_t)blockIdx.x * 2 * n;
    double *d = u + n;
    const double *flux = blockIdx.x == 0 ? flux_top : flux_bot;
    double *ug = blockIdx.x == 0 ? u_top : u_bot;
    const double *lo = coef, *im = coef + n, *cp = coef + 2 * n;
    const double inv_iter = 1.0 / (double)num_iter;
    const double dd = 0.5 * dtx * D / h, dv = 0.25 * dtx * v;
    for (int j = threadIdx.x; j < n; j += CH_THREADS) u[j] = ug[j];
    __syncthreads();
    for (int it = 0; it < num_iter; ++it) {
        // src/fHSL.cpp:126-135: spread the step's flux over the sub-steps
        for (int j = threadIdx.x; j < n; j += CH_THREADS) u[j] += flux[j] / (double)num_iter;
        __syncthreads();
        // CN load vector (fenics/AdvectionDiffusion.h:2444-2648, closed form)
        auto rhs = [&](int j) {
            double b = 0.0;
            const double uj = u[j];
            if (j > 0) {  // element (j-1, j), local row 1
                const double du = uj - u[j - 1];
                b += h * (u[j - 1] * (1.0 / 6.0) + uj * (1.0 / 3.0)) - dv * du - dd * du;
            }
            if (j < n - 1) {  // element (j, j+1), local row 0
                const double du = u[j + 1] - uj;
                b += h * (uj * (1.0 / 3.0) + u[j + 1] * (1.0 / 6.0)) - dv * du + dd * du;
            }
            if (j == 0) b += -dtx * r1 * (0.5 * uj - s1);
            if (j == n - 1) b += -dtx * r2 * (0.5 * uj - s2);
            return b;
        };
        // forward elimination: d[i] = (rhs[i] - lo[i] d[i-1]) * im[i]
        recurrence_scan<false>(
            n, [&](int i) { return -lo[i] * im[i]; }, [&](int i) { return rhs(i) * im[i]; }, d, wsum);
        // back substitution: x[i] = d[i] - cp[i] x[i+1]
        recurrence_scan<true>(
            n, [&](int i) { return -cp[i]; }, [&](int i) { return d[i]; }, u, wsum);
    }
    (void)inv_iter;
    for (int j = threadIdx.x; j < n; j += CH_THREADS) ug[j] = u[j];
}

// -oint grad(u).n ds over the four walls (fenics/boundary.h:2652-2741 on the "right" mesh: each wall
// facet sees the one-sided difference of the triangle it belongs to).  u holds global rows
// [row0, row0+ny_loc); every facet term is added by the rank that owns the row it is attached to
// (left facet i -> row i+1, right facet i -> row i, bottom -> row 0, top -> row nH-1).
// Single block, deterministic.
__global__ void __launch_bounds__(1024)
k_boundary_functional(int nW, int nH, double hx, double hy, const double *__restrict__ u, double *out,
                      int row0, int g0, int g1)
{
    double acc = 0.0;
    const double ry = hx / hy, rx = hy / hx;
    const int nb = nW - 1;
    auto U = [&](int gi, int j) { return u[(size_t)(gi - row0) * nW + j]; };
    if (g0 == 0)  // bottom: lower triangle (v0,v1,v3): du/dy = (u_TR - u_BR)/hy
        for (int j = threadIdx.x; j < nb; j += blockDim.x) acc += (U(1, j + 1) - U(0, j + 1)) * ry;
    if (g1 == nH)  // top: upper triangle (v0,v2,v3): -du/dy = (u_BL - u_TL)/hy
        for (int j = threadIdx.x; j < nb; j += blockDim.x) acc += (U(nH - 2, j) - U(nH - 1, j)) * ry;
    for (int gi = g0 + threadIdx.x; gi < g1; gi += blockDim.x) {
        if (gi >= 1) acc += (U(gi, 1) - U(gi, 0)) * rx;                     // left facet gi-1: du/dx = (u_TR - u_TL)/hx
        if (gi <= nH - 2) acc += (U(gi, nW - 2) - U(gi, nW - 1)) * rx;      // right facet gi: -du/dx = (u_BL - u_BR)/hx
    }
    __shared__ double sm[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    acc = warp_sum(acc);
    if (lane == 0) sm[warp] = acc;
    __syncthreads();
    if (warp == 0) {
        double w = lane < (blockDim.x >> 5) ? sm[lane] : 0.0;
        w = warp_sum(w);
        if (lane == 0) *out = w;
    }
}

int channels_setup(eqgpu_solver *s)
{
    const eqgpu_params &p = s->p;
    if (!p.channels) return 0;
    if (p.channel_iters <= 0 || p.well_scaling <= 0) {
        s->set_error("channels enabled but channel_iters / well_scaling not set");
        return EQGPU_EINVAL;
    }
    // Thomas factors of the constant CN matrix (fenics/AdvectionDiffusion.h:2277-2288,
    // :2349-2353): lo/di/up per node, ends carry 0.5*dt*r.
    const int n = p.nW;
    const double h = p.hx, dtx = p.dt / (double)p.channel_iters;
    const double dd = 0.5 * dtx * p.D / h, dv = 0.25 * dtx * p.channel_v;
    const double a00 = h / 3.0 - dv + dd, a01 = h / 6.0 + dv - dd;
    const double a10 = h / 6.0 - dv - dd, a11 = h / 3.0 + dv + dd;
    std::vector<double> coef(3 * (size_t)n, 0.0);
    double *lo = coef.data(), *im = lo + n, *cp = im + n;
    double cprev = 0.0;
    for (int i = 0; i < n; ++i) {
        double di = 0.0, up = 0.0, l = 0.0;
        if (i > 0) { di += a11; l = a10; }
        if (i < n - 1) { di += a00; up = a01; }
        if (i == 0) di += 0.5 * dtx * p.channel_r[0];
        if (i == n - 1) di += 0.5 * dtx * p.channel_r[1];
        const double m = di - l * cprev;
        lo[i] = l;
        im[i] = 1.0 / m;
        cp[i] = up / m;
        cprev = cp[i];
    }
    EQ_CUDA(cudaMalloc(&s->chan_coef, sizeof(double) * (3 * (size_t)n + 4 * (size_t)n)));
    EQ_CUDA(cudaMemcpy(s->chan_coef, coef.data(), sizeof(double) * 3 * n, cudaMemcpyHostToDevice));
    return 0;
}

int channels_step(eqgpu_solver *s)
{
    const eqgpu_params &p = s->p;
    const int n = p.nW;
    k_channel_flux<<<(n + 255) / 256, 256, 0, s->stream>>>(p.nW, p.nH, s->u, p.hx, p.dt, p.D,
                                                           p.well_scaling, s->flux_bot, s->flux_top);
    const size_t smem = sizeof(double) * 2 * (size_t)n;
    double *scratch = nullptr;
    size_t dyn = smem;
    if (smem > 200 * 1024) { scratch = s->chan_coef + 3 * (size_t)n; dyn = 0; }
    if (dyn > 48 * 1024)
        EQ_CUDA(cudaFuncSetAttribute(k_channel_substeps, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
    k_channel_substeps<<<2, CH_THREADS, dyn, s->stream>>>(
        n, p.channel_iters, p.hx, p.dt / (double)p.channel_iters, p.D, p.channel_v, p.channel_r[0], 0.0,
        p.channel_r[1], 0.0, s->chan_coef, s->flux_top, s->flux_bot, s->chan_top, s->chan_bot, scratch);
    s->launches += 2;
    EQ_CUDA(cudaGetLastError());
    return 0;
}

// Functional kernel + copy of the scalar to pinned memory, stream-ordered, no host synchronisation.
int boundary_functional_enqueue(eqgpu_solver *s)
{
    const eqgpu_params &p = s->p;
    const double hy = p.hy > 0 ? p.hy : p.hx;
    const Level &l0 = s->levels[0];
    k_boundary_functional<<<1, 1024, 0, s->stream>>>(p.nW, p.nH, p.hx, hy, s->u, s->flux_dev, l0.dev.row0, l0.g0,
                                                     l0.g1);
    s->launches++;
    if (s->slab) {
        int rc = slab_allreduce(s, s->flux_dev, s->flux_dev, 1);
        if (rc) return rc;
    }
    EQ_CUDA(cudaMemcpyAsync(s->flux_host, s->flux_dev, sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    return 0;
}

// After the stream has been synchronised: src/fHSL.cpp:160
void boundary_functional_finish(eqgpu_solver *s) { s->st.total_boundary_flux = s->p.D * s->p.dt * (*s->flux_host); }

int boundary_functional(eqgpu_solver *s)
{
    int rc = boundary_functional_enqueue(s);
    if (rc) return rc;
    EQ_CUDA(cudaStreamSynchronize(s->stream));
    boundary_functional_finish(s);
    return 0;
}
