// extern "C" entry points of libeqgpu.so (include/eqgpu.h).
#include <algorithm>
#include "eqgpu_internal.cuh"
#include <cstring>
#include <new>

static std::string g_create_error;

extern "C" {

void eqgpu_default_params(eqgpu_params *p)
{
    if (!p) return;
    memset(p, 0, sizeof *p);
    p->abi_version = EQGPU_ABI_VERSION;
    p->nW = 201;            // 100 sim-um x 2 nodes/um + 1   (src/main.cpp:511-516,534)
    p->nH = 41;
    p->hx = 0.5; p->hy = 0.5;
    p->dt = 0.1;            // src/main.cpp:457
    p->D = 1200.0;          // C4HSL 3e4 / lengthScaling^2 (src/eQinit.h:12, src/main.cpp:329-330)
    for (int w = 0; w < 4; ++w) { p->bc_type[w] = EQGPU_BC_DIRICHLET; p->bc_value[w] = 0.0; }  // DIRICHLET_0 / NOWALLED
    p->channels = 0;
    p->channel_iters = 48;  // src/main.cpp:417-429 with the shipped numbers
    p->channel_v = 120.0;   // src/main.cpp:476,364
    p->well_scaling = 10.0 * (25.0 / 5.0) * 0.5;  // src/fHSL.cpp:47
    p->rtol = 1e-12;
    p->max_iters = 200;
}

const char *eqgpu_last_error(const eqgpu_solver *s) { return s ? s->err.c_str() : g_create_error.c_str(); }

static int create_common(const eqgpu_params *p, int rank, int world, const void *nccl_id, eqgpu_solver **out);

int eqgpu_create(const eqgpu_params *p, eqgpu_solver **out) { return create_common(p, 0, 1, nullptr, out); }

int eqgpu_create_slab(const eqgpu_params *p, int rank, int world, const void *nccl_unique_id, eqgpu_solver **out)
{
    if (world < 1 || rank < 0 || rank >= world || (world > 1 && !nccl_unique_id)) {
        g_create_error = "bad slab rank/world/id";
        return EQGPU_EINVAL;
    }
    if (p && p->channels) { g_create_error = "row-slab mode does not run the flow channels yet"; return EQGPU_EINVAL; }
    return create_common(p, rank, world, nccl_unique_id, out);
}

int eqgpu_nccl_unique_id(void *out128) { return out128 ? slab_unique_id(out128) : EQGPU_EINVAL; }

// Pure host arithmetic (no device): the owned row range [g0, g1) of `rank` on every multigrid level of an
// nH-row mesh split over `world` ranks -- the same rule solver_setup applies (even cuts on level 0; a
// coarse row belongs to the owner of its coincident fine row).  Returns the number of levels written.
int eqgpu_slab_plan(int32_t nH, int32_t world, int32_t rank, int32_t max_levels, int32_t *g0, int32_t *g1,
                    int32_t *rows)
{
    if (nH < 3 || world < 1 || rank < 0 || rank >= world || max_levels < 1 || !g0 || !g1 || !rows) return EQGPU_EINVAL;
    auto cut = [&](int r) { return r >= world ? nH : (int)(((long long)nH * r / world) & ~1LL); };
    int n = nH, a = cut(rank), b = cut(rank + 1), l = 0;
    while (true) {
        g0[l] = a; g1[l] = b; rows[l] = n;
        ++l;
        if (l >= max_levels || n < 5 || (world > 1 && (n / 2) / world < 12)) break;
        const int nc = n / 2 + 1;
        int ca = nc, cb = 0;
        for (int I = 0; I < nc; ++I) {
            const int fi = 2 * I < n - 1 ? 2 * I : n - 1;
            if (fi >= a && fi < b) { if (I < ca) ca = I; if (I + 1 > cb) cb = I + 1; }
        }
        if (cb <= ca) break;
        n = nc; a = ca; b = cb;
    }
    return l;
}

int eqgpu_solver_path(eqgpu_solver *s)
{
    if (!s) return EQGPU_EINVAL;
    int rt_levels = 0;
    if (s->rt_smooth && !s->tensor)
        for (const auto &lv : s->levels) rt_levels += (lv.rt_pre.on && lv.rt_post.on) ? 1 : 0;
    return (s->fused ? 1 : 0) | (s->slab ? 2 : 0) | (s->slab_fused ? 4 : 0) | (s->use_cluster ? 8 : 0) |
           (s->tile_coarsest ? 16 : 0) | (s->tensor ? 32 : 0) | (std::min(rt_levels, 15) << 8);
}

int eqgpu_set_nonconvergence_policy(eqgpu_solver *s, int policy)
{
    if (!s) return EQGPU_EINVAL;
    if (policy != 0 && policy != 1) { s->set_error("non-convergence policy must be 0 or 1"); return EQGPU_EINVAL; }
    s->noconv_policy = policy;
    return 0;
}

int eqgpu_unconverged_steps(eqgpu_solver *s, int64_t *count)
{
    if (!s || !count) return EQGPU_EINVAL;
    *count = s->unconverged;
    return 0;
}

int eqgpu_set_warm_start(eqgpu_solver *s, int mode)
{
    if (!s) return EQGPU_EINVAL;
    if (mode < 0 || mode > 7) { s->set_error("warm-start mode must be 0..7"); return EQGPU_EINVAL; }
    s->warm = mode;
    s->nh_cap = 5; s->low_gain_steps = 0; s->probe_countdown = 0;
    return 0;
}

int eqgpu_comm_stats(eqgpu_solver *s, int64_t out[4])
{
    if (!s || !out) return EQGPU_EINVAL;
    out[0] = s->comm_allreduce_calls; out[1] = s->comm_allreduce_doubles; out[2] = s->comm_exchange_groups;
    out[3] = s->comm_halo_bytes;
    return 0;
}

int eqgpu_comm_peer_stats(eqgpu_solver *s, int64_t out[2])
{
    if (!s || !out) return EQGPU_EINVAL;
    out[0] = s->peer_ok ? s->comm_peer_exchanges : -1;   // -1: the peer-memory path is not in use (NCCL carries everything)
    out[1] = s->peer_ok ? s->comm_peer_allreduces : -1;
    return 0;
}

int eqgpu_get_warm_start(eqgpu_solver *s) { return s ? s->warm : EQGPU_EINVAL; }

int eqgpu_ls_solve3(const double *G, const double *f, double bb, double *c, double *pred)
{
    if (!G || !f || !c || !pred) return EQGPU_EINVAL;
    solver_ls_solve3(G, f, bb, c, pred);
    return 0;
}

int eqgpu_ring_solve(int K, const double *G, const double *f, double *c)
{
    if (!G || !f || !c || K < 1 || K > 7) return EQGPU_EINVAL;
    solver_ring_solve(K, G, f, c);
    return 0;
}

int eqgpu_last_guess(eqgpu_solver *s) { return s ? s->last_guess : EQGPU_EINVAL; }

int eqgpu_slab_rows(eqgpu_solver *s, int32_t *g0, int32_t *g1)
{
    if (!s || !g0 || !g1) return EQGPU_EINVAL;
    *g0 = s->levels[0].g0; *g1 = s->levels[0].g1;
    return 0;
}

static int create_common(const eqgpu_params *p, int rank, int world, const void *nccl_id, eqgpu_solver **out)
{
    if (!p || !out) { g_create_error = "null argument"; return EQGPU_EINVAL; }
    *out = nullptr;
    if (p->abi_version != EQGPU_ABI_VERSION) { g_create_error = "ABI version mismatch"; return EQGPU_EINVAL; }
    if (p->nW < 3 || p->nH < 3 || !(p->hx > 0) || !(p->dt > 0) || !(p->D > 0)) {
        g_create_error = "need nW,nH >= 3 and hx, dt, D > 0";
        return EQGPU_EINVAL;
    }
    for (int w = 0; w < 4; ++w) {
        const int t = p->bc_type[w];
        const bool lr = (w == EQGPU_LEFT || w == EQGPU_RIGHT);
        if (t < 0 || t > 3 || (t == EQGPU_BC_ROBIN && !lr) || (t == EQGPU_BC_DIRICHLET_CHANNEL && lr)) {
            g_create_error = "invalid boundary type for wall";
            return EQGPU_EINVAL;
        }
    }
    if (p->discretisation != EQGPU_DISC_P1 && p->discretisation != EQGPU_DISC_FD) {
        g_create_error = "unknown discretisation";
        return EQGPU_EINVAL;
    }
    if (p->discretisation == EQGPU_DISC_FD && (p->channels || world > 1)) {
        g_create_error = "the finite-difference discretisation (diffusionPETSc) has no flow channels and no row slabs";
        return EQGPU_EINVAL;
    }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        g_create_error = std::string("no usable CUDA device (there is no CPU fallback): ") +
                         (e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
        return EQGPU_ECUDA;
    }
    if (p->device < 0 || p->device >= ndev) { g_create_error = "device ordinal out of range"; return EQGPU_EINVAL; }
    eqgpu_solver *s = new (std::nothrow) eqgpu_solver();
    if (!s) { g_create_error = "out of host memory"; return EQGPU_EINVAL; }
    s->p = *p;
    if (!(s->p.hy > 0)) s->p.hy = s->p.hx;
    s->slab = world > 1;
    s->slab_rank = rank;
    s->slab_world = world;
    auto fail = [&](int rc) {
        g_create_error = s->err;
        slab_peer_teardown(s);
        slab_destroy_comm(s);
        solver_teardown(s);
        if (s->own_stream && s->stream) cudaStreamDestroy(s->stream);
        delete s;
        return rc;
    };
    if ((e = cudaSetDevice(p->device)) != cudaSuccess) { s->set_error(cudaGetErrorString(e)); return fail(EQGPU_ECUDA); }
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, p->device)) != cudaSuccess) { s->set_error(cudaGetErrorString(e)); return fail(EQGPU_ECUDA); }
    s->num_sms = prop.multiProcessorCount;
    if (p->stream) s->stream = (cudaStream_t)p->stream;
    else {
        if ((e = cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking)) != cudaSuccess) {
            s->set_error(cudaGetErrorString(e));
            return fail(EQGPU_ECUDA);
        }
        s->own_stream = true;
    }
    int rc = 0;
    if (s->slab) {
        rc = slab_init_comm(s, nccl_id);
        if (rc) return fail(rc);
    }
    rc = solver_setup(s);
    if (rc) return fail(rc);
    rc = channels_setup(s);
    if (rc) return fail(rc);
    rc = solver_refresh_levels(s);
    if (rc) return fail(rc);
    if (s->slab) {   // every vector exists now: map the other ranks' memory for the halo pulls and the scalar all-reduce
        rc = slab_peer_setup(s);
        if (rc) return fail(rc);
    }
    if (cudaStreamSynchronize(s->stream) != cudaSuccess) { s->set_error("setup sync failed"); return fail(EQGPU_ECUDA); }
    *out = s;
    return EQGPU_OK;
}

void eqgpu_destroy(eqgpu_solver *s)
{
    if (!s) return;
    cudaSetDevice(s->p.device);
    cudaStreamSynchronize(s->stream);
    slab_peer_teardown(s);
    slab_destroy_comm(s);
    solver_teardown(s);
    cudaFree(s->cells); cudaFree(s->cell_vals); cudaFree(s->cell_counts); cudaFree(s->cell_amt); cudaFree(s->bin_ints); cudaFree(s->tensor_owner);
    if (s->own_stream) cudaStreamDestroy(s->stream);
    delete s;
}

#define CHECK_S(s) do { if (!(s)) return EQGPU_EINVAL; cudaSetDevice((s)->p.device); } while (0)

int eqgpu_set_field(eqgpu_solver *s, const double *h)
{
    CHECK_S(s);
    if (!h) { s->set_error("null field"); return EQGPU_EINVAL; }
    // the host array is always the whole nW x nH field; a slab takes its window (owned + halo rows)
    EQ_CUDA(cudaMemcpyAsync(s->u, h + (size_t)s->levels[0].dev.row0 * s->p.nW, sizeof(double) * s->N,
                            cudaMemcpyHostToDevice, s->stream));
    EQ_CUDA(cudaStreamSynchronize(s->stream));
    return 0;
}

int eqgpu_get_field(eqgpu_solver *s, double *h)
{
    CHECK_S(s);
    if (!h) { s->set_error("null field"); return EQGPU_EINVAL; }
    // ... and gives back its owned rows only
    const Level &l0 = s->levels[0];
    EQ_CUDA(cudaMemcpyAsync(h + (size_t)l0.g0 * s->p.nW, s->u + (size_t)l0.dev.own0 * s->p.nW,
                            sizeof(double) * (size_t)(l0.g1 - l0.g0) * s->p.nW, cudaMemcpyDeviceToHost, s->stream));
    EQ_CUDA(cudaStreamSynchronize(s->stream));
    return 0;
}

int eqgpu_set_tensor(eqgpu_solver *s, const double *d11, const double *d22, const double *d12)
{
    CHECK_S(s);
    if (!d11 && !d22 && !d12) {
        s->tensor = false;
        return solver_refresh_levels(s);
    }
    if (!d11 || !d22 || !d12) { s->set_error("give all three tensor components or none"); return EQGPU_EINVAL; }
    if (s->p.discretisation == EQGPU_DISC_FD) { s->set_error("the variable tensor belongs to the P1 discretisation"); return EQGPU_ESTATE; }
    if (s->slab) { s->set_error("variable tensor is single-GPU only for now"); return EQGPU_ESTATE; }
    const size_t bytes = sizeof(double) * s->N;
    if (!s->d11) {
        EQ_CUDA(cudaMalloc(&s->d11, bytes));
        EQ_CUDA(cudaMalloc(&s->d22, bytes));
        EQ_CUDA(cudaMalloc(&s->d12, bytes));
    }
    EQ_CUDA(cudaMemcpyAsync(s->d11, d11, bytes, cudaMemcpyHostToDevice, s->stream));
    EQ_CUDA(cudaMemcpyAsync(s->d22, d22, bytes, cudaMemcpyHostToDevice, s->stream));
    EQ_CUDA(cudaMemcpyAsync(s->d12, d12, bytes, cudaMemcpyHostToDevice, s->stream));
    s->tensor = true;
    int rc = solver_refresh_levels(s);
    if (rc) return rc;
    EQ_CUDA(cudaStreamSynchronize(s->stream));
    return 0;
}

int eqgpu_cells_tensor(eqgpu_solver *s, double Dx, double Dy)
{
    CHECK_S(s);
    if (s->slab) { s->set_error("variable tensor is single-GPU only for now"); return EQGPU_ESTATE; }
    if (!(Dx > 0) || !(Dy > 0)) { s->set_error("axial / transverse scalings must be positive"); return EQGPU_EINVAL; }
    if (s->p.discretisation == EQGPU_DISC_FD && !(Dx == 1.0 && Dy == 1.0)) {
        s->set_error("the variable tensor belongs to the P1 discretisation");
        return EQGPU_ESTATE;
    }
    int rc = cells_tensor(s, Dx, Dy);
    if (rc) return rc;
    // Dx == Dy == 1 (the shipped values, src/eQinit.h:64-65): the grids differ from 1,1,0 only by the
    // rounding of c^2 + s^2 (<= 1 ulp), so the solve stays on the constant-coefficient kernels.
    s->tensor = !(Dx == 1.0 && Dy == 1.0);
    return solver_refresh_levels(s);
}

int eqgpu_get_tensor(eqgpu_solver *s, double *d11, double *d22, double *d12)
{
    CHECK_S(s);
    if (!d11 || !d22 || !d12) { s->set_error("null tensor output"); return EQGPU_EINVAL; }
    if (s->slab) { s->set_error("variable tensor is single-GPU only for now"); return EQGPU_ESTATE; }
    const size_t n = s->N, bytes = sizeof(double) * n;
    if (!s->d11) {   // never set: the isotropic default of src/fHSL.cpp:313-323
        for (size_t g = 0; g < n; ++g) { d11[g] = 1.0; d22[g] = 1.0; d12[g] = 0.0; }
        return 0;
    }
    EQ_CUDA(cudaMemcpyAsync(d11, s->d11, bytes, cudaMemcpyDeviceToHost, s->stream));
    EQ_CUDA(cudaMemcpyAsync(d22, s->d22, bytes, cudaMemcpyDeviceToHost, s->stream));
    EQ_CUDA(cudaMemcpyAsync(d12, s->d12, bytes, cudaMemcpyDeviceToHost, s->stream));
    EQ_CUDA(cudaStreamSynchronize(s->stream));
    return 0;
}

int eqgpu_set_boundary_value(eqgpu_solver *s, double v)
{
    CHECK_S(s);
    for (int w = 0; w < 4; ++w)
        if (s->p.bc_type[w] == EQGPU_BC_DIRICHLET) s->dir_val[w] = v;
    return 0;
}

int eqgpu_step(eqgpu_solver *s)
{
    CHECK_S(s);
    return solver_step(s);
}

int eqgpu_step_host(eqgpu_solver *s, double *v)
{
    CHECK_S(s);
    if (!v) { s->set_error("null solution_vector"); return EQGPU_EINVAL; }
    const Level &l0 = s->levels[0];
    EQ_CUDA(cudaMemcpyAsync(s->u, v + (size_t)l0.dev.row0 * s->p.nW, sizeof(double) * s->N, cudaMemcpyHostToDevice,
                            s->stream));
    int rc = solver_step(s);
    if (rc) return rc;
    EQ_CUDA(cudaMemcpyAsync(v + (size_t)l0.g0 * s->p.nW, s->u + (size_t)l0.dev.own0 * s->p.nW,
                            sizeof(double) * (size_t)(l0.g1 - l0.g0) * s->p.nW, cudaMemcpyDeviceToHost, s->stream));
    EQ_CUDA(cudaStreamSynchronize(s->stream));
    return 0;
}

int eqgpu_get_stats(eqgpu_solver *s, eqgpu_stats *out)
{
    CHECK_S(s);
    if (!out) return EQGPU_EINVAL;
    s->st.kernel_launches = s->launches;
    *out = s->st;
    return 0;
}

int eqgpu_get_channels(eqgpu_solver *s, double *top, double *bottom)
{
    CHECK_S(s);
    const size_t b = sizeof(double) * s->p.nW;
    if (top) EQ_CUDA(cudaMemcpyAsync(top, s->chan_top, b, cudaMemcpyDeviceToHost, s->stream));
    if (bottom) EQ_CUDA(cudaMemcpyAsync(bottom, s->chan_bot, b, cudaMemcpyDeviceToHost, s->stream));
    EQ_CUDA(cudaStreamSynchronize(s->stream));
    return 0;
}

int eqgpu_set_channels(eqgpu_solver *s, const double *top, const double *bottom)
{
    CHECK_S(s);
    const size_t b = sizeof(double) * s->p.nW;
    if (top) EQ_CUDA(cudaMemcpyAsync(s->chan_top, top, b, cudaMemcpyHostToDevice, s->stream));
    if (bottom) EQ_CUDA(cudaMemcpyAsync(s->chan_bot, bottom, b, cudaMemcpyHostToDevice, s->stream));
    EQ_CUDA(cudaStreamSynchronize(s->stream));
    return 0;
}

int eqgpu_get_channel_flux(eqgpu_solver *s, double *ft, double *fb)
{
    CHECK_S(s);
    const size_t b = sizeof(double) * s->p.nW;
    if (ft) EQ_CUDA(cudaMemcpyAsync(ft, s->flux_top, b, cudaMemcpyDeviceToHost, s->stream));
    if (fb) EQ_CUDA(cudaMemcpyAsync(fb, s->flux_bot, b, cudaMemcpyDeviceToHost, s->stream));
    EQ_CUDA(cudaStreamSynchronize(s->stream));
    return 0;
}

static int cells_reserve(eqgpu_solver *s, int64_t n)
{
    if (n <= s->cells_cap) return 0;
    cudaFree(s->cells); cudaFree(s->cell_vals); cudaFree(s->cell_counts); cudaFree(s->cell_amt);
    s->cells = nullptr; s->cell_vals = nullptr; s->cell_counts = nullptr; s->cell_amt = nullptr;
    s->cells_cap = 0;
    const int64_t cap = n + n / 4 + 64;
    EQ_CUDA(cudaMalloc(&s->cells, sizeof(double) * EQGPU_CELL_STRIDE * cap));
    EQ_CUDA(cudaMalloc(&s->cell_vals, sizeof(double) * cap));
    EQ_CUDA(cudaMalloc(&s->cell_amt, sizeof(double) * cap));
    EQ_CUDA(cudaMemset(s->cell_amt, 0, sizeof(double) * cap));
    EQ_CUDA(cudaMalloc(&s->cell_counts, sizeof(int32_t) * cap));
    s->cells_cap = cap;
    return 0;
}

int eqgpu_cells_upload(eqgpu_solver *s, const double *rec, int64_t n, double npm)
{
    CHECK_S(s);
    if (n < 0 || (n > 0 && !rec) || !(npm > 0)) { s->set_error("bad cell upload arguments"); return EQGPU_EINVAL; }
    int rc = cells_reserve(s, n);
    if (rc) return rc;
    s->ncells = n;
    s->counts_valid = false;
    s->npm = npm;
    if (n) EQ_CUDA(cudaMemcpyAsync(s->cells, rec, sizeof(double) * EQGPU_CELL_STRIDE * n, cudaMemcpyHostToDevice, s->stream));
    EQ_CUDA(cudaStreamSynchronize(s->stream));
    return 0;
}

int eqgpu_cells_upload_device(eqgpu_solver *s, const double *d_rec, int64_t n, double npm)
{
    CHECK_S(s);
    if (n < 0 || (n > 0 && !d_rec) || !(npm > 0)) { s->set_error("bad cell upload arguments"); return EQGPU_EINVAL; }
    int rc = cells_reserve(s, n);
    if (rc) return rc;
    s->ncells = n;
    s->counts_valid = false;
    s->npm = npm;
    if (n) EQ_CUDA(cudaMemcpyAsync(s->cells, d_rec, sizeof(double) * EQGPU_CELL_STRIDE * n, cudaMemcpyDeviceToDevice, s->stream));
    return 0;
}

int eqgpu_cells_raster(eqgpu_solver *s, int32_t *counts, int64_t *nodes, int32_t cap)
{
    CHECK_S(s);
    if (!counts || cap < 0 || (cap > 0 && !nodes)) { s->set_error("bad raster arguments"); return EQGPU_EINVAL; }
    if (s->ncells == 0) return 0;
    long long *d_nodes = nullptr;
    if (cap > 0) {
        EQ_CUDA(cudaMalloc(&d_nodes, sizeof(long long) * (size_t)cap * s->ncells));
        if (cudaMemsetAsync(d_nodes, 0xff, sizeof(long long) * (size_t)cap * s->ncells, s->stream) != cudaSuccess) {
            cudaFree(d_nodes);
            s->set_error("raster scratch memset failed");
            return EQGPU_ECUDA;
        }
    }
    int rc = cells_raster(s, s->cell_counts, d_nodes, cap);
    if (!rc) {
        cudaMemcpyAsync(counts, s->cell_counts, sizeof(int32_t) * s->ncells, cudaMemcpyDeviceToHost, s->stream);
        if (cap > 0)
            cudaMemcpyAsync(nodes, d_nodes, sizeof(long long) * (size_t)cap * s->ncells, cudaMemcpyDeviceToHost, s->stream);
        if (cudaStreamSynchronize(s->stream) != cudaSuccess) { s->set_error("raster copy failed"); rc = EQGPU_ECUDA; }
    }
    cudaFree(d_nodes);
    return rc;
}

int eqgpu_cells_gather(eqgpu_solver *s, double *out)
{
    CHECK_S(s);
    if (!out && s->ncells) { s->set_error("null output"); return EQGPU_EINVAL; }
    if (s->ncells == 0) return 0;
    int rc = cells_gather(s, s->cell_vals);
    if (rc) return rc;
    EQ_CUDA(cudaMemcpyAsync(out, s->cell_vals, sizeof(double) * s->ncells, cudaMemcpyDeviceToHost, s->stream));
    EQ_CUDA(cudaStreamSynchronize(s->stream));
    return 0;
}

int eqgpu_cells_scatter(eqgpu_solver *s, const double *amount)
{
    CHECK_S(s);
    if (!amount && s->ncells) { s->set_error("null amounts"); return EQGPU_EINVAL; }
    if (s->ncells == 0) return 0;
    EQ_CUDA(cudaMemcpyAsync(s->cell_amt, amount, sizeof(double) * s->ncells, cudaMemcpyHostToDevice, s->stream));
    int rc = cells_scatter(s, s->cell_amt);
    if (rc) return rc;
    EQ_CUDA(cudaStreamSynchronize(s->stream));
    return 0;
}

int eqgpu_set_scatter_mode(eqgpu_solver *s, int mode)
{
    CHECK_S(s);
    if (mode != 0 && mode != 1) { s->set_error("scatter mode must be 0 (direct) or 1 (binned)"); return EQGPU_EINVAL; }
    s->scatter_mode = mode;
    return 0;
}

int eqgpu_cells_set_amounts(eqgpu_solver *s, const double *amount)
{
    CHECK_S(s);
    if (!amount && s->ncells) { s->set_error("null amounts"); return EQGPU_EINVAL; }
    if (s->ncells == 0) return 0;
    EQ_CUDA(cudaMemcpyAsync(s->cell_amt, amount, sizeof(double) * s->ncells, cudaMemcpyHostToDevice, s->stream));
    EQ_CUDA(cudaStreamSynchronize(s->stream));
    return 0;
}

int eqgpu_cells_gather_resident(eqgpu_solver *s)
{
    CHECK_S(s);
    return cells_gather(s, s->cell_vals);
}

int eqgpu_cells_scatter_resident(eqgpu_solver *s)
{
    CHECK_S(s);
    return cells_scatter(s, s->cell_amt);
}

int eqgpu_cells_get_gathered(eqgpu_solver *s, double *out)
{
    CHECK_S(s);
    if (!out && s->ncells) { s->set_error("null output"); return EQGPU_EINVAL; }
    if (s->ncells == 0) return 0;
    EQ_CUDA(cudaMemcpyAsync(out, s->cell_vals, sizeof(double) * s->ncells, cudaMemcpyDeviceToHost, s->stream));
    EQ_CUDA(cudaStreamSynchronize(s->stream));
    return 0;
}

int eqgpu_bench_kernel(eqgpu_solver *s, const char *name, int reps, double *avg_ms, double *alg_bytes)
{
    CHECK_S(s);
    if (!name || reps <= 0 || !avg_ms || !alg_bytes) return EQGPU_EINVAL;
    return solver_bench(s, name, reps, avg_ms, alg_bytes);
}

int eqgpu_apply_operator(eqgpu_solver *s, const double *hx, double *hy, int constrained)
{
    CHECK_S(s);
    if (!hx || !hy) return EQGPU_EINVAL;
    if (s->slab) { s->set_error("verification hooks are single-GPU only"); return EQGPU_ESTATE; }
    const size_t bytes = sizeof(double) * s->N;
    EQ_CUDA(cudaMemcpyAsync(s->pv, hx, bytes, cudaMemcpyHostToDevice, s->stream));
    int rc = solver_apply(s, s->pv, s->Ap, constrained != 0);
    if (rc) return rc;
    EQ_CUDA(cudaMemcpyAsync(hy, s->Ap, bytes, cudaMemcpyDeviceToHost, s->stream));
    EQ_CUDA(cudaMemsetAsync(s->pv, 0, bytes, s->stream));
    EQ_CUDA(cudaStreamSynchronize(s->stream));
    return 0;
}

int eqgpu_apply_preconditioner(eqgpu_solver *s, const double *hr, double *hz)
{
    CHECK_S(s);
    if (!hr || !hz) return EQGPU_EINVAL;
    const size_t bytes = sizeof(double) * s->N;
    EQ_CUDA(cudaMemcpyAsync(s->r, hr, bytes, cudaMemcpyHostToDevice, s->stream));
    int rc = solver_precond(s);
    if (rc) return rc;
    EQ_CUDA(cudaMemcpyAsync(hz, s->z, bytes, cudaMemcpyDeviceToHost, s->stream));
    EQ_CUDA(cudaStreamSynchronize(s->stream));
    return 0;
}

int eqgpu_build_rhs(eqgpu_solver *s, const double *hu0, double *hb)
{
    CHECK_S(s);
    if (!hu0 || !hb) return EQGPU_EINVAL;
    if (s->slab) { s->set_error("verification hooks are single-GPU only"); return EQGPU_ESTATE; }
    const size_t bytes = sizeof(double) * s->N;
    EQ_CUDA(cudaMemcpyAsync(s->pv, hu0, bytes, cudaMemcpyHostToDevice, s->stream));
    int rc = solver_rhs(s, s->pv, s->Ap);
    if (rc) return rc;
    EQ_CUDA(cudaMemcpyAsync(hb, s->Ap, bytes, cudaMemcpyDeviceToHost, s->stream));
    EQ_CUDA(cudaMemsetAsync(s->pv, 0, bytes, s->stream));
    EQ_CUDA(cudaStreamSynchronize(s->stream));
    return 0;
}

int eqgpu_field_device_ptr(eqgpu_solver *s, void **p)
{
    CHECK_S(s);
    if (!p) return EQGPU_EINVAL;
    *p = s->u;
    return 0;
}

int eqgpu_sync(eqgpu_solver *s)
{
    CHECK_S(s);
    EQ_CUDA(cudaStreamSynchronize(s->stream));
    return 0;
}

}  // extern "C"
