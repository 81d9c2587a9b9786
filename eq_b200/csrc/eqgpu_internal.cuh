// Internal declarations shared by the translation units of libeqgpu.so.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>
#include "../../include/eqgpu.h"

#define EQ_CUDA(call)                                                          \
    do {                                                                       \
        cudaError_t e__ = (call);                                              \
        if (e__ != cudaSuccess) {                                              \
            s->set_error(std::string(#call) + ": " + cudaGetErrorString(e__)); \
            return EQGPU_ECUDA;                                                \
        }                                                                      \
    } while (0)

// Stencil band order: centre, E, W, N, S, NE, SW on the "right"-diagonal mesh.
enum { B_C = 0, B_E, B_W, B_N, B_S, B_NE, B_SW, NBAND };

// One grid of the multigrid hierarchy: a tensor-product mesh of nx x ny nodes
// whose rectangles are split by the BL->TR diagonal ("right" RectangleMesh,
// src/fHSL.cpp:171).  Level 0 is uniform; coarse levels keep every even node
// plus the last one, so only their last cell can be narrower.
struct LevelDev {
    int nx, ny;
    // padded cell sizes: hx[j] = width of the cell WEST of node j (hx[0] = 0),
    // hx[j+1] = width of the cell EAST of node j (hx[nx] = 0).  ihx = 1/hx or 0.
    const double *hx, *ihx, *hy, *ihy;
    double tau;           // dt * D
    double rob_l, rob_r;  // dt * r_left, dt * r_right (0 when the wall is not Robin)
    unsigned dirmask;     // bit0 left, bit1 right, bit2 top, bit3 bottom are Dirichlet
    // uniform fast path: nodes 1<=j<=jreg_hi, 1<=i<=ireg_hi see four regular cells
    int jreg_hi, ireg_hi;
    double cC, cEW, cNS, cD, icC;
    // mass row on regular nodes (load vector b = M u0): centre and each of the six neighbours
    double mC, mO;
    // 1: diffusionPETSc's discretisation (diffuclass.cpp:786-862) -- 5-point finite differences with unit
    // (lumped) mass and ghost-node Robin rows, i.e. lumped nodal mass (aw+ae)(bs+bn)/4 and lumped Robin edge
    // mass on the same tensor-product grid; 0: the consistent P1 mass of fenics/hslD.ufl
    int lumped;
    // optional nodal tensor fields (level-local, injected); null = isotropic
    const double *d11, *d22, *d12;
    // row-slab decomposition: the arrays hold global rows [row0, row0+ny) of a gny-row grid, of
    // which local rows [own0, own1) are owned (the others are halo copies).  Single GPU: 0, ny, 0, ny.
    int row0, gny, own0, own1;
    // tile kernels (global row indices): rows [slo, shi) are present in storage (pointers are passed
    // pre-offset so that row gi lives at ptr[gi*nx]), rows [wlo, whi) may be written by this rank, and
    // tile rows start at the even row tbase.  Single GPU: 0, ny, 0, ny, 0.
    int slo, shi, wlo, whi, tbase;
    // closed form of the padded cell sizes (uniform cells, only the last one may be narrower): regular and last cell
    // widths / heights of the GLOBAL grid, so that set-up code needs no global loads (mg_stream.cuh)
    double hxr, hxl, hyr, hyl;
    // variable tensor: the assembled rows, packed by symmetry -- centre, east, north, north-east coefficient of every node
    // (west / south / south-west are the neighbour's east / north / north-east); null = evaluate the row from d11/d22/d12
    const double *kC, *kE, *kN, *kNE;
    // tile-list launches (mg_rt.cuh): when set, a tile kernel's CTA q works on tile (tlist[2q], tlist[2q+1]) of the level's
    // tl_gx x tl_gy tile grid instead of (blockIdx.x, blockIdx.y), and a reduction it takes part in is shared with another
    // kernel: one partial per tile of the whole grid, tl_gx * tl_gy arrivals in all
    const int *tlist;
    int tl_gx, tl_gy;
};

struct SmoothW { double w[4]; };          // per-sweep Jacobi weights (Chebyshev roots)

struct CGScalars {
    double rz_old, rz_new, pAp, rr, bnorm2, stop2, rr0;
    // slab mode: kernels leave their rank-local sums here; an out-of-place all-reduce then writes the
    // global value into the field above (idempotent when replayed after convergence)
    double part_rz, part_pAp, part_rr, part_b2, part_rr0, part_rrD, part_rrE;
    int iters, done, max_iters, pad;
    // deferred x update (single-GPU fused path): k_update_r leaves the step length and the iteration number
    // here, k_update_x applies x += alpha_x * p later, off the critical path; x_applied == x_stamp: nothing pending
    double alpha_x;
    int x_stamp, x_applied;
    // warm start: squared residual of the extrapolated guess (k_init_tile) and the guess k_impose picked
    double rrD, rrE;
    int guess, pad2;
    // least-squares guess (warm mode 4): coefficients of h0, h0-h1, h1-h2 and the predicted squared residual
    double lsc[3], rrL;
    double rr_init;   // squared residual of the starting guess (measured in k_impose with ls, else the candidate's)
    double rrF, part_rrF;   // cubic extrapolation of the last four solutions (warm mode 5)
    double rrG;             // quartic extrapolation of the last five (warm mode 6)
    // image ring (warm mode 7): weights of the backward differences nabla^j h0 in the guess (1 + least-squares
    // correction), the squared residual the fit predicts, the ring depth used
    double ringw[8], rrR;
    int ring_k, pad3;
    // slab mode, start of a step: the five rank-local sums of k_init_tile travel in ONE all-reduce
    double red_src[6], red_dst[6];
};

struct Level {
    LevelDev dev{};    // local view (row 0 = first stored row): unfused kernels
    LevelDev gdev{};   // global view for the tile kernels (== dev on a single GPU)
    double *d_hy_g = nullptr, *d_ihy_g = nullptr;  // slab mode: padded row sizes of the whole grid
    std::vector<double> hx_host, hy_host;  // unpadded cell sizes (global grid)
    int g0 = 0, g1 = 0;                    // owned global rows [g0, g1)
    double *d_hx = nullptr, *d_ihx = nullptr, *d_hy = nullptr, *d_ihy = nullptr;
    double *x = nullptr, *b = nullptr, *t = nullptr;  // solution, rhs, scratch
    double *t11 = nullptr, *t22 = nullptr, *t12 = nullptr;
    double *kC = nullptr, *kE = nullptr, *kN = nullptr, *kNE = nullptr;   // assembled tensor rows (solver_refresh_levels)
    // streaming smoothers (mg_stream.cuh): TMA descriptors of the level's b and t vectors (row pitch 16-byte aligned)
    CUtensorMap map_b{}, map_t{};
    bool tma = false;
    // register-tile smoothers (mg_rt.cuh): 64 x 64 TMA boxes of b and t, the rectangle of regular tiles they take, the list
    // of the remaining (perimeter) tiles for the shared-memory tile kernels; one set per kernel shape (pre, post)
    CUtensorMap map_b64{}, map_t64{};
    struct RtPlan {
        bool on = false;
        int gx = 0, gy = 0, n = 0, nperim = 0;      // tiling, tiles of the register-tile kernel, tiles of the general kernel
        int *d_tiles = nullptr, *d_tlist = nullptr;  // their (bx, by) lists
    } rt_pre, rt_post;
    bool rt_tma = false;   // 16-byte row pitch: TMA staging (else the register-tile kernels load their rows directly)
    size_t n() const { return (size_t)dev.nx * dev.ny; }
};

struct eqgpu_solver {
    eqgpu_params p;
    std::string err;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int num_sms = 148;
    size_t N = 0;
    std::vector<Level> levels;
    // fine-level vectors
    double *u = nullptr;   // resident field == PCG iterate x
    double *r = nullptr, *pv = nullptr, *pv2 = nullptr, *Ap = nullptr, *z = nullptr;
    LevelDev *d_levels = nullptr;  // device copy of every level descriptor (k_tail)
    int tail_first = 0;            // first level handled by the single-CTA tail kernel
    size_t tail_smem = 0;
    bool fused = true;
    bool tail_fits = true;         // the deepest levels fit one CTA's shared memory (k_tail / k_ctail usable)
    bool tile_coarsest = true;     // coarsest level solved by the deep-halo tile kernel instead of a tail kernel
    bool defer_x = false;          // x += alpha p runs beside the coarse levels of the next iteration (k_update_x)
    cudaStream_t side_stream = nullptr;
    bool x_forked = false;
    bool functional_enqueued = false;  // pcg() already ran the flux functional in its speculative step tail
    int xupd_blocks = 296;         // CTAs of the deferred k_update_x (few: it runs beside the coarse levels)
    bool join_pdl = false;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    // warm start (single-GPU isotropic fused path): the last three solutions, newest first; k_init_tile tries
    // the previous solution and its linear / quadratic extrapolation as starting guesses
    double *uh[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    // quartic candidate: A (h3 - h4) of this step is A (h2 - h3) of the previous one, so k_init_tile keeps its d3
    // in dk[dk_cur ^ 1] and reads the previous step's from dk[dk_cur]; valid after a step that wrote it and
    // rotated the history
    double *dk[2] = {nullptr, nullptr};
    int dk_cur = 0;
    bool dk_valid = false;
    int hist = 0;                  // valid entries of uh[]
    int warm = 3;                  // 0 off, 1 previous solution, 2 + linear, 3 + quadratic extrapolation, 4 = 3 + the
                                   // residual-minimising combination of the last three solutions, 5 = 3 + cubic
                                   // extrapolation of the last four, 6 = 5 + quartic of the last five
                                   // (solver_setup: 4 up to 512^2 nodes, 6 above)
    int last_guess = 0;
    // adaptive history depth: when the extrapolations stop paying (a colony whose rods move: the change of the sources is
    // not predictable from the history) only the previous solution is walked; the full depth is probed again periodically
    int nh_cap = 5, low_gain_steps = 0, probe_countdown = 0;
    bool warm_adaptive = true;
    // warm mode 7 (opt-in; profiles/r01_guess_study.md): ring of the last RING_MAX solutions and of their images
    // A_ff h (free rows), slot (ring_head + i) % RING_MAX = i-th newest; ring_b keeps this step's reduced right-hand
    // side until the step ends, when the new image is b - r_final (no operator walk)
    double *ring_h[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    double *ring_a[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    double *ring_b = nullptr, *ring_partials = nullptr;
    int ring_n = 0, ring_head = 0, ring_prev_iters = 0;
    int ls_form = 1;               // least-squares guess: 1 = correction to h0 fitted to r1 on {A h0, d1, d1-d2}; 0 = first form
    bool init_tile = true;         // shared-tile k_init_tile instead of the per-node k_init (isotropic, one GPU)
    bool pdl = false;              // programmatic dependent launch between the kernels of a PCG iteration
    bool t32_wide = false;         // 32-node tiles of a level with at most one tile per SM: one node per thread (1024 threads; measured no faster: the small levels are bound by kernel hand-off and load latency, not by the sweeps)
    int t32_below = 148;           // levels with fewer 64-node tiles than this run on 32-node tiles
    // row-slab mode (eqgpu_create_slab): this rank owns rows [levels[l].g0, levels[l].g1) of every level
    bool slab = false;
    int slab_rank = 0, slab_world = 1;
    void *nccl_comm = nullptr;
    long long comm_allreduce_calls = 0, comm_allreduce_doubles = 0, comm_exchange_groups = 0, comm_halo_bytes = 0;   // cumulative, this rank
    int slab_group_depth = 0;      // open slab_group_begin() brackets
    // peer-memory halos and scalar all-reduce (slab.cu): the other ranks' flag blocks and staging buffers mapped through
    // CUDA IPC; an exchange is one small kernel that stores my boundary rows into the neighbours' staging buffers and
    // copies theirs out of mine once their arrival flag is up -- no NCCL call on the data path
    unsigned long long *peer_flags = nullptr;                   // my flag block (PEER_FLAG_WORDS words)
    unsigned long long *peer_flags_of[16] = {};                 // every rank's flag block as mapped here (mine included)
    void *peer_stage = nullptr, *peer_stage_lo = nullptr, *peer_stage_hi = nullptr;   // my staging buffers, rank-1's, rank+1's
    unsigned long long peer_stage_cap = 0, peer_batch_fill = 0;   // 16-byte slots per (side, parity) buffer; filled by the open batch
    int *peer_err = nullptr;                                    // mapped pinned: set by a kernel whose wait timed out
    std::vector<void *> peer_opened;                            // cudaIpcOpenMemHandle results to close
    bool peer_ok = false;
    unsigned long long peer_xseq = 0, peer_arseq = 0;           // exchanges / all-reduces issued so far (same on every rank)
    long long peer_timeout_ns = 20000000000LL;
    void *peer_batch = nullptr;                                 // jobs collected between slab_group_begin/end
    long long comm_peer_exchanges = 0, comm_peer_allreduces = 0;
    int halo = 1;                  // halo rows kept per neighbour (1 unfused, 6 for the tile kernels)
    bool slab_fused = false;
    int scatter_mode = 0;          // 0 direct global atomics, 1 shared-memory-binned
    int *bin_ints = nullptr;       // binned scatter scratch
    long long bin_cap_cells = 0;
    int bin_cap_tiles = 0;
    CUtensorMap map_z{}, map_pv{}, map_pv2{};   // level-0 z and the two search-direction buffers (ks_apply_p)
    bool tma_p = false;
    const double *map_pv_ptr = nullptr;   // the buffer map_pv describes
    bool rt_smooth = true;         // register-tile smoothers (mg_rt.cuh) on the interior tiles of the large levels
    int rt_min_tiles = 148;        // ... of levels with at least this many regular tiles
    int rt_ctas = 0;               // persistent CTAs of a register-tile kernel (2 per SM)
    bool rt_lean = false;          // TMA levels: the three-CTAs-per-SM instances (right-hand side in shared memory; opt-in)
    cudaStream_t rt_stream = nullptr;   // perimeter tiles run beside the interior ones
    cudaEvent_t ev_rt_fork = nullptr, ev_rt_join = nullptr;
    unsigned *rt_sched = nullptr;  // tile counters of the persistent kernels
    bool pdl_block = false;        // the next launch follows a stream join: no programmatic edge
    bool stream_pipe = true;       // warp-specialised sweep pipelines (PIPE::kp_*) on TMA-capable levels
    bool stream_uni = true;        // constant-bank coefficient instances where every column is regular or Dirichlet
    bool stream_apply = false;     // ks_apply_p instead of the tile k_apply_p
    bool stream_smooth = false;    // warp-streaming smoothers (mg_stream.cuh) instead of the shared-memory tile kernels
    int stream_min_nodes = 0;      // ... on levels with at least this many nodes
    bool use_cluster = false;      // deepest levels on a 16-CTA cluster (k_ctail) instead of one CTA (k_tail)
    int ctail_first = 0, ctail_ncta = 0;
    size_t ctail_smem = 0;
    cudaGraphExec_t graph_exec = nullptr, graph_exec2 = nullptr;  // one fused PCG iteration each (p ping / pong)
    int graph_phase = 0;
    int graph_launches = 0;
    double *d11 = nullptr, *d22 = nullptr, *d12 = nullptr;
    int *tensor_owner = nullptr;   // per node: highest record index of the rods covering it (cells_tensor)
    // Dirichlet data
    double dir_val[4] = {0, 0, 0, 0};
    double *chan_top = nullptr, *chan_bot = nullptr;       // channel u (nW)
    double *flux_top = nullptr, *flux_bot = nullptr;       // per-step channel flux
    double *chan_coef = nullptr;                           // tridiagonal factors
    // reductions
    double *partials = nullptr;  // [nslots][max_blocks]
    unsigned *counters = nullptr;
    CGScalars *sc = nullptr;
    CGScalars *sc_host = nullptr;  // pinned
    double *flux_dev = nullptr;    // total boundary functional
    double *flux_host = nullptr;   // pinned
    int max_blocks = 0;
    // cells
    double *cells = nullptr;
    int64_t ncells = 0, cells_cap = 0;
    double npm = 0;
    double *cell_vals = nullptr;   // device per-cell gather result
    double *cell_amt = nullptr;    // device per-cell deposit amounts (nM)
    int32_t *cell_counts = nullptr;
    bool counts_valid = false;     // cell_counts match the uploaded cell set (left by the last gather/raster)
    double *stage_host = nullptr;  // pinned staging for field copies
    // stats
    eqgpu_stats st{};
    int64_t launches = 0;
    int nu1 = 3;                   // sweeps on level 1 (default: as the coarser levels)
    int nu = 3, nuc = 3, ncoarse = 24;  // smoothing sweeps on level 0 / on the coarser levels
    double omega = 0.8;
    bool tensor = false;
    int noconv_policy = 0;         // 0: EQGPU_ENOCONV; 1: report and continue with the best iterate
    int64_t unconverged = 0;

    void set_error(const std::string &m) { err = m; }
};

// ---- solver.cu ----
int solver_setup(eqgpu_solver *s);
void solver_teardown(eqgpu_solver *s);
int solver_step(eqgpu_solver *s);
int solver_apply(eqgpu_solver *s, const double *dx, double *dy, bool constrained);
int solver_rhs(eqgpu_solver *s, const double *du0, double *db);
int solver_precond(eqgpu_solver *s);   // z = B r on the solver's own vectors (verification hook)
int solver_refresh_levels(eqgpu_solver *s);
int solver_bench(eqgpu_solver *s, const char *name, int reps, double *avg_ms, double *alg_bytes);
// ---- smooth_rt.cu ----
int rt_setup(eqgpu_solver *s);       // after the hierarchy exists: plans, descriptors, counters
void rt_teardown(eqgpu_solver *s);
// interior tiles of level l on `st` (the caller has forked the perimeter tiles); false: this level does not use them
void rt_launch_pre(eqgpu_solver *s, cudaStream_t st, int l, int nu, const SmoothW &sw, bool pdl_ok);
void rt_launch_post(eqgpu_solver *s, cudaStream_t st, int l, int nu, const SmoothW &sw, bool dot, double *out_dot);
void solver_ls_solve3(const double G[6], const double f[3], double bb, double c[3], double *pred);
// host-side entry for the CPU tests of the K x K ring fit (eqgpu_ring_solve); G packed upper triangle, row-major
void solver_ring_solve(int K, const double *G, const double *f, double *c);
// ---- slab.cu ----
int slab_init_comm(eqgpu_solver *s, const void *unique_id);
void slab_destroy_comm(eqgpu_solver *s);
int slab_exchange(eqgpu_solver *s, const LevelDev &L, double *v, int depth = 1);  // halo rows of a level vector (local view)
int slab_allreduce(eqgpu_solver *s, const double *src, double *dst, int count);  // sum over ranks, stream-ordered
int slab_exchange2(eqgpu_solver *s, const LevelDev &L, double *v1, double *v2, int depth);   // two vectors, one NCCL group
int slab_group_begin(eqgpu_solver *s);   // bracket several exchanges into one NCCL group (no kernel in between)
int slab_group_end(eqgpu_solver *s);
void solver_trace_mark(cudaStream_t st, const char *label);   // EQGPU_TRACE debugging aid (solver.cu)
int slab_peer_setup(eqgpu_solver *s);     // after every vector exists: map the neighbours' memory (no-op unless enabled)
void slab_peer_teardown(eqgpu_solver *s);
int slab_peer_check(eqgpu_solver *s);     // after a stream sync: did a peer wait time out?  (error set, peer path switched off)
int slab_unique_id(void *out128);
// ---- cells.cu ----
int cells_raster(eqgpu_solver *s, int32_t *d_counts, long long *d_nodes, int cap);
int cells_gather(eqgpu_solver *s, double *d_out);
int cells_scatter(eqgpu_solver *s, const double *d_amount);
int cells_tensor(eqgpu_solver *s, double Dx, double Dy);
// ---- channels.cu ----
int channels_setup(eqgpu_solver *s);
int channels_step(eqgpu_solver *s);
int boundary_functional(eqgpu_solver *s);
int boundary_functional_enqueue(eqgpu_solver *s);   // kernel + scalar copy, no host sync
void boundary_functional_finish(eqgpu_solver *s);   // after the stream was synchronised

#ifdef __CUDACC__
// --------------------------------------------------------------------------
// device helpers
// --------------------------------------------------------------------------
// Programmatic dependent launch (sm_90+): a kernel launched with the programmatic-serialization attribute
// may start while its predecessor is still running; pdl_wait() blocks until the predecessor has completed
// and its writes are visible, pdl_trigger() lets the successor's CTAs be scheduled into free SM slots.
// Both are no-ops for ordinary launches.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// Asynchronous global -> shared copies (LDGSTS), 8-byte granularity because tile origins are odd node
// offsets; completion by commit groups.  k_init_tile uses them to have all its input tiles in flight at once.
__device__ __forceinline__ void cp_async8(double *smem_dst, const double *gsrc)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ bool is_dirichlet(const LevelDev &L, int i, int j)
{
    unsigned m = L.dirmask;
    return ((m & 1u) && j == 0) || ((m & 2u) && j == L.nx - 1) ||
           ((m & 4u) && i == L.ny - 1) || ((m & 8u) && i == 0);
}

// Where a kernel reads the padded cell sizes from: global memory (jo = io = 0)
// or a shared-memory copy of the window [jo, jo+n) x [io, io+m) of a tile.
struct Spacing {
    const double *hx, *ihx, *hy, *ihy;
    int jo, io;
};
__device__ __forceinline__ Spacing global_spacing(const LevelDev &L)
{
    Spacing S; S.hx = L.hx; S.ihx = L.ihx; S.hy = L.hy; S.ihy = L.ihy; S.jo = 0; S.io = 0;
    return S;
}

// Row of A = M + dt*K + dt*R at node (i,j) of a tensor-product "right" mesh,
// isotropic constant tensor (closed form of fenics/hslD.h:3123-3350 summed over
// the six triangles around the node; DESIGN.md "operator").
__device__ __forceinline__ void stencil_iso(const LevelDev &L, const Spacing &S, int i, int j, double c[NBAND])
{
    if (i >= 1 && i <= L.ireg_hi && j >= 1 && j <= L.jreg_hi) {
        c[B_C] = L.cC; c[B_E] = L.cEW; c[B_W] = L.cEW; c[B_N] = L.cNS; c[B_S] = L.cNS;
        c[B_NE] = L.cD; c[B_SW] = L.cD;
        return;
    }
    const int jj = j - S.jo, ii = i - S.io;
    const double aw = S.hx[jj], ae = S.hx[jj + 1], bs = S.hy[ii], bn = S.hy[ii + 1];
    const double iaw = S.ihx[jj], iae = S.ihx[jj + 1], ibs = S.ihy[ii], ibn = S.ihy[ii + 1];
    const double sy = 0.5 * L.tau * (bs + bn), sx = 0.5 * L.tau * (aw + ae);
    if (L.lumped) {
        // MyMatMult's row (diffuclass.cpp:786-862) times the node's cell share w*h^2 (w = 1, 1/2 on a wall,
        // 1/4 in a corner), which makes the ghost-node rows symmetric: interior (1+4F)u - F(uE+uW+uN+uS);
        // wall rows -2F to the inner neighbour, -F along the wall, Robin term 2hF(Dc/Nc) on the diagonal
        c[B_E] = -sy * iae; c[B_W] = -sy * iaw; c[B_N] = -sx * ibn; c[B_S] = -sx * ibs;
        c[B_NE] = 0.0; c[B_SW] = 0.0;
        double cl = sy * (iae + iaw) + sx * (ibn + ibs) + 0.25 * (aw + ae) * (bs + bn);
        if (j == 0) cl += L.rob_l * 0.5 * (bs + bn);
        if (j == L.nx - 1) cl += L.rob_r * 0.5 * (bs + bn);
        c[B_C] = cl;
        return;
    }
    const double my = (bs + bn) * (1.0 / 24.0), mx = (aw + ae) * (1.0 / 24.0);
    c[B_E] = ae * my - sy * iae;
    c[B_W] = aw * my - sy * iaw;
    c[B_N] = bn * mx - sx * ibn;
    c[B_S] = bs * mx - sx * ibs;
    c[B_NE] = ae * bn * (1.0 / 12.0);
    c[B_SW] = aw * bs * (1.0 / 12.0);
    double cc = sy * (iae + iaw) + sx * (ibn + ibs) +
                (2.0 * aw * bs + ae * bs + aw * bn + 2.0 * ae * bn) * (1.0 / 12.0);
    // Robin edge mass dt*r*|e|*{1/3,1/6} (fenics/hslD.h:3284-3441)
    if (j == 0 && L.rob_l != 0.0) {
        cc += L.rob_l * (bs + bn) * (1.0 / 3.0);
        c[B_N] += L.rob_l * bn * (1.0 / 6.0);
        c[B_S] += L.rob_l * bs * (1.0 / 6.0);
    }
    if (j == L.nx - 1 && L.rob_r != 0.0) {
        cc += L.rob_r * (bs + bn) * (1.0 / 3.0);
        c[B_N] += L.rob_r * bn * (1.0 / 6.0);
        c[B_S] += L.rob_r * bs * (1.0 / 6.0);
    }
    c[B_C] = cc;
}

__device__ __forceinline__ void stencil_iso(const LevelDev &L, int i, int j, double c[NBAND])
{
    stencil_iso(L, global_spacing(L), i, j, c);
}

// Consistent P1 mass row only (for the load vector b = M u0).
__device__ __forceinline__ void stencil_mass(const LevelDev &L, int i, int j, double c[NBAND])
{
    const double aw = L.hx[j], ae = L.hx[j + 1], bs = L.hy[i], bn = L.hy[i + 1];
    if (L.lumped) {   // b_i = u0_i of diffuclass.cpp (TimeStep: KSPSolve(b = globalVector)), times the cell share
        c[B_E] = 0.0; c[B_W] = 0.0; c[B_N] = 0.0; c[B_S] = 0.0; c[B_NE] = 0.0; c[B_SW] = 0.0;
        c[B_C] = 0.25 * (aw + ae) * (bs + bn);
        return;
    }
    const double my = (bs + bn) * (1.0 / 24.0), mx = (aw + ae) * (1.0 / 24.0);
    c[B_E] = ae * my; c[B_W] = aw * my; c[B_N] = bn * mx; c[B_S] = bs * mx;
    c[B_NE] = ae * bn * (1.0 / 12.0);
    c[B_SW] = aw * bs * (1.0 / 12.0);
    c[B_C] = (2.0 * aw * bs + ae * bs + aw * bn + 2.0 * ae * bn) * (1.0 / 12.0);
}

// General-tensor row: six triangles around the node, each with the mean of its
// three vertex tensor values (fenics/hslD.h:3181-3259; the 3-point rule with P1
// weights integrates a P1 coefficient exactly = area * vertex mean).
// T(i,j) reads the nodal tensor (d11,d22,d12) at a node, already times tau.
struct Ten { double a, b, c; };
__device__ __forceinline__ Ten ten_at(const LevelDev &L, int i, int j)
{
    const size_t g = (size_t)i * L.nx + j;
    Ten t; t.a = __ldg(L.d11 + g); t.b = __ldg(L.d22 + g); t.c = __ldg(L.d12 + g);
    return t;
}
__device__ __forceinline__ Ten ten_mean(const Ten &p, const Ten &q, const Ten &r, double tau)
{
    Ten t;
    t.a = tau * (p.a + q.a + r.a) * (1.0 / 3.0);
    t.b = tau * (p.b + q.b + r.b) * (1.0 / 3.0);
    t.c = tau * (p.c + q.c + r.c) * (1.0 / 3.0);
    return t;
}
__device__ __forceinline__ void stencil_tensor(const LevelDev &L, int i, int j, double c[NBAND])
{
    stencil_mass(L, i, j, c);
    const double aw = L.hx[j], ae = L.hx[j + 1], bs = L.hy[i], bn = L.hy[i + 1];
    const double iaw = L.ihx[j], iae = L.ihx[j + 1], ibs = L.ihy[i], ibn = L.ihy[i + 1];
    const bool hasW = j > 0, hasE = j < L.nx - 1, hasS = i > 0, hasN = i < L.ny - 1;
    const Ten P = ten_at(L, i, j);
    Ten E = P, W = P, N = P, S = P, NE = P, SW = P;
    if (hasE) E = ten_at(L, i, j + 1);
    if (hasW) W = ten_at(L, i, j - 1);
    if (hasN) N = ten_at(L, i + 1, j);
    if (hasS) S = ten_at(L, i - 1, j);
    if (hasN && hasE) NE = ten_at(L, i + 1, j + 1);
    if (hasS && hasW) SW = ten_at(L, i - 1, j - 1);
    double cc = 0.0;
    // K_e entries for a rectangle a x b:
    //  lower (BL,BR,TR): BL-BL (b/2a)d11; BL-BR -(b/2a)d11+d12/2; BL-TR -d12/2;
    //                    BR-BR (b/2a)d11-d12+(a/2b)d22; BR-TR d12/2-(a/2b)d22; TR-TR (a/2b)d22
    //  upper (BL,TL,TR): BL-BL (a/2b)d22; BL-TL d12/2-(a/2b)d22; BL-TR -d12/2;
    //                    TL-TL (b/2a)d11-d12+(a/2b)d22; TL-TR -(b/2a)d11+d12/2; TR-TR (b/2a)d11
    if (hasS && hasW) {  // SW cell (aw x bs), node is TR; BL=SW, BR=S, TL=W
        const double ba = 0.5 * bs * iaw, ab = 0.5 * aw * ibs;
        Ten tl = ten_mean(SW, S, P, L.tau), tu = ten_mean(SW, W, P, L.tau);
        cc += ab * tl.b + ba * tu.a;
        c[B_S] += 0.5 * tl.c - ab * tl.b;
        c[B_W] += 0.5 * tu.c - ba * tu.a;
        c[B_SW] += -0.5 * tl.c - 0.5 * tu.c;
    }
    if (hasS && hasE) {  // SE cell (ae x bs), node is TL (upper only); BL=S, TR=E
        const double ba = 0.5 * bs * iae, ab = 0.5 * ae * ibs;
        Ten tu = ten_mean(S, P, E, L.tau);
        cc += ba * tu.a - tu.c + ab * tu.b;
        c[B_S] += 0.5 * tu.c - ab * tu.b;
        c[B_E] += 0.5 * tu.c - ba * tu.a;
    }
    if (hasN && hasW) {  // NW cell (aw x bn), node is BR (lower only); BL=W, TR=N
        const double ba = 0.5 * bn * iaw, ab = 0.5 * aw * ibn;
        Ten tl = ten_mean(W, P, N, L.tau);
        cc += ba * tl.a - tl.c + ab * tl.b;
        c[B_W] += 0.5 * tl.c - ba * tl.a;
        c[B_N] += 0.5 * tl.c - ab * tl.b;
    }
    if (hasN && hasE) {  // NE cell (ae x bn), node is BL; BR=E, TL=N, TR=NE
        const double ba = 0.5 * bn * iae, ab = 0.5 * ae * ibn;
        Ten tl = ten_mean(P, E, NE, L.tau), tu = ten_mean(P, N, NE, L.tau);
        cc += ba * tl.a + ab * tu.b;
        c[B_E] += 0.5 * tl.c - ba * tl.a;
        c[B_N] += 0.5 * tu.c - ab * tu.b;
        c[B_NE] += -0.5 * tl.c - 0.5 * tu.c;
    }
    if (j == 0 && L.rob_l != 0.0) {
        cc += L.rob_l * (bs + bn) * (1.0 / 3.0);
        c[B_N] += L.rob_l * bn * (1.0 / 6.0);
        c[B_S] += L.rob_l * bs * (1.0 / 6.0);
    }
    if (j == L.nx - 1 && L.rob_r != 0.0) {
        cc += L.rob_r * (bs + bn) * (1.0 / 3.0);
        c[B_N] += L.rob_r * bn * (1.0 / 6.0);
        c[B_S] += L.rob_r * bs * (1.0 / 6.0);
    }
    c[B_C] += cc;
}

// Row of the operator.  TENSOR 0: isotropic closed form; 1: variable tensor evaluated on the fly (150 flops and 21 tensor
// loads per node); 2: variable tensor read from the assembled arrays (LevelDev::kC ..., one evaluation per node and tensor
// update instead of one per node, sweep and kernel -- and 40 registers instead of 96, i.e. four resident blocks, not two).
template <int TENSOR>
__device__ __forceinline__ void stencil_row(const LevelDev &L, int i, int j, double c[NBAND])
{
    if (TENSOR == 2) {
        const size_t g = (size_t)i * L.nx + j;
        c[B_C] = __ldg(L.kC + g);
        c[B_E] = __ldg(L.kE + g);
        c[B_N] = __ldg(L.kN + g);
        c[B_NE] = __ldg(L.kNE + g);
        c[B_W] = j > 0 ? __ldg(L.kE + g - 1) : 0.0;
        c[B_S] = i > 0 ? __ldg(L.kN + g - L.nx) : 0.0;
        c[B_SW] = (i > 0 && j > 0) ? __ldg(L.kNE + g - L.nx - 1) : 0.0;
    } else if (TENSOR == 1) {
        stencil_tensor(L, i, j, c);
    } else {
        stencil_iso(L, i, j, c);
    }
}

// sum_k c[k] * x[nbr_k]; neighbours outside the grid carry zero coefficients
// and are not read.
__device__ __forceinline__ double stencil_dot(const LevelDev &L, int i, int j,
                                              const double c[NBAND], const double *__restrict__ x)
{
    const size_t g = (size_t)i * L.nx + j;
    const bool hasW = j > 0, hasE = j < L.nx - 1, hasS = i > 0, hasN = i < L.ny - 1;
    double s = c[B_C] * __ldg(x + g);
    if (hasE) s += c[B_E] * __ldg(x + g + 1);
    if (hasW) s += c[B_W] * __ldg(x + g - 1);
    if (hasN) s += c[B_N] * __ldg(x + g + L.nx);
    if (hasS) s += c[B_S] * __ldg(x + g - L.nx);
    if (hasN && hasE) s += c[B_NE] * __ldg(x + g + L.nx + 1);
    if (hasS && hasW) s += c[B_SW] * __ldg(x + g - L.nx - 1);
    return s;
}

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Deterministic grid reduction of NV values per thread: warp shuffles, one
// shared-memory pass per block, block partials to global, and the last block
// to finish sums the partials in a fixed order.  Returns true in thread 0 of
// that last block, with the totals in out[].
template <int NV>
__device__ __forceinline__ bool grid_reduce(double (&v)[NV], double *partials, unsigned *counter,
                                            double (&out)[NV])
{
    __shared__ double sm[NV][32];
    __shared__ bool last;
    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    const int nthreads = blockDim.x * blockDim.y;
    const int lane = tid & 31, warp = tid >> 5, nwarps = (nthreads + 31) >> 5;
    const int bid = blockIdx.y * gridDim.x + blockIdx.x;
    const int nblocks = gridDim.x * gridDim.y;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        double w = warp_sum(v[k]);
        if (lane == 0) sm[k][warp] = w;
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            double w = lane < nwarps ? sm[k][lane] : 0.0;
            w = warp_sum(w);
            if (lane == 0) partials[(size_t)k * nblocks + bid] = w;
        }
    }
    if (tid == 0) {
        __threadfence();
        unsigned t = atomicAdd(counter, 1u);
        last = (t == (unsigned)nblocks - 1u);
    }
    __syncthreads();
    if (!last) return false;
    __threadfence();
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        double w = 0.0;
        for (int b = tid; b < nblocks; b += nthreads) w += __ldcg(partials + (size_t)k * nblocks + b);
        w = warp_sum(w);
        __syncthreads();
        if (lane == 0) sm[k][warp] = w;
        __syncthreads();
        if (warp == 0) {
            double q = lane < nwarps ? sm[k][lane] : 0.0;
            q = warp_sum(q);
            out[k] = q;
        }
    }
    if (tid == 0) *counter = 0u;
    return tid == 0;
}

// Reduction shared by the CTAs of SEVERAL kernels (tile-list launches): every tile of a level leaves one partial sum in
// partials[tile], whichever kernel worked on it; a CTA then reports how many tiles it has finished, and the CTA whose report
// completes the `total` tiles sums the partials in a fixed order (warp 0, lane-strided, then the shuffle tree), so the
// result does not depend on which kernel or CTA came last.  Block-level part: the sum of v over the CTA in a fixed order,
// valid in thread 0 (sm: >= 32 doubles of shared memory; contains a __syncthreads).
__device__ __forceinline__ double cta_sum(double v, double *sm)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = (blockDim.x + 31) >> 5;
    const double w = warp_sum(v);
    if (lane == 0) sm[warp] = w;
    __syncthreads();
    double tot = 0.0;
    if (tid == 0)
        for (int k = 0; k < nwarps; ++k) tot += sm[k];
    return tot;
}
// Called by all threads of a CTA after thread 0 has stored its partial(s); mine = tiles this CTA reports.
// Returns true in thread 0 of the completing CTA with the total in *out.
__device__ __forceinline__ bool tiles_arrive(double *partials, unsigned *counter, unsigned mine, unsigned total, double *out)
{
    __shared__ bool last_;
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned t = atomicAdd(counter, mine);
        last_ = (t + mine == total);
    }
    __syncthreads();
    if (!last_) return false;
    __threadfence();
    if (threadIdx.x < 32) {
        double w = 0.0;
        for (unsigned b = threadIdx.x; b < total; b += 32) w += __ldcg(partials + b);
        w = warp_sum(w);
        if (threadIdx.x == 0) { *out = w; *counter = 0u; }
    }
    return threadIdx.x == 0;
}
#endif
