/*
 * eqgpu.h -- C ABI of the B200 (sm_100a) HSL diffusion solver.
 *
 * This is the drop-in boundary for ONE hot path of jwinkle/eQ: the per-timestep
 * implicit HSL solve and the cell<->mesh coupling around it.  Every entry point
 * names the reference interface it replaces (paths relative to the reference
 * root).  The C++ host class eq_b200/host/gpuHSL.h wraps these behind eQ's own
 * `eQ::diffusionSolver` interface (src/eQ.h:302-330); INTEGRATION.md shows the
 * edits a maintainer makes in src/simulation.{h,cpp}.
 *
 * Conventions: plain pointers and sizes only; the caller owns every host
 * buffer; all fields are fp64 in natural node order g = iy*nW + jx (the
 * identity dof map, src/fHSL.h:89-114 / src/simulation.cpp:367-386); one
 * solver <-> one CUDA device and stream, no process-global state.  Every
 * function returns 0 on success or a negative EQGPU_E* code;
 * eqgpu_last_error() gives the message.  There is NO CPU fallback: without a
 * usable CUDA device eqgpu_create fails with EQGPU_ECUDA.
 */
#ifndef EQGPU_H
#define EQGPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EQGPU_ABI_VERSION 1

#if defined(__GNUC__)
#define EQGPU_API __attribute__((visibility("default")))
#else
#define EQGPU_API
#endif

enum {
    EQGPU_OK = 0,
    EQGPU_EINVAL = -1,   /* bad argument / configuration           */
    EQGPU_ECUDA = -2,    /* CUDA runtime error or no device        */
    EQGPU_ENOCONV = -3,  /* PCG hit max_iters before reaching rtol */
    EQGPU_ESTATE = -4    /* call made in the wrong state           */
};

/* Wall order everywhere: LEFT (x=0), RIGHT (x=W), TOP (y=H), BOTTOM (y=0). */
enum { EQGPU_LEFT = 0, EQGPU_RIGHT = 1, EQGPU_TOP = 2, EQGPU_BOTTOM = 3 };

/* Boundary types = the decode of parameters["boundaries"][wall][1] = {a,b,v}
 * in src/fHSL.cpp:468-539 (a==0 Dirichlet value v, v==-1 on top/bottom means
 * "the channel Function"; b==0 homogeneous Neumann; else Robin), and of the
 * legacy DIRICHLET_0 / DIRICHLET_UPDATE paths (:545-569). */
enum {
    EQGPU_BC_NEUMANN = 0,
    EQGPU_BC_DIRICHLET = 1,          /* bc_value = boundary value            */
    EQGPU_BC_ROBIN = 2,              /* bc_value = rate r (left/right only)  */
    EQGPU_BC_DIRICHLET_CHANNEL = 3   /* top/bottom only: value = channel u   */
};

/* Discretisation = which of the reference's two eQ::diffusionSolver implementations the solver stands in for:
 * fenicsInterface (P1 finite elements, fenics/hslD.ufl) or diffusionPETSc (diffuclass.cpp: 5-point finite
 * differences, unit mass, ghost-node Neumann/Robin rows `-2F u_inner ... + (1 + (4 + 2h Dc/Nc) F) u`,
 * :786-862, right-hand side u0 + 2Fh BV/Nc, :191-275).  With EQGPU_DISC_FD a wall's coefficients map as
 * Dirichlet (Nc = 0): bc_value = BV/Dc; Neumann: Dc = 0, BV = 0; Robin (left/right): bc_value = r = D*Dc/Nc,
 * robin_s = BV/Dc.  At a corner a Dirichlet wall wins, and left/right values win over top/bottom ones
 * (the write order of ApplyBoundaryConditions).  The variable tensor and row slabs are P1-only. */
enum { EQGPU_DISC_P1 = 0, EQGPU_DISC_FD = 1 };

/* What fenicsInterface::initDiffusion reads from eQ::diffusionSolver::params
 * (src/eQ.h:305-321) and from eQ::data::parameters (SURVEY.md 8b "globals"). */
typedef struct eqgpu_params {
    int32_t abi_version;      /* EQGPU_ABI_VERSION */
    int32_t nW, nH;           /* node counts = cells+1 (src/fHSL.cpp:242-243,281-283) */
    double hx, hy;            /* node spacing; hy <= 0 means hy = hx (1/nodesPerMicron) */
    double dt;                /* params.dt                                   */
    double D;                 /* params.D_HSL                                */
    int32_t bc_type[4];       /* EQGPU_BC_* per wall                         */
    double bc_value[4];       /* Dirichlet value or Robin rate per wall      */
    double robin_s[2];        /* s_left, s_right (src/fHSL.cpp:359-360)      */
    int32_t channels;         /* 1: MICROFLUIDIC_TRAP && !H_TRAP branch of stepDiffusion (:110-152) */
    int32_t channel_iters;    /* channelSolverNumberIterations (src/main.cpp:426-429) */
    double channel_v;         /* simulationFlowRate (src/fHSL.cpp:336)       */
    double channel_r[2];      /* Robin rates at the channel ends (:342-362)  */
    double well_scaling;      /* src/fHSL.cpp:47                             */
    double rtol;              /* PCG relative residual target; <=0 -> 1e-12  */
    int32_t max_iters;        /* <=0 -> 200                                  */
    int32_t device;           /* CUDA device ordinal                         */
    void *stream;             /* cudaStream_t to run on, or NULL for an own stream */
    int32_t smooth_sweeps;    /* multigrid pre/post sweeps; <=0 -> default   */
    int32_t max_levels;       /* <=0 -> automatic                            */
    int32_t discretisation;   /* EQGPU_DISC_P1 (default) or EQGPU_DISC_FD          */
    int32_t reserved[7];
} eqgpu_params;

typedef struct eqgpu_stats {
    int32_t iterations;        /* PCG iterations of the last step            */
    int32_t levels;            /* multigrid levels in use                    */
    double relres;             /* ||r|| / ||b|| reached                      */
    double total_boundary_flux;/* fenicsInterface::totalBoundaryFlux (src/fHSL.cpp:160) */
    int64_t kernel_launches;   /* kernels launched by this solver so far     */
    int64_t steps;             /* steps taken so far                         */
} eqgpu_stats;

typedef struct eqgpu_solver eqgpu_solver;

/* Fills *p with the shipped defaults (src/main.cpp:457-548, src/eQinit.h:12). */
EQGPU_API void eqgpu_default_params(eqgpu_params *p);

/* ctor + initDiffusion (src/fHSL.cpp:17-24,37-53,195-328). */
EQGPU_API int eqgpu_create(const eqgpu_params *p, eqgpu_solver **out);
/* Row-slab variant for meshes split over the GPUs of one box (BASELINE configs[4]; the reference's
 * precedent is the DMDA decomposition of diffuclass.cpp:364-370 and the per-layer sub-communicator of
 * src/simulation.cpp:657-666).  One process per GPU calls this with its rank; rank r owns a contiguous
 * block of rows, keeps halo rows per neighbour (six with the fused tile kernels), exchanges halos and sums the CG
 * scalars through peer memory over NVLink (eqgpu_comm_peer_stats) or, as fallback, over NCCL (loaded at run time
 * from the calling process; also used for the set-up).  nccl_unique_id: 128 bytes from eqgpu_nccl_unique_id
 * on rank 0, distributed by the caller.  Host field pointers always address the WHOLE nW x nH field;
 * a slab reads its window and writes back its owned rows (eqgpu_slab_rows).  Per-cell calls take the
 * full cell list on every rank and return rank-summed samples. */
EQGPU_API int eqgpu_create_slab(const eqgpu_params *p, int rank, int world, const void *nccl_unique_id,
                                eqgpu_solver **out);
EQGPU_API int eqgpu_nccl_unique_id(void *out128);
EQGPU_API int eqgpu_slab_rows(eqgpu_solver *s, int32_t *g0, int32_t *g1);
/* Host-only: owned rows [g0[l], g1[l]) of `rank` and the row count rows[l] on each multigrid level. */
EQGPU_API int eqgpu_slab_plan(int32_t nH, int32_t world, int32_t rank, int32_t max_levels, int32_t *g0, int32_t *g1,
                              int32_t *rows);
/* dtor / finalize (src/fHSL.cpp:655-661). */
EQGPU_API void eqgpu_destroy(eqgpu_solver *s);
/* Message of the last failure on this solver (s == NULL: last create failure). */
EQGPU_API const char *eqgpu_last_error(const eqgpu_solver *s);

/* solution_vector access (src/fHSL.h:412; the buffer Simulation Isend/Irecv's,
 * src/simulation.cpp:491-505).  Host <-> device copies of nW*nH doubles. */
EQGPU_API int eqgpu_set_field(eqgpu_solver *s, const double *host_u);
EQGPU_API int eqgpu_get_field(eqgpu_solver *s, double *host_u);

/* D11/D22/D12 (src/fHSL.h:434, src/simulation.cpp:503-505).  All NULL restores
 * the isotropic 1,1,0 default (src/fHSL.cpp:313-323). */
EQGPU_API int eqgpu_set_tensor(eqgpu_solver *s, const double *d11, const double *d22, const double *d12);

/* fenicsInterface::setBoundaryValues (src/fHSL.cpp:601-604): every Dirichlet
 * wall takes value v from the next step on (DIRICHLET_UPDATE). */
EQGPU_API int eqgpu_set_boundary_value(eqgpu_solver *s, double v);

/* fenicsInterface::stepDiffusion (src/fHSL.cpp:98-161) on the device-resident
 * field: trap solve, channel flux + sub-steps, boundary-flux functional. */
EQGPU_API int eqgpu_step(eqgpu_solver *s);
/* Same with the reference's host-vector contract: solution_vector in, solved
 * field out (H2D + step + D2H inside the call). */
EQGPU_API int eqgpu_step_host(eqgpu_solver *s, double *solution_vector);

EQGPU_API int eqgpu_get_stats(eqgpu_solver *s, eqgpu_stats *out);

/* topChannelData / bottomChannelData (src/fHSL.cpp:145-151), nW doubles each. */
EQGPU_API int eqgpu_get_channels(eqgpu_solver *s, double *top, double *bottom);
EQGPU_API int eqgpu_set_channels(eqgpu_solver *s, const double *top, const double *bottom);
/* fluxTopChannel / fluxBottomChannel of the last step (src/fHSL.cpp:54-96). */
EQGPU_API int eqgpu_get_channel_flux(eqgpu_solver *s, double *flux_top, double *flux_bottom);

/* ---- cells: eQabm::updateCells' lambdas (src/abm/eQabm.cpp:268-359) -------
 * A cell record is EQGPU_CELL_STRIDE doubles:
 *  0,1 bodyA position   2,3 bodyA rot (cos,sin)   4 offset   5 newOffset
 *  6 radius   7,8 polePositionA   9,10 polePositionB   11,12 centre   13 length
 * (what findInteriorPoints / pointIsInCell read: src/abm/cpmEcoli.cpp:313-327,
 *  src/abm/Ecoli.cpp:36-63). */
#define EQGPU_CELL_STRIDE 16
EQGPU_API int eqgpu_cells_upload(eqgpu_solver *s, const double *records, int64_t ncells, double nodes_per_micron);
/* Same, from records that already sit in DEVICE memory (a controller that stages several steps of records in HBM,
 * or builds them on the device): a stream-ordered device-to-device copy, no host synchronisation. */
EQGPU_API int eqgpu_cells_upload_device(eqgpu_solver *s, const double *d_records, int64_t ncells, double nodes_per_micron);
/* findInteriorPoints: counts[k] points of cell k, node ids (iy*nW+jx) in the
 * reference's push_back order into nodes[k*cap .. ] (at most cap written). */
EQGPU_API int eqgpu_cells_raster(eqgpu_solver *s, int32_t *counts, int64_t *nodes, int32_t cap);
/* readHSL for every cell: out[k] = mean of the field over the cell's points. */
EQGPU_API int eqgpu_cells_gather(eqgpu_solver *s, double *out);
/* writeHSL for every cell: amount_nM[k] is Strain's deltaHSL for this layer. */
EQGPU_API int eqgpu_cells_scatter(eqgpu_solver *s, const double *amount_nM);

/* setDiffusionTensor for every cell (src/abm/eQabm.cpp:246-248,306-325,407): the D11/D22/D12 fields are
 * reset to 1,1,0 and each rod writes Dx c^2 + Dy s^2, Dx s^2 + Dy c^2, (Dx - Dy) s c on its points (Dx, Dy =
 * parameters["AnisotropicDiffusion_Axial" / "_Transverse"]); a later record wins a shared node, as the
 * reference's list order does.  The result becomes the solver's tensor, as after eqgpu_set_tensor (with
 * Dx == Dy == 1, the shipped values, the solve stays on the constant-coefficient kernels: the fields
 * then differ from 1,1,0 by rounding only).  eqgpu_get_tensor copies the three fields to the host. */
EQGPU_API int eqgpu_cells_tensor(eqgpu_solver *s, double Dx, double Dy);
EQGPU_API int eqgpu_get_tensor(eqgpu_solver *s, double *d11, double *d22, double *d12);

/* writeHSL strategy: 0 = one global fp64 atomic per (rod, node) (default; exact for non-overlapping rods),
 * 1 = rods binned by 64x64-node tile and accumulated with shared-memory atomics before one global atomic
 * per touched node (dense / overlapping colonies). */
EQGPU_API int eqgpu_set_scatter_mode(eqgpu_solver *s, int mode);

/* Device-resident variants (no host copy inside): gather into the solver's
 * per-cell buffer / scatter the amounts last given to eqgpu_cells_set_amounts.
 * eqgpu_cells_get_gathered copies the last gather result to the host. */
EQGPU_API int eqgpu_cells_set_amounts(eqgpu_solver *s, const double *amount_nM);
EQGPU_API int eqgpu_cells_gather_resident(eqgpu_solver *s);
EQGPU_API int eqgpu_cells_scatter_resident(eqgpu_solver *s);
EQGPU_API int eqgpu_cells_get_gathered(eqgpu_solver *s, double *out);

/* ---- verification hooks (used by the parity tests) ------------------------
 * y = A x with A = M + dt*K(D) + dt*R exactly as DOLFIN would assemble it
 * (unconstrained, constrained = 0) or with Dirichlet rows replaced by identity
 * and columns eliminated (constrained = 1). */
EQGPU_API int eqgpu_apply_operator(eqgpu_solver *s, const double *host_x, double *host_y, int constrained);
/* b = L(u0) = M u0 + Robin load (unconstrained load vector). */
EQGPU_API int eqgpu_build_rhs(eqgpu_solver *s, const double *host_u0, double *host_b);
/* Device pointer to the resident field (for benchmarks that stage inputs in HBM). */
EQGPU_API int eqgpu_field_device_ptr(eqgpu_solver *s, void **dev_ptr);
EQGPU_API int eqgpu_sync(eqgpu_solver *s);
/* What a step does when PCG stops at max_iters above rtol.  policy 0 (default): the step returns EQGPU_ENOCONV (the field
 * holds the best iterate).  policy 1: report and continue -- the step returns EQGPU_OK, stats.relres tells how far it got,
 * and the step is counted in eqgpu_unconverged_steps.  The reference prints solver diagnostics and carries on
 * (src/fHSL.cpp:104-108 has no error path); policy 1 is that behaviour for 24 000-step runs that must not abort. */
EQGPU_API int eqgpu_set_nonconvergence_policy(eqgpu_solver *s, int policy);
EQGPU_API int eqgpu_unconverged_steps(eqgpu_solver *s, int64_t *count);
/* Which code path the solver selected (diagnostics / tests): bit 0 fused tile kernels, 1 row-slab mode,
 * 2 fused kernels in slab mode, 3 cluster tail, 4 tiled coarsest solve, 5 variable tensor active; bits 8-11: number of
 * multigrid levels whose interior tiles run on the register-tile smoothers (TMA-staged, smooth_rt.cu). */
EQGPU_API int eqgpu_solver_path(eqgpu_solver *s);
/* Starting guess of the iterative solve that stands in for LinearVariationalSolver::solve()
 * (src/fHSL.cpp:106; the reference's LU has no such notion).  mode 0: the field as given or zero, whichever
 * has the smaller residual; 1: also the previous step's solution; 2: also the linear extrapolation of the
 * two previous solutions; 3: also the quadratic extrapolation of the last three; 4: also the
 * residual-minimising combination of the last three solutions (a 3x3 least-squares problem solved on the
 * device; its span contains the previous solution and both extrapolations).  5: mode 3 plus the cubic
 * extrapolation of the last four solutions; 6: mode 5 plus the quartic extrapolation of the last five.
 * 7 (opt-in; first run on a B200 in the last seconds of round 1, profiles/r01_ring_first_run.json: 1.46 iterations
 * per step against mode 6's 2.66 on the 2048^2 bench, same field): image ring -- the last seven
 * solutions and their images, guess = fixed extrapolation plus a least-squares correction in the
 * backward-difference basis (profiles/r01_guess_study.md).
 * Default: 4 for meshes up to 512^2 nodes, 6 above (measured: the
 * least-squares combination saves up to four iterations per step on small or quasi-steady problems and costs
 * half an iteration at 2048^2).  The stopping test is relative to the right-hand side in every mode, so
 * the mode changes the iteration count, not the accuracy.  History lives on the device, survives
 * eqgpu_set_field, and is kept on the single-GPU isotropic path only (elsewhere the call is accepted and
 * mode 0 is what runs). */
EQGPU_API int eqgpu_set_warm_start(eqgpu_solver *s, int mode);
/* Row-slab mode: cumulative communication counters of this rank since creation -- out[0] all-reduce calls,
 * out[1] doubles all-reduced, out[2] halo exchanges (one NCCL group each), out[3] halo bytes sent.  Zeros on one GPU. */
EQGPU_API int eqgpu_comm_stats(eqgpu_solver *s, int64_t out[4]);
/* Row-slab mode, transport of the data path.  By default the ranks of one box map each other's vectors (CUDA IPC over
 * NVLink / NVSwitch) and a halo exchange is one small kernel that stores my boundary rows into the neighbours' staging buffers and copies theirs out of mine, a scalar
 * all-reduce one warp that sums every rank's partials in rank order -- no NCCL call per exchange; NCCL remains the
 * transport of the set-up, of long reductions (per-cell samples) and of everything when peer mapping is unavailable or
 * EQGPU_SLAB_PEER=0.  out[0] halo-exchange kernels, out[1] peer all-reduces so far; -1, -1 when NCCL carries everything.
 * The reference's counterpart is PETSc's DMDA ghost update and the KSP's MPI_Allreduce (diffuclass.cpp:364-370,410). */
EQGPU_API int eqgpu_comm_peer_stats(eqgpu_solver *s, int64_t out[2]);
/* Verification hook (single GPU, isotropic operator): z = B r, one application of the multigrid preconditioner
 * (the V-cycle of the PCG iteration) to a host vector r (zero on Dirichlet rows); host arrays of nW*nH doubles.
 * Lets the tests compare the streaming smoothers with the shared-memory tile kernels and check the symmetry of B. */
EQGPU_API int eqgpu_apply_preconditioner(eqgpu_solver *s, const double *r, double *z);
/* The mode in effect (the size-dependent default, EQGPU_WARM, or the last eqgpu_set_warm_start). */
EQGPU_API int eqgpu_get_warm_start(eqgpu_solver *s);
/* Which guess the last step started from: 0 field as given, 1 zero, 2 previous solution, 3 linear,
 * 4 quadratic extrapolation, 5 least-squares combination, 6 cubic, 7 quartic extrapolation, 8 image-ring guess
 * (mode 7). */
EQGPU_API int eqgpu_last_guess(eqgpu_solver *s);
/* Host-only (no device): the 3x3 least-squares solve warm-start mode 4 runs on the device, for the CPU tests.
 * G = {a0.a0, a0.a1, a0.a2, a1.a1, a1.a2, a2.a2}, f = {a0.b, a1.b, a2.b}, bb = b.b; c minimises
 * ||b - c0 a0 - c1 a1 - c2 a2||, *pred is the predicted squared residual (1e300 if the solve failed). */
EQGPU_API int eqgpu_ls_solve3(const double *G, const double *f, double bb, double *c, double *pred);
/* Host-only (no device): the K x K (K <= 7) normal-equation solve of warm-start mode 7 (image ring: least-squares
 * correction to the fixed extrapolation in the backward-difference basis; opt-in).
 * G = packed upper triangle of the Gram matrix, row-major (K(K+1)/2 entries), f = its right-hand side; c = the
 * correction coefficients (all zero if the solve failed: the fixed extrapolation stands). */
EQGPU_API int eqgpu_ring_solve(int K, const double *G, const double *f, double *c);
/* Times `reps` back-to-back launches of one named kernel on the solver's
 * stream with CUDA events (for bench.py's roofline line).  Returns the average
 * milliseconds per launch and the algorithmic bytes one launch must move
 * (DESIGN.md "kernels").  Unknown name -> EQGPU_EINVAL. */
EQGPU_API int eqgpu_bench_kernel(eqgpu_solver *s, const char *name, int reps, double *avg_ms, double *alg_bytes);

#ifdef __cplusplus
}
#endif
#endif /* EQGPU_H */
