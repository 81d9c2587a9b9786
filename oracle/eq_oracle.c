/*
 * eq_oracle.c -- CPU restatement of eQ's HSL diffusion hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ may be imported, linked or
 * executed by the product path (eq_b200/).  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs use it, as the checker.
 *
 * PARITY PINNING: the reference (jwinkle/eQ) ships no tests, golden vectors or
 * fixtures for this path, and DOLFIN/PETSc/Chipmunk are
 * absent, so the reference executable cannot be built here.  What IS pinned: the element
 * kernels below are checked bit-for-bit against the reference's own
 * FFC-generated tabulate_tensor bodies, compiled in place from
 * /root/reference/fenics/{hslD,AdvectionDiffusion,boundary}.h under a UFC/DOLFIN
 * shim (oracle/Makefile -> oracle/_ref/libeq_ufc_ref.so; tests/test_oracle_ref.py),
 * and golden vectors produced by those compiled kernels are committed under
 * tests/golden/.
 * The P1 step IS pinned against the reference's own class: fenicsInterface
 * (src/fHSL.{h,cpp}, src/Expressions.h and every generated form header it
 * includes) is compiled in place on a one-process DOLFIN interface shim
 * (oracle/shim_dolfin/, oracle/fenics_ref.cpp -> oracle/_ref/libeq_fenics_ref.so)
 * and initDiffusion / createHSL / setRobinBoundaryConditions / stepDiffusion /
 * computeBoundaryFlux / setBoundaryValues themselves run on the reference's own
 * element kernels; oracle.py's problem_from_parameters + step (eqo_assemble,
 * eqo_dirichlet_mask, eqo_compute_boundary_flux, eqo_channel_substeps,
 * eqo_boundary_functional below, SuperLU) agree with it to 1e-11 on every
 * boundary / trap type the reference decodes, the anisotropic tensor and three
 * geometries (28 golden cases, tests/golden/fenics_ref.json).  What stays
 * restated there is DOLFIN 2019.1.0 [ext] -- RectangleMesh("right") /
 * IntervalMesh numbering, the assembly loop, DirichletBC::apply (identity
 * rows, list order) and the direct solve -- now inside the shim, written from
 * DOLFIN's documented behaviour: the reference does not vendor DOLFIN, no
 * DOLFIN build exists in this image, and the reference ships no golden vectors
 * (SURVEY.md section 4), so that layer is "parity unpinned" in the strict sense.
 * The finite-difference solver IS pinned end to end: the reference's own class
 * diffusionPETSc (diffuclass.{h,cpp}) is compiled in place on a one-process
 * PETSc/MPI/boost interface shim (oracle/shim_petsc/, oracle/petsc_ref.cpp ->
 * oracle/_ref/libeq_fd_ref.so) and run; eqo_fd_matmult / eqo_fd_apply_bc below
 * equal its MyMatMult / ApplyBoundaryConditions bit for bit and its
 * stepDiffusion agrees with the exact solve of the restated system to 1e-13
 * (tests/test_oracle_golden.py, golden vectors tests/golden/fd_ref.json).  Only
 * PETSc's Krylov solver [ext] is replaced there (plain BiCGStab).
 * The cell <-> mesh coupling IS pinned too: the reference's own eQabm, Ecoli,
 * cpmEcoli and Strain classes are compiled in place on a Chipmunk 7.0.1
 * interface shim (oracle/shim_cpm/, oracle/cell_ref.cpp ->
 * oracle/_ref/libeq_cell_ref.so) and eQabm::updateCells itself is run:
 * eqo_make_cell, eqo_point_in_cell, eqo_update_cells_sequential and
 * eqo_cells_tensor equal it bit for bit on fresh, grown, bent and ratcheted
 * rods (golden vectors tests/golden/cells_ref.json).  What stays restated
 * there is Chipmunk's rigid-body transform arithmetic [ext], now in the shim.
 *
 * Every function cites the reference file:line it follows (paths relative to
 * the reference root).  All arithmetic is fp64; compile WITHOUT -ffast-math and
 * with -ffp-contract=off so the operation order below is what executes.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define EQO_API __attribute__((visibility("default")))

/* UFC reference-cell tables (ufc_geometry.h, UFC 2018.1.0 [ext]); facet f is
 * the edge opposite local vertex f of the reference triangle. */
static const double tri_ref_facet_jac[3][2][1] = {
    {{-1.0}, {1.0}}, {{0.0}, {1.0}}, {{1.0}, {0.0}}};
static const double tri_ref_facet_normals[3][2] = {
    {0.7071067811865476, 0.7071067811865476}, {-1.0, 0.0}, {0.0, -1.0}};

/* ------------------------------------------------------------------------- */
/* Element kernels (restated from the FFC-generated code, same op order)      */
/* ------------------------------------------------------------------------- */

/* fenics/hslD.h:3123-3259  hsld_cell_integral_0_otherwise::tabulate_tensor
 * A_e = |det| * mass + dt * |det| * w_q * grad(phi_i) . (D*Dq) grad(phi_j)
 * w = {D11[3], D22[3], D12[3], D, dt}. */
EQO_API void eqo_hsld_cell_a(double *A, const double *d11, const double *d22,
                             const double *d12, double D, double dt,
                             const double *xy)
{
    static const double weights3[3] = {0.1666666666666667, 0.1666666666666667,
                                       0.1666666666666667};
    static const double dphi[2] = {-1.0, 1.0};
    static const double FE[3][3] = {
        {0.6666666666666669, 0.1666666666666666, 0.1666666666666667},
        {0.1666666666666667, 0.1666666666666666, 0.6666666666666665},
        {0.1666666666666667, 0.6666666666666666, 0.1666666666666666}};
    const double J_c0 = xy[0] * dphi[0] + xy[2] * dphi[1];
    const double J_c3 = xy[1] * dphi[0] + xy[5] * dphi[1];
    const double J_c1 = xy[0] * dphi[0] + xy[4] * dphi[1];
    const double J_c2 = xy[1] * dphi[0] + xy[3] * dphi[1];
    double sp[8];
    sp[0] = J_c0 * J_c3;
    sp[1] = J_c1 * J_c2;
    sp[2] = sp[0] + -1 * sp[1];
    sp[3] = J_c0 / sp[2];
    sp[4] = -1 * J_c1 / sp[2];
    sp[5] = J_c3 / sp[2];
    sp[6] = -1 * J_c2 / sp[2];
    sp[7] = fabs(sp[2]);
    double TP0[2] = {0, 0}, TP1[2] = {0, 0}, TP2[2] = {0, 0}, TP3[2] = {0, 0};
    for (int iq = 0; iq < 3; ++iq) {
        double w1 = 0.0, w2 = 0.0, w0 = 0.0;
        for (int ic = 0; ic < 3; ++ic) w1 += d22[ic] * FE[iq][ic];
        for (int ic = 0; ic < 3; ++ic) w2 += d12[ic] * FE[iq][ic];
        for (int ic = 0; ic < 3; ++ic) w0 += d11[ic] * FE[iq][ic];
        double sv[35];
        sv[0] = w1 * D;
        sv[1] = sv[0] * sp[3];
        sv[2] = sv[0] * sp[4];
        sv[3] = w2 * D;
        sv[4] = sv[3] * sp[6];
        sv[5] = sv[3] * sp[5];
        sv[6] = sv[1] + sv[4];
        sv[7] = sv[5] + sv[2];
        sv[8] = sv[6] * sp[3];
        sv[9] = sv[6] * sp[4];
        sv[10] = sv[7] * sp[3];
        sv[11] = sv[7] * sp[4];
        sv[12] = w0 * D;
        sv[13] = sv[12] * sp[6];
        sv[14] = sv[12] * sp[5];
        sv[15] = sv[3] * sp[3];
        sv[16] = sv[3] * sp[4];
        sv[17] = sv[15] + sv[13];
        sv[18] = sv[14] + sv[16];
        sv[19] = sv[17] * sp[6];
        sv[20] = sv[17] * sp[5];
        sv[21] = sv[18] * sp[6];
        sv[22] = sv[18] * sp[5];
        sv[23] = sv[8] + sv[19];
        sv[24] = sv[20] + sv[9];
        sv[25] = sv[10] + sv[21];
        sv[26] = sv[22] + sv[11];
        sv[27] = sv[23] * dt;
        sv[28] = sv[24] * dt;
        sv[29] = sv[25] * dt;
        sv[30] = sv[26] * dt;
        sv[31] = sv[27] * sp[7];
        sv[32] = sv[28] * sp[7];
        sv[33] = sv[29] * sp[7];
        sv[34] = sv[30] * sp[7];
        const double fw0 = sv[34] * weights3[iq];
        for (int j = 0; j < 2; ++j) TP0[j] += fw0 * dphi[j];
        const double fw1 = sv[32] * weights3[iq];
        for (int j = 0; j < 2; ++j) TP1[j] += fw1 * dphi[j];
        const double fw2 = sv[33] * weights3[iq];
        for (int j = 0; j < 2; ++j) TP2[j] += fw2 * dphi[j];
        const double fw3 = sv[31] * weights3[iq];
        for (int j = 0; j < 2; ++j) TP3[j] += fw3 * dphi[j];
    }
    A[0] = 0.08333333333333338 * sp[7];
    A[1] = 0.04166666666666666 * sp[7];
    A[2] = 0.04166666666666667 * sp[7];
    A[3] = 0.04166666666666666 * sp[7];
    A[4] = 0.08333333333333333 * sp[7];
    A[5] = 0.04166666666666665 * sp[7];
    A[6] = 0.04166666666666667 * sp[7];
    A[7] = 0.04166666666666665 * sp[7];
    A[8] = 0.08333333333333329 * sp[7];
    static const int DM0[2] = {0, 2};
    for (int i = 0; i < 2; ++i)
        for (int j = 0; j < 2; ++j) A[3 * i + j] += dphi[i] * TP0[j];
    for (int i = 0; i < 2; ++i)
        for (int j = 0; j < 2; ++j) A[3 * i + DM0[j]] += dphi[i] * TP1[j];
    for (int i = 0; i < 2; ++i)
        for (int j = 0; j < 2; ++j) A[3 * DM0[i] + j] += dphi[i] * TP2[j];
    for (int i = 0; i < 2; ++i)
        for (int j = 0; j < 2; ++j) A[3 * DM0[i] + DM0[j]] += dphi[i] * TP3[j];
}

/* Facet length as the generated code computes it (hslD.h:3312-3333). */
static double facet_len(const double *xy, int facet)
{
    const double J_c0 = xy[0] * -1.0 + xy[2] * 1.0;
    const double J_c1 = xy[0] * -1.0 + xy[4] * 1.0;
    const double J_c2 = xy[1] * -1.0 + xy[3] * 1.0;
    const double J_c3 = xy[1] * -1.0 + xy[5] * 1.0;
    double a = J_c0 * tri_ref_facet_jac[facet][0][0];
    double b = J_c1 * tri_ref_facet_jac[facet][1][0];
    double c = a + b;
    double d = c * c;
    double e = tri_ref_facet_jac[facet][0][0] * J_c2;
    double f = tri_ref_facet_jac[facet][1][0] * J_c3;
    double g = e + f;
    double hh = g * g;
    return sqrt(d + hh);
}

/* fenics/hslD.h:3284-3350 (ds(1), rate r1) and :3375-3441 (ds(2), rate r2):
 * Robin edge matrix dt*r*|e|*PI0[facet]. */
EQO_API void eqo_hsld_facet_a(double *A, double dt, double r, const double *xy,
                              int facet)
{
    static const double PI0[3][3][3] = {
        {{0.0, 0.0, 0.0},
         {0.0, 0.3333333333333334, 0.1666666666666667},
         {0.0, 0.1666666666666667, 0.3333333333333334}},
        {{0.3333333333333334, 0.0, 0.1666666666666667},
         {0.0, 0.0, 0.0},
         {0.1666666666666667, 0.0, 0.3333333333333334}},
        {{0.3333333333333334, 0.1666666666666667, 0.0},
         {0.1666666666666667, 0.3333333333333334, 0.0},
         {0.0, 0.0, 0.0}}};
    const double sp0 = dt * r;
    const double sp11 = sp0 * facet_len(xy, facet);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) A[3 * i + j] = sp11 * PI0[facet][i][j];
}

/* fenics/hslD.h:3466-3529  hsld_cell_integral_1_otherwise:
 * b_e[i] = sum_q w_q |det| (dt*f + u0(q)) phi_i(q). */
EQO_API void eqo_hsld_cell_L(double *b, const double *u0, double dt, double f,
                             const double *xy)
{
    static const double weights3[3] = {0.1666666666666667, 0.1666666666666667,
                                       0.1666666666666667};
    static const double FE[3][3] = {
        {0.6666666666666669, 0.1666666666666666, 0.1666666666666667},
        {0.1666666666666667, 0.1666666666666666, 0.6666666666666665},
        {0.1666666666666667, 0.6666666666666666, 0.1666666666666666}};
    const double J_c0 = xy[0] * -1.0 + xy[2] * 1.0;
    const double J_c3 = xy[1] * -1.0 + xy[5] * 1.0;
    const double J_c1 = xy[0] * -1.0 + xy[4] * 1.0;
    const double J_c2 = xy[1] * -1.0 + xy[3] * 1.0;
    double sp[5];
    sp[0] = dt * f;
    sp[1] = J_c0 * J_c3;
    sp[2] = J_c1 * J_c2;
    sp[3] = sp[1] + -1 * sp[2];
    sp[4] = fabs(sp[3]);
    double BF0[3] = {0, 0, 0};
    for (int iq = 0; iq < 3; ++iq) {
        double w0 = 0.0;
        for (int ic = 0; ic < 3; ++ic) w0 += u0[ic] * FE[iq][ic];
        double sv0 = -1 * (-1 * sp[0]) + -1 * (-1 * w0);
        double sv1 = sv0 * sp[4];
        const double fw0 = sv1 * weights3[iq];
        for (int i = 0; i < 3; ++i) BF0[i] += fw0 * FE[iq][i];
    }
    for (int i = 0; i < 3; ++i) b[i] = 0.0 + BF0[i];
}

/* fenics/hslD.h:3554-3609 / :3634-3689: Robin load dt*r*s*|e|*PI0[facet]. */
EQO_API void eqo_hsld_facet_L(double *b, double dt, double r, double s,
                              const double *xy, int facet)
{
    static const double PI0[3][3] = {
        {0.0, 0.5, 0.5}, {0.5, 0.0, 0.5}, {0.5, 0.5, 0.0}};
    double sp0 = dt * r;
    double sp1 = sp0 * (-1 * s);
    double sp12 = -1 * sp1 * facet_len(xy, facet);
    for (int i = 0; i < 3; ++i) b[i] = sp12 * PI0[facet][i];
}

/* fenics/boundary.h:2652-2741: -grad(u).n |e| on one exterior facet (P1 u). */
EQO_API double eqo_boundary_facet(const double *u, const double *xy, int facet)
{
    const double w0_d1 = u[0] * -1.0 + u[2] * 1.0;
    const double J_c0 = xy[0] * -1.0 + xy[2] * 1.0;
    const double J_c3 = xy[1] * -1.0 + xy[5] * 1.0;
    const double J_c1 = xy[0] * -1.0 + xy[4] * 1.0;
    const double J_c2 = xy[1] * -1.0 + xy[3] * 1.0;
    double w0_d0 = 0.0;
    w0_d0 += u[0] * -1.0;
    w0_d0 += u[1] * 1.0;
    double sp[39];
    sp[0] = J_c0 * J_c3;
    sp[1] = J_c1 * J_c2;
    sp[2] = sp[0] + -1 * sp[1];
    sp[3] = J_c0 / sp[2];
    sp[4] = w0_d1 * sp[3];
    sp[5] = -1 * J_c1 / sp[2];
    sp[6] = w0_d0 * sp[5];
    sp[7] = sp[4] + sp[6];
    sp[8] = tri_ref_facet_normals[facet][1] * sp[3];
    sp[9] = tri_ref_facet_normals[facet][0] * sp[5];
    sp[10] = sp[8] + sp[9];
    sp[11] = sp[10] * sp[10];
    sp[12] = J_c3 / sp[2];
    sp[13] = tri_ref_facet_normals[facet][0] * sp[12];
    sp[14] = -1 * J_c2 / sp[2];
    sp[15] = tri_ref_facet_normals[facet][1] * sp[14];
    sp[16] = sp[13] + sp[15];
    sp[17] = sp[16] * sp[16];
    sp[18] = sp[11] + sp[17];
    sp[19] = sqrt(sp[18]);
    sp[20] = sp[10] / sp[19];
    sp[21] = sp[7] * sp[20];
    sp[22] = w0_d0 * sp[12];
    sp[23] = w0_d1 * sp[14];
    sp[24] = sp[22] + sp[23];
    sp[25] = sp[16] / sp[19];
    sp[26] = sp[24] * sp[25];
    sp[27] = sp[21] + sp[26];
    sp[38] = -1 * sp[27] * facet_len(xy, facet);
    return 0.0 + sp[38] * 1.0;
}

/* fenics/AdvectionDiffusion.h:2246-2289: 1-D CN element matrix (2x2).
 * w = {dt, D, v}; xc = {x0, x1}. */
EQO_API void eqo_ad_cell_a(double *A, double dt, double D, double v,
                           const double *xc)
{
    const double J_c0 = xc[0] * -1.0 + xc[1] * 1.0;
    double sp[8];
    sp[0] = dt * D;
    sp[1] = sp[0] * (1.0 / J_c0);
    sp[2] = 0.5 * (1.0 / J_c0) * sp[1];
    sp[3] = dt * v;
    sp[4] = 0.5 * (1.0 / J_c0) * sp[3];
    sp[5] = fabs(J_c0);
    sp[6] = sp[2] * sp[5];
    sp[7] = sp[4] * sp[5];
    A[0] = 0.3333333333333334 * sp[5] + -0.5 * sp[7] + sp[6];
    A[1] = 0.1666666666666667 * sp[5] + 0.5 * sp[7] - sp[6];
    A[2] = 0.1666666666666667 * sp[5] + -0.5 * sp[7] - sp[6];
    A[3] = 0.3333333333333334 * sp[5] + 0.5 * sp[7] + sp[6];
}

/* fenics/AdvectionDiffusion.h:2314-2354 / :2379-2419: end-point Robin matrix
 * 0.5*dt*r on the facet vertex (facet f = vertex f of the interval). */
EQO_API void eqo_ad_facet_a(double *A, double dt, double r, int facet)
{
    static const double PI0[2][2][2] = {{{1.0, 0.0}, {0.0, 0.0}},
                                        {{0.0, 0.0}, {0.0, 1.0}}};
    double sp0 = dt * r;
    A[0] = 0.5 * sp0 * PI0[facet][0][0];
    A[1] = 0.5 * sp0 * PI0[facet][0][1];
    A[2] = 0.5 * sp0 * PI0[facet][1][0];
    A[3] = 0.5 * sp0 * PI0[facet][1][1];
}

/* fenics/AdvectionDiffusion.h:2444-2510: CN element load.
 * w = {u0[2], dt, D, v}. */
EQO_API void eqo_ad_cell_L(double *b, const double *u0, double dt, double D,
                           double v, const double *xc)
{
    static const double weights2[2] = {0.5, 0.5};
    static const double FEQ[2][2] = {{0.7886751345948129, 0.2113248654051871},
                                     {0.2113248654051871, 0.7886751345948129}};
    double w0_d0 = 0.0;
    w0_d0 += u0[0] * -1.0;
    w0_d0 += u0[1] * 1.0;
    const double J_c0 = xc[0] * -1.0 + xc[1] * 1.0;
    double sp[8];
    sp[0] = w0_d0 * (1.0 / J_c0);
    sp[1] = dt * D;
    sp[2] = sp[1] * (1.0 / J_c0);
    sp[3] = 0.5 * sp[0] * sp[2];
    sp[4] = dt * v;
    sp[5] = 0.5 * sp[0] * sp[4];
    sp[6] = fabs(J_c0);
    sp[7] = -1 * sp[3] * sp[6];
    double BF0[2] = {0, 0};
    for (int iq = 0; iq < 2; ++iq) {
        double w0 = 0.0;
        for (int ic = 0; ic < 2; ++ic) w0 += u0[ic] * FEQ[iq][ic];
        double sv0 = -1 * sp[5] + -1 * (-1 * w0);
        double sv1 = sv0 * sp[6];
        const double fw0 = sv1 * weights2[iq];
        for (int i = 0; i < 2; ++i) BF0[i] += fw0 * FEQ[iq][i];
    }
    b[0] = -sp[7];
    b[1] = sp[7];
    for (int i = 0; i < 2; ++i) b[i] += BF0[i];
}

/* fenics/AdvectionDiffusion.h:2535-2579 / :2604-2648: end-point Robin load
 * -dt*r*(0.5*u0 - s) on the facet vertex. */
EQO_API void eqo_ad_facet_L(double *b, const double *u0, double dt, double r,
                            double s, int facet)
{
    static const double FEF[2][2] = {{1.0, 0.0}, {0.0, 1.0}};
    double w0 = 0.0;
    for (int ic = 0; ic < 2; ++ic) w0 += u0[ic] * FEF[facet][ic];
    double sp0 = 0.5 * w0 + -1 * s;
    double sp1 = dt * r;
    double sp2 = sp0 * sp1;
    b[0] = -1 * sp2 * FEF[facet][0];
    b[1] = -1 * sp2 * FEF[facet][1];
}

/* ------------------------------------------------------------------------- */
/* Mesh + assembly (DOLFIN RectangleMesh "right" + Assembler, [ext])          */
/* ------------------------------------------------------------------------- */

/* Band order used everywhere: offsets of the 7-point P1 stencil on the
 * "right"-diagonal mesh in natural ordering g = iy*nW + jx. */
enum { B_C = 0, B_E, B_W, B_N, B_S, B_NE, B_SW, NBAND };

static inline int band_of(long diff, long nW)
{
    if (diff == 0) return B_C;
    if (diff == 1) return B_E;
    if (diff == -1) return B_W;
    if (diff == nW) return B_N;
    if (diff == -nW) return B_S;
    if (diff == nW + 1) return B_NE;
    if (diff == -nW - 1) return B_SW;
    return -1;
}

EQO_API void eqo_band_offsets(long nW, long *off)
{
    off[B_C] = 0; off[B_E] = 1; off[B_W] = -1; off[B_N] = nW; off[B_S] = -nW;
    off[B_NE] = nW + 1; off[B_SW] = -nW - 1;
}

/* Vertex coordinates as DOLFIN's RectangleMesh::build lays them out
 * (src/fHSL.cpp:164-172; row-major, x fastest) [ext]. */
static inline void vertex_xy(long v, long nW, long nH, double W, double H,
                             double *x, double *y)
{
    long ix = v % nW, iy = v / nW;
    *x = 0.0 + ((double)ix) * (W - 0.0) / (double)(nW - 1);
    *y = 0.0 + ((double)iy) * (H - 0.0) / (double)(nH - 1);
}

/* Assemble the un-constrained system of fenics/hslD.ufl:36-42 the way
 * LinearVariationalSolver::solve does each step (src/fHSL.cpp:104-106):
 *   bands[k*N + g] = A[g][g+off_k],   b[g] = L(phi_g).
 * d11/d22/d12 are nodal tensor fields (src/Expressions.h:112-118 evaluates to
 * the nodal value at every vertex); NULL means 1,1,0 (src/fHSL.cpp:313-323).
 * use_robin mirrors "meshFunction set" (src/fHSL.cpp:455-457, fHSL.h:215-216):
 * left wall = ds(1) with (r1,s1), right wall = ds(2) with (r2,s2). */
EQO_API int eqo_assemble(long nW, long nH, double W, double H, double D,
                         double dt, const double *d11, const double *d22,
                         const double *d12, int use_robin, double r1, double s1,
                         double r2, double s2, const double *u0, double f,
                         double *bands, double *b)
{
    const long N = nW * nH;
    if (bands) memset(bands, 0, sizeof(double) * NBAND * N);
    if (b) memset(b, 0, sizeof(double) * N);
    for (long cy = 0; cy < nH - 1; ++cy) {
        for (long cx = 0; cx < nW - 1; ++cx) {
            const long v0 = cy * nW + cx, v1 = v0 + 1, v2 = v0 + nW, v3 = v2 + 1;
            /* "right": diagonal v0-v3; cells (v0,v1,v3) and (v0,v2,v3) */
            for (int t = 0; t < 2; ++t) {
                long vs[3] = {v0, t == 0 ? v1 : v2, v3};
                double xy[6], a11[3], a22[3], a12[3], uu[3];
                for (int k = 0; k < 3; ++k) {
                    vertex_xy(vs[k], nW, nH, W, H, &xy[2 * k], &xy[2 * k + 1]);
                    a11[k] = d11 ? d11[vs[k]] : 1.0;
                    a22[k] = d22 ? d22[vs[k]] : 1.0;
                    a12[k] = d12 ? d12[vs[k]] : 0.0;
                    uu[k] = u0 ? u0[vs[k]] : 0.0;
                }
                if (bands) {
                    double Ae[9];
                    eqo_hsld_cell_a(Ae, a11, a22, a12, D, dt, xy);
                    for (int i = 0; i < 3; ++i)
                        for (int j = 0; j < 3; ++j) {
                            int k = band_of(vs[j] - vs[i], nW);
                            if (k < 0) return -1;
                            bands[(size_t)k * N + vs[i]] += Ae[3 * i + j];
                        }
                }
                if (b) {
                    double be[3];
                    eqo_hsld_cell_L(be, uu, dt, f, xy);
                    for (int i = 0; i < 3; ++i) b[vs[i]] += be[i];
                }
                if (use_robin) {
                    /* left wall: upper triangle edge v0-v2 = local (0,1) -> facet 2
                     * right wall: lower triangle edge v1-v3 = local (1,2) -> facet 0 */
                    int facet = -1;
                    double r = 0, s = 0;
                    if (t == 1 && cx == 0) { facet = 2; r = r1; s = s1; }
                    if (t == 0 && cx == nW - 2) { facet = 0; r = r2; s = s2; }
                    if (facet >= 0) {
                        if (bands) {
                            double Af[9];
                            eqo_hsld_facet_a(Af, dt, r, xy, facet);
                            for (int i = 0; i < 3; ++i)
                                for (int j = 0; j < 3; ++j) {
                                    int k = band_of(vs[j] - vs[i], nW);
                                    bands[(size_t)k * N + vs[i]] += Af[3 * i + j];
                                }
                        }
                        if (b) {
                            double bf[3];
                            eqo_hsld_facet_L(bf, dt, r, s, xy, facet);
                            for (int i = 0; i < 3; ++i) b[vs[i]] += bf[i];
                        }
                    }
                }
            }
        }
    }
    return 0;
}

/* DirichletBC::apply (DOLFIN [ext]): identity row, rhs = g; columns kept. */
EQO_API void eqo_apply_dirichlet_rows(long N, double *bands, double *b,
                                      const uint8_t *mask, const double *g)
{
    for (long i = 0; i < N; ++i)
        if (mask[i]) {
            for (int k = 0; k < NBAND; ++k) bands[(size_t)k * N + i] = 0.0;
            bands[(size_t)B_C * N + i] = 1.0;
            b[i] = g[i];
        }
}

/* Symmetric elimination of the same constraints (what a CG solver needs):
 * b_f -= A_fd g_d, zero row+column, unit diagonal, b_d = g_d. */
EQO_API void eqo_apply_dirichlet_sym(long nW, long nH, double *bands, double *b,
                                     const uint8_t *mask, const double *g)
{
    const long N = nW * nH;
    long off[NBAND];
    eqo_band_offsets(nW, off);
    for (long i = 0; i < N; ++i) {
        if (mask[i]) continue;
        for (int k = 1; k < NBAND; ++k) {
            long j = i + off[k];
            double a = bands[(size_t)k * N + i];
            if (a != 0.0 && j >= 0 && j < N && mask[j]) {
                b[i] -= a * g[j];
                bands[(size_t)k * N + i] = 0.0;
            }
        }
    }
    eqo_apply_dirichlet_rows(N, bands, b, mask, g);
}

/* Dirichlet node set and values per wall.  src/fHSL.cpp:468-539 pushes the
 * DirichletBC objects in the order left, right, top, bottom; DOLFIN applies
 * them in list order, so at a corner the LAST one wins [ext].
 * wall order here: 0=left 1=right 2=top 3=bottom.  is_dir[w] != 0 marks a
 * Dirichlet wall; val[w] its constant; top_vals/bottom_vals (length nW, may
 * be NULL) give per-node values ("-1 => channel Function", :511-515,528-532). */
EQO_API void eqo_dirichlet_mask(long nW, long nH, const int *is_dir,
                                const double *val, const double *top_vals,
                                const double *bottom_vals, uint8_t *mask,
                                double *g)
{
    const long N = nW * nH;
    memset(mask, 0, N);
    for (long i = 0; i < N; ++i) g[i] = 0.0;
    if (is_dir[0])
        for (long iy = 0; iy < nH; ++iy) { mask[iy * nW] = 1; g[iy * nW] = val[0]; }
    if (is_dir[1])
        for (long iy = 0; iy < nH; ++iy) {
            mask[iy * nW + nW - 1] = 1; g[iy * nW + nW - 1] = val[1];
        }
    if (is_dir[2])
        for (long jx = 0; jx < nW; ++jx) {
            long v = (nH - 1) * nW + jx;
            mask[v] = 1; g[v] = top_vals ? top_vals[jx] : val[2];
        }
    if (is_dir[3])
        for (long jx = 0; jx < nW; ++jx) {
            mask[jx] = 1; g[jx] = bottom_vals ? bottom_vals[jx] : val[3];
        }
}

EQO_API void eqo_band_matvec(long nW, long nH, const double *bands,
                             const double *x, double *y)
{
    const long N = nW * nH;
    long off[NBAND];
    eqo_band_offsets(nW, off);
#pragma omp parallel for schedule(static)
    for (long i = 0; i < N; ++i) {
        double s = 0.0;
        for (int k = 0; k < NBAND; ++k) {
            long j = i + off[k];
            double a = bands[(size_t)k * N + i];
            if (a != 0.0 && j >= 0 && j < N) s += a * x[j];
        }
        y[i] = s;
    }
}

/* Plain (Jacobi-preconditioned) CG on the symmetric banded system; used for
 * meshes where a sparse direct factorisation does not fit, and as the
 * "PETSc-equivalent" CPU baseline (diffuclass.cpp:390-392 uses an
 * un-preconditioned Krylov solver on a matrix-free stencil).
 * Returns the iteration count (negative if not converged). */
EQO_API long eqo_cg(long nW, long nH, const double *bands, const double *b,
                    double *x, double rtol, long maxit, double *relres_out)
{
    const long N = nW * nH;
    double *r = malloc(sizeof(double) * N), *p = malloc(sizeof(double) * N),
           *q = malloc(sizeof(double) * N), *z = malloc(sizeof(double) * N);
    const double *dg = bands + (size_t)B_C * N;
    eqo_band_matvec(nW, nH, bands, x, q);
    double bb = 0, rz = 0, rr = 0;
#pragma omp parallel for reduction(+ : bb, rz, rr) schedule(static)
    for (long i = 0; i < N; ++i) {
        r[i] = b[i] - q[i];
        z[i] = r[i] / dg[i];
        p[i] = z[i];
        bb += b[i] * b[i];
        rz += r[i] * z[i];
        rr += r[i] * r[i];
    }
    if (bb == 0.0) bb = 1.0;
    long it = 0;
    const double stop = rtol * rtol * bb;
    while (rr > stop && it < maxit) {
        eqo_band_matvec(nW, nH, bands, p, q);
        double pq = 0;
#pragma omp parallel for reduction(+ : pq) schedule(static)
        for (long i = 0; i < N; ++i) pq += p[i] * q[i];
        const double alpha = rz / pq;
        double rz2 = 0;
        rr = 0;
#pragma omp parallel for reduction(+ : rz2, rr) schedule(static)
        for (long i = 0; i < N; ++i) {
            x[i] += alpha * p[i];
            r[i] -= alpha * q[i];
            z[i] = r[i] / dg[i];
            rz2 += r[i] * z[i];
            rr += r[i] * r[i];
        }
        const double beta = rz2 / rz;
        rz = rz2;
#pragma omp parallel for schedule(static)
        for (long i = 0; i < N; ++i) p[i] = z[i] + beta * p[i];
        ++it;
    }
    if (relres_out) *relres_out = sqrt(rr / bb);
    free(r); free(p); free(q); free(z);
    return rr <= stop ? it : -it;
}

EQO_API int eqo_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

EQO_API void eqo_set_num_threads(int n)
{
#ifdef _OPENMP
    omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* ------------------------------------------------------------------------- */
/* diffusionPETSc (diffuclass.cpp): finite-difference solver, matrix-free       */
/* ------------------------------------------------------------------------- */
/* Wall arrays are ordered left, right, top, bottom; a wall obeys
 * Dc*u + Nc*du/dn = BV and Nc == 0 marks a Dirichlet wall.  Natural order
 * g = i + j*nX (allXCoordinates/allYCoordinates, diffuclass.cpp:96-99).
 *
 * y = A x as MyMatMult computes it (diffuclass.cpp:786-862, the loop that is
 * compiled; the commented-out twin above it is ignored).  Ghost nodes of a
 * Neumann/Robin wall are eliminated, which doubles the inner neighbour and
 * adds 2h(Dc/Nc)F to the diagonal; Dirichlet rows are the identity with the
 * columns kept.  Deviation, as in oracle.py's fd_assemble: a corner between a
 * Neumann/Robin top/bottom wall and a Dirichlet side wall is an identity row
 * (upstream indexes xarray[j][i-1] outside the grid there). */
EQO_API void eqo_fd_matmult(long nX, long nY, double h, double F,
                            const double *Dc, const double *Nc,
                            const double *x, double *y)
{
    const double gl = Nc[0] != 0.0 ? 2.0 * h * Dc[0] / Nc[0] : 0.0;
    const double gr = Nc[1] != 0.0 ? 2.0 * h * Dc[1] / Nc[1] : 0.0;
    const double gt = Nc[2] != 0.0 ? 2.0 * h * Dc[2] / Nc[2] : 0.0;
    const double gb = Nc[3] != 0.0 ? 2.0 * h * Dc[3] / Nc[3] : 0.0;
#pragma omp parallel for schedule(static)
    for (long j = 0; j < nY; ++j) {
        const int top = (j == nY - 1), bot = (j == 0);
        for (long i = 0; i < nX; ++i) {
            const long g = i + j * nX;
            const int lef = (i == 0), rig = (i == nX - 1);
            if (!(top || bot || lef || rig)) {   /* :857-860 */
                y[g] = -F * x[g - nX] - F * x[g + nX] - F * x[g - 1] - F * x[g + 1] + (1.0 + 4.0 * F) * x[g];
                continue;
            }
            const int ywall = (top && Nc[2] != 0.0) ? 2 : ((bot && Nc[3] != 0.0) ? 3 : -1);
            if (ywall >= 0) {                    /* :792-846: a non-Dirichlet top or bottom row */
                const long gin = (ywall == 2) ? g - nX : g + nX;
                const double gy = (ywall == 2) ? gt : gb;
                if (lef && Nc[0] != 0.0)
                    y[g] = -2.0 * F * x[g + 1] - 2.0 * F * x[gin] + (1.0 + (4.0 + gy + gl) * F) * x[g];
                else if (rig && Nc[1] != 0.0)
                    y[g] = -2.0 * F * x[g - 1] - 2.0 * F * x[gin] + (1.0 + (4.0 + gy + gr) * F) * x[g];
                else if (lef || rig)
                    y[g] = x[g];                 /* Dirichlet side wall wins the corner (see above) */
                else
                    y[g] = -2.0 * F * x[gin] - F * x[g - 1] - F * x[g + 1] + (1.0 + (4.0 + gy) * F) * x[g];
            } else if (lef && Nc[0] != 0.0 && !top && !bot) {   /* :838-842 */
                y[g] = -F * x[g - nX] - F * x[g + nX] - 2.0 * F * x[g + 1] + (1.0 + (4.0 + gl) * F) * x[g];
            } else if (rig && Nc[1] != 0.0 && !top && !bot) {   /* :844-848 */
                y[g] = -F * x[g - nX] - F * x[g + nX] - 2.0 * F * x[g - 1] + (1.0 + (4.0 + gr) * F) * x[g];
            } else {
                y[g] = x[g];                     /* :850-852 */
            }
        }
    }
}

/* ApplyBoundaryConditions on b (= u0 on entry), diffuclass.cpp:191-275.
 * tBV / bBV are per-node (topBoundaryValue[i], :218,236); left/right scalars. */
EQO_API void eqo_fd_apply_bc(long nX, long nY, double h, double F,
                             const double *Nc, const double *tBV,
                             const double *bBV, double lBV, double rBV,
                             double *b)
{
    const double twoFh = 2 * F * h;
    const double lN = Nc[0], rN = Nc[1], tN = Nc[2], bN = Nc[3];
    const long top = nY - 1;
    for (long j = 0; j < nY; ++j)
        for (long i = 0; i < nX; ++i) {
            if (!(i == 0 || i == nX - 1 || j == 0 || j == top)) continue;
            double *v = b + i + j * nX;
            if (j == top) {
                if (tN != 0) {
                    if ((i != 0 || lN != 0) && (i != nX - 1 || rN != 0)) *v += (twoFh * tBV[i]) / tN;
                } else *v = tBV[i];
            } else if (j == 0) {
                if (bN != 0) {
                    if ((i != 0 || lN != 0) && (i != nX - 1 || rN != 0)) *v += (twoFh * bBV[i]) / bN;
                } else *v = bBV[i];
            }
            if (i == nX - 1) {
                if (rN != 0) {
                    if ((j != 0 || bN != 0) && (j != top || tN != 0)) *v += (twoFh * rBV) / rN;
                } else *v = rBV;
            } else if (i == 0) {
                if (lN != 0) {
                    if ((j != 0 || bN != 0) && (j != top || tN != 0)) *v += (twoFh * lBV) / lN;
                } else *v = lBV;
            }
        }
}

/* TimeStep's KSPSolve (diffuclass.cpp:405-413) with the solver InitializeDiffusion
 * configures (:386-392): no preconditioner, KSPFBCGSR -- flexible BiCGStab, which
 * without a preconditioner is plain BiCGStab [ext: PETSc] -- zero initial guess,
 * stop at ||r|| <= rtol*||b|| (PETSc's default rtol is 1e-5).  x must not alias b.
 * Returns the iteration count (negative if maxit was hit). */
EQO_API long eqo_fd_bicgstab(long nX, long nY, double h, double F,
                             const double *Dc, const double *Nc,
                             const double *b, double *x, double rtol,
                             long maxit, double *relres_out)
{
    const long N = nX * nY;
    double *r = malloc(sizeof(double) * N), *rh = malloc(sizeof(double) * N),
           *p = malloc(sizeof(double) * N), *v = malloc(sizeof(double) * N),
           *sv = malloc(sizeof(double) * N), *t = malloc(sizeof(double) * N);
    double bb = 0.0;
#pragma omp parallel for reduction(+ : bb) schedule(static)
    for (long g = 0; g < N; ++g) {
        x[g] = 0.0; r[g] = b[g]; rh[g] = b[g]; p[g] = 0.0; v[g] = 0.0;
        bb += b[g] * b[g];
    }
    if (bb == 0.0) bb = 1.0;
    const double stop = rtol * rtol * bb;
    double rho = 1.0, alpha = 1.0, omega = 1.0, rr = bb;
    long it = 0;
    while (rr > stop && it < maxit) {
        double rho1 = 0.0;
#pragma omp parallel for reduction(+ : rho1) schedule(static)
        for (long g = 0; g < N; ++g) rho1 += rh[g] * r[g];
        const double beta = (rho1 / rho) * (alpha / omega);
#pragma omp parallel for schedule(static)
        for (long g = 0; g < N; ++g) p[g] = r[g] + beta * (p[g] - omega * v[g]);
        eqo_fd_matmult(nX, nY, h, F, Dc, Nc, p, v);
        double rv = 0.0;
#pragma omp parallel for reduction(+ : rv) schedule(static)
        for (long g = 0; g < N; ++g) rv += rh[g] * v[g];
        alpha = rho1 / rv;
#pragma omp parallel for schedule(static)
        for (long g = 0; g < N; ++g) sv[g] = r[g] - alpha * v[g];
        eqo_fd_matmult(nX, nY, h, F, Dc, Nc, sv, t);
        double ts = 0.0, tt = 0.0;
#pragma omp parallel for reduction(+ : ts, tt) schedule(static)
        for (long g = 0; g < N; ++g) { ts += t[g] * sv[g]; tt += t[g] * t[g]; }
        omega = tt > 0.0 ? ts / tt : 0.0;
        rr = 0.0;
#pragma omp parallel for reduction(+ : rr) schedule(static)
        for (long g = 0; g < N; ++g) {
            x[g] += alpha * p[g] + omega * sv[g];
            r[g] = sv[g] - omega * t[g];
            rr += r[g] * r[g];
        }
        rho = rho1;
        ++it;
        if (omega == 0.0) break;
    }
    if (relres_out) *relres_out = sqrt(rr / bb);
    free(r); free(rh); free(p); free(v); free(sv); free(t);
    return rr <= stop ? it : -it;
}

/* ------------------------------------------------------------------------- */
/* Boundary flux (src/fHSL.cpp:54-96 and :156-160 + fenics/boundary.ufl)      */
/* ------------------------------------------------------------------------- */

/* src/fHSL.cpp:54-96 computeBoundaryFlux: one-sided FD flux into the channel
 * nodes below/above each column, scaled to the channel volume element. */
EQO_API void eqo_compute_boundary_flux(long nW, long nH, const double *u,
                                       double h, double dt, double D,
                                       double wellScaling, double *flux_bottom,
                                       double *flux_top)
{
    const double ds = 1.0 * h;
    for (long j = 0; j < nW; ++j) {
        double gradc = (u[1 * nW + j] - u[0 * nW + j]) / h;
        flux_bottom[j] = ds * (dt * D * gradc) / wellScaling;
        gradc = (u[(nH - 2) * nW + j] - u[(nH - 1) * nW + j]) / h;
        flux_top[j] = ds * (dt * D * gradc) / wellScaling;
    }
}

/* src/fHSL.cpp:156-157: assemble -oint grad(u).n ds over every exterior facet
 * (fenics/boundary.ufl:9-12, kernel fenics/boundary.h:2652-2741).  DOLFIN
 * visits cells in index order; each boundary cell adds its facets. */
EQO_API double eqo_boundary_functional(long nW, long nH, double W, double H,
                                       const double *u)
{
    double total = 0.0;
    for (long cy = 0; cy < nH - 1; ++cy)
        for (long cx = 0; cx < nW - 1; ++cx) {
            const long v0 = cy * nW + cx, v1 = v0 + 1, v2 = v0 + nW, v3 = v2 + 1;
            for (int t = 0; t < 2; ++t) {
                long vs[3] = {v0, t == 0 ? v1 : v2, v3};
                double xy[6], uu[3];
                for (int k = 0; k < 3; ++k) {
                    vertex_xy(vs[k], nW, nH, W, H, &xy[2 * k], &xy[2 * k + 1]);
                    uu[k] = u[vs[k]];
                }
                if (t == 0) {
                    if (cy == 0) total += eqo_boundary_facet(uu, xy, 2);       /* bottom: v0-v1 */
                    if (cx == nW - 2) total += eqo_boundary_facet(uu, xy, 0);  /* right: v1-v3 */
                } else {
                    if (cx == 0) total += eqo_boundary_facet(uu, xy, 2);       /* left: v0-v2 */
                    if (cy == nH - 2) total += eqo_boundary_facet(uu, xy, 0);  /* top: v2-v3 */
                }
            }
        }
    return total;
}

/* ------------------------------------------------------------------------- */
/* 1-D flow channels (src/fHSL.h:240-332, src/fHSL.cpp:117-143)               */
/* ------------------------------------------------------------------------- */

/* Assemble the tridiagonal CN system of fenics/AdvectionDiffusion.ufl:62-64 on
 * IntervalMesh(n-1 cells, 0, W) (src/fHSL.cpp:271-274): lo/di/up (length n)
 * and rhs from u0.  Left end = ds(1) with (r1,s1), right end = ds(2). */
EQO_API void eqo_channel_assemble(long n, double W, double dt, double D,
                                  double v, double r1, double s1, double r2,
                                  double s2, const double *u0, double *lo,
                                  double *di, double *up, double *rhs)
{
    for (long i = 0; i < n; ++i) {
        if (lo) { lo[i] = 0; di[i] = 0; up[i] = 0; }
        if (rhs) rhs[i] = 0;
    }
    for (long c = 0; c < n - 1; ++c) {
        double xc[2] = {0.0 + (double)c * (W - 0.0) / (double)(n - 1),
                        0.0 + (double)(c + 1) * (W - 0.0) / (double)(n - 1)};
        double uu[2] = {u0 ? u0[c] : 0.0, u0 ? u0[c + 1] : 0.0};
        if (lo) {
            double A[4];
            eqo_ad_cell_a(A, dt, D, v, xc);
            di[c] += A[0]; up[c] += A[1]; lo[c + 1] += A[2]; di[c + 1] += A[3];
        }
        if (rhs) {
            double be[2];
            eqo_ad_cell_L(be, uu, dt, D, v, xc);
            rhs[c] += be[0]; rhs[c + 1] += be[1];
        }
        if (c == 0) {
            if (lo) { double A[4]; eqo_ad_facet_a(A, dt, r1, 0); di[0] += A[0]; }
            if (rhs) { double be[2]; eqo_ad_facet_L(be, uu, dt, r1, s1, 0); rhs[0] += be[0]; }
        }
        if (c == n - 2) {
            if (lo) { double A[4]; eqo_ad_facet_a(A, dt, r2, 1); di[n - 1] += A[3]; }
            if (rhs) { double be[2]; eqo_ad_facet_L(be, uu, dt, r2, s2, 1); rhs[n - 1] += be[1]; }
        }
    }
}

/* Direct solve of the tridiagonal system (stands in for the sparse LU behind
 * LVS->solve(), src/fHSL.cpp:138,141 [ext]).  lo/di/up are not modified. */
EQO_API void eqo_tridiag_solve(long n, const double *lo, const double *di,
                               const double *up, const double *rhs, double *x)
{
    double *c = malloc(sizeof(double) * n), *d = malloc(sizeof(double) * n);
    c[0] = up[0] / di[0];
    d[0] = rhs[0] / di[0];
    for (long i = 1; i < n; ++i) {
        double m = di[i] - lo[i] * c[i - 1];
        c[i] = up[i] / m;
        d[i] = (rhs[i] - lo[i] * d[i - 1]) / m;
    }
    x[n - 1] = d[n - 1];
    for (long i = n - 2; i >= 0; --i) x[i] = d[i] - c[i] * x[i + 1];
    free(c); free(d);
}

/* src/fHSL.cpp:117-143: numIterations CN sub-steps of dt/numIterations, the
 * per-step flux spread evenly over the sub-steps, for one channel. */
EQO_API void eqo_channel_substeps(long n, double W, double dt, long num_iter,
                                  double D, double v, double r1, double s1,
                                  double r2, double s2, const double *flux,
                                  double *u)
{
    double *lo = malloc(sizeof(double) * n), *di = malloc(sizeof(double) * n),
           *up = malloc(sizeof(double) * n), *rhs = malloc(sizeof(double) * n);
    const double dtx = dt / (double)num_iter;
    eqo_channel_assemble(n, W, dtx, D, v, r1, s1, r2, s2, NULL, lo, di, up, NULL);
    for (long it = 0; it < num_iter; ++it) {
        for (long j = 0; j < n; ++j) u[j] += flux[j] / (double)num_iter;
        eqo_channel_assemble(n, W, dtx, D, v, r1, s1, r2, s2, u, NULL, NULL, NULL, rhs);
        eqo_tridiag_solve(n, lo, di, up, rhs, u);
    }
    free(lo); free(di); free(up); free(rhs);
}

/* src/fHSL.cpp:331-364 setRobinBoundaryConditions. */
EQO_API void eqo_robin_rates(double v, double D, double L_left, double L_right,
                             double *r_left, double *r_right)
{
    double lvdl = (L_left * v) / D;
    double lvdr = (L_right * v) / D;
    if (v > 1.0e-6) {
        *r_left = v * (1.0 / (1.0 - exp(-lvdl)));
        *r_right = v * (1.0 / (exp(lvdr) - 1.0));
    } else {
        *r_left = D / L_left;
        *r_right = D / L_right;
    }
}

/* ------------------------------------------------------------------------- */
/* Cells: rasterise, gather, scatter (src/abm/eQabm.cpp:234-449)              */
/* ------------------------------------------------------------------------- */

/* Cell record: 16 doubles, the state eQabm::updateCells reads per cell.
 *  0,1  bodyA position (cpBodyGetPosition)      2,3  bodyA rot = (cos,sin)
 *  4    offset  (vertsA[0].x = -offset)         5    newOffset (vertsA[1].x)
 *  6    radius                                  7,8  polePositionA (x,y)
 *  9,10 polePositionB                           11,12 centre (x,y)
 *  13   length                                  14,15 cos, sin of cpmCell->angle
 *       (the mean body angle setDiffusionTensor receives, src/abm/Ecoli.h:42,
 *        src/abm/cpmEcoli.cpp:387-389; evaluated by the HOST's libm, as upstream) */
#define CELL_STRIDE 16

/* src/eQ.h:119-130 ij_from_xy: size_t(round(x*n)); C round = half away from 0. */
static inline void ij_from_xy(double x, double y, double n, size_t *i, size_t *j)
{
    *j = (size_t)round(x * n);
    *i = (size_t)round(y * n);
}

/* src/abm/cpmEcoli.cpp:313-327 pointIsInCell, with Chipmunk 7.0.1
 * cpBodyWorldToLocal / cpTransformRigidInverse / cpTransformPoint / cpvcross
 * (cpBody.c, cpTransform.h, cpVect.h [ext]; centre of gravity = 0). */
EQO_API int eqo_point_in_cell(const double *c, double px, double py)
{
    const double posx = c[0], posy = c[1], rx = c[2], ry = c[3];
    const double off = c[4], noff = c[5], rad = c[6];
    /* body->transform = NewTranspose(rot.x,-rot.y,p.x, rot.y,rot.x,p.y) */
    const double ta = rx, tb = ry, tc = -ry, td = rx, ttx = posx, tty = posy;
    /* cpTransformRigidInverse */
    const double ia = td, ic = -tc, itx = (tc * tty - ttx * td);
    const double ib = -tb, id = ta, ity = (ttx * tb - ta * tty);
    /* cpTransformPoint */
    const double lx = ia * px + ic * py + itx;
    const double ly = ib * px + id * py + ity;
    /* vertsA: UL(-off,rad) UR(noff,rad) LR(noff,-rad) LL(-off,-rad)
     * (cpmEcoli.cpp:127-130,177-180,413-415); edges :71-76 */
    const double v0x = -off, v0y = rad, v1x = noff, v1y = rad;
    const double v2x = noff, v2y = -rad, v3x = -off, v3y = -rad;
    const double e0x = v1x - v0x, e0y = v1y - v0y;
    const double e1x = v2x - v1x, e1y = v2y - v1y;
    const double e2x = v3x - v2x, e2y = v3y - v2y;
    const double e3x = v0x - v3x, e3y = v0y - v3y;
    const double p0x = lx - v1x, p0y = ly - v1y;
    const double p1x = lx - v2x, p1y = ly - v2y;
    const double p2x = lx - v3x, p2y = ly - v3y;
    const double p3x = lx - v0x, p3y = ly - v0y;
    return ((e0x * p0y - e0y * p0x) < 0.0) && ((e1x * p1y - e1y * p1x) < 0.0) &&
           ((e2x * p2y - e2y * p2x) < 0.0) && ((e3x * p3y - e3y * p3x) < 0.0);
}

/* src/abm/eQabm.cpp:268-305 findInteriorPoints.  Writes node indices
 * g = i*nW + j in the reference's push_back order (row-major over the bbox).
 * Returns the number of points (>= 1).  cap bounds the write. */
EQO_API long eqo_raster_cell(const double *c, double npm, long nH, long nW,
                             long nodesToEdge, long *nodes, long cap)
{
    size_t ai, aj, bi, bj;
    ij_from_xy(c[7], c[8], npm, &ai, &aj);
    ij_from_xy(c[9], c[10], npm, &bi, &bj);
    size_t i1, i2, j1, j2;
    if (ai > bi) { i1 = bi; i2 = ai; } else { i1 = ai; i2 = bi; }
    if (aj > bj) { j1 = bj; j2 = aj; } else { j1 = aj; j2 = bj; }
    const size_t nte = (size_t)nodesToEdge;
    i1 = (i1 >= nte) ? (i1 - nte) : 0;
    j1 = (j1 >= nte) ? (j1 - nte) : 0;
    i2 = ((i2 + nte) >= (size_t)(nH - 1)) ? (size_t)(nH - 1) : (i2 + nte);
    j2 = ((j2 + nte) >= (size_t)(nW - 1)) ? (size_t)(nW - 1) : (j2 + nte);
    long cnt = 0;
    for (size_t pi = i1; pi <= i2; pi++)
        for (size_t pj = j1; pj <= j2; pj++) {
            /* src/eQ.h:115-118 xy_from_ij */
            double x = (double)pj / npm, y = (double)pi / npm;
            if (eqo_point_in_cell(c, x, y)) {
                if (cnt < cap) nodes[cnt] = (long)(pi * (size_t)nW + pj);
                cnt++;
            }
        }
    if (cnt == 0) {
        size_t ci, cj;
        ij_from_xy(c[11], c[12], npm, &ci, &cj);
        if (cap > 0) nodes[0] = (long)(ci * (size_t)nW + cj);
        cnt = 1;
    }
    return cnt;
}

EQO_API void eqo_raster(const double *cells, long ncells, double npm, long nH,
                        long nW, long nodesToEdge, long *counts, long *nodes,
                        long cap)
{
    for (long k = 0; k < ncells; ++k)
        counts[k] = eqo_raster_cell(cells + k * CELL_STRIDE, npm, nH, nW,
                                    nodesToEdge, nodes + k * cap, cap);
}

/* src/eQcell.h:40-93 volume helpers. */
static double cell_volume(double L)
{
    const double poleRadius = 1.0 / 2.0;
    const double cyl = M_PI * (poleRadius * poleRadius);
    const double poleVolume = 4.0 / 3.0 * M_PI * (poleRadius * poleRadius * poleRadius);
    return (L - 1.0) * cyl + poleVolume;
}

/* src/abm/eQabm.cpp:338-359 writeHSL amplitude: nM -> per-grid-point update. */
EQO_API double eqo_deposit_per_point(double hsl_nM, double L, double npm,
                                     long npoints)
{
    const double nanoMolarPerMoleculePerCubicMicron = 1.0 / 0.602;
    double numberHSL = hsl_nM / nanoMolarPerMoleculePerCubicMicron * cell_volume(L);
    double extra = 1.0 - cell_volume(L) / (L * 1.0 * 1.0);
    double updatePerSquareMicron = numberHSL / extra;
    double updateForOneGridPoint = updatePerSquareMicron * npm * npm;
    return updateForOneGridPoint / (double)npoints;
}

/* src/abm/eQabm.cpp:326-337 readHSL for every cell (no deposits in between). */
EQO_API void eqo_gather(const double *cells, long ncells, double npm, long nH,
                        long nW, long nodesToEdge, const double *u, double *out)
{
    long cap = 4096;
    long *nodes = malloc(sizeof(long) * cap);
    for (long k = 0; k < ncells; ++k) {
        long n = eqo_raster_cell(cells + k * CELL_STRIDE, npm, nH, nW,
                                 nodesToEdge, nodes, cap);
        double HSL = 0.0;
        for (long p = 0; p < n; ++p) HSL += u[nodes[p]];
        out[k] = HSL / (double)n;
    }
    free(nodes);
}

/* src/abm/eQabm.cpp:338-359 writeHSL for every cell, list order. */
EQO_API void eqo_scatter(const double *cells, long ncells, double npm, long nH,
                         long nW, long nodesToEdge, const double *amount_nM,
                         double *u)
{
    long cap = 4096;
    long *nodes = malloc(sizeof(long) * cap);
    for (long k = 0; k < ncells; ++k) {
        const double *c = cells + k * CELL_STRIDE;
        long n = eqo_raster_cell(c, npm, nH, nW, nodesToEdge, nodes, cap);
        double dHSL = eqo_deposit_per_point(amount_nM[k], c[13], npm, n);
        for (long p = 0; p < n; ++p) u[nodes[p]] += dHSL;
    }
    free(nodes);
}

/* Reference-order coupling (src/abm/eQabm.cpp:254-425): per cell, read then
 * write, so a later cell's read sees earlier cells' deposits.  The per-cell
 * deposit is amount = a0[k] + a1 * (gathered value)  (stand-in for
 * Strain::computeProteins, which is out of scope). */
EQO_API void eqo_update_cells_sequential(const double *cells, long ncells,
                                         double npm, long nH, long nW,
                                         long nodesToEdge, const double *a0,
                                         double a1, double *u, double *gathered)
{
    long cap = 4096;
    long *nodes = malloc(sizeof(long) * cap);
    for (long k = 0; k < ncells; ++k) {
        const double *c = cells + k * CELL_STRIDE;
        long n = eqo_raster_cell(c, npm, nH, nW, nodesToEdge, nodes, cap);
        double HSL = 0.0;
        for (long p = 0; p < n; ++p) HSL += u[nodes[p]];
        HSL = HSL / (double)n;
        gathered[k] = HSL;
        double amount = a0[k] + a1 * HSL;
        double dHSL = eqo_deposit_per_point(amount, c[13], npm, n);
        for (long p = 0; p < n; ++p) u[nodes[p]] += dHSL;
    }
    free(nodes);
}

/* src/abm/Ecoli.cpp:36-63 updatePoleCenters + the fresh (un-ratcheted) body
 * geometry of src/abm/cpmEcoli.cpp:100-130: fills a cell record from
 * (centre, angle, length, width).  rot = (cos a, sin a) as cpBodySetAngle
 * (cpvforangle [ext]). */
EQO_API void eqo_make_cell(double cx, double cy, double angle, double length,
                           double width, double trapW, double trapH,
                           double *rec)
{
    double ca = cos(angle), sa = sin(angle);
    rec[0] = cx; rec[1] = cy; rec[2] = ca; rec[3] = sa;
    rec[4] = (length - width) * 0.5;
    rec[5] = rec[4];
    rec[6] = width * 0.5;
    for (int r = 1, k = 0; k < 2; ++k, r = -1) {
        double s = (r)*0.5 * length - width / 2.0;
        double px = cx + ca * s, py = cy + sa * s;
        if (px < 0.0) px = 0.0;
        if (py < 0.0) py = 0.0;
        if (px > trapW) px = trapW;
        if (py > trapH) py = trapH;
        rec[7 + 2 * k] = px; rec[8 + 2 * k] = py;
    }
    rec[11] = cx; rec[12] = cy; rec[13] = length;
    rec[14] = ca; rec[15] = sa;  /* fresh cell: angle == both body angles (cpmEcoli.cpp:107,155-158) */
}

/* src/abm/eQabm.cpp:246-248 (D11/D22/D12 grids reset to 1,1,0 every step) and
 * :306-325,407 setDiffusionTensor for every cell in list order: on the cell's
 * interior points D11 = Dx c^2 + Dy s^2, D22 = Dx s^2 + Dy c^2,
 * D12 = (Dx - Dy) s c with c = cos(theta), s = sin(theta) (record slots 14,15);
 * a later cell overwrites an earlier one on a shared node; points outside the
 * grid are skipped (gridFunction::isValidIndex, src/eQ.h:66-69). */
EQO_API void eqo_cells_tensor(const double *cells, long ncells, double npm,
                              long nH, long nW, long nodesToEdge, double Dx,
                              double Dy, double *d11, double *d22, double *d12)
{
    const long N = nH * nW;
    for (long g = 0; g < N; ++g) { d11[g] = 1.0; d22[g] = 1.0; d12[g] = 0.0; }
    long cap = 4096;
    long *nodes = malloc(sizeof(long) * cap);
    for (long k = 0; k < ncells; ++k) {
        const double *c = cells + k * CELL_STRIDE;
        long n = eqo_raster_cell(c, npm, nH, nW, nodesToEdge, nodes, cap);
        const double ct = c[14], st = c[15];
        double cos2t = ct * ct;
        double sin2t = st * st;
        double sincost = st * ct;
        for (long p = 0; p < n; ++p) {
            const long i = nodes[p] / nW, j = nodes[p] - i * nW;
            if (nodes[p] >= 0 && i < nH && j < nW) {
                d11[nodes[p]] = Dx * cos2t + Dy * sin2t;
                d22[nodes[p]] = Dx * sin2t + Dy * cos2t;
                d12[nodes[p]] = (Dx - Dy) * sincost;
            }
        }
    }
    free(nodes);
}
