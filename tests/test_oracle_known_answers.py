"""Known-answer tests of the oracle pipeline (SURVEY.md 8c list): the reference ships none, so these pin the
restated assembly / boundary conditions / solve / cells / channels against closed forms and invariants."""
import ctypes as C

import numpy as np
import pytest


def test_interior_stencil_closed_form(oracle):
    # h=0.5, tau=dt*D=120: centre h^2/2+4tau, axis h^2/12-tau, NE/SW h^2/12 (SURVEY.md 8a closed form)
    p = oracle.Problem(nW=7, nH=6, bc_type=(0, 0, 0, 0))
    bands, _ = oracle.assemble(p, None)
    c = 3 * 7 + 3
    want = [480.125, -119.97916666666667, -119.97916666666667, -119.97916666666667, -119.97916666666667,
            0.020833333333333332, 0.020833333333333332]
    assert np.allclose(bands[:, c], want, rtol=1e-13)


def test_row_sums_are_mass_row_sums(oracle):
    # K*1 = 0, so A*1 = M*1 = L(1); both assembled independently
    p = oracle.Problem(nW=23, nH=17, bc_type=(0, 0, 0, 0))
    bands, b = oracle.assemble(p, np.ones(p.N))
    assert np.allclose(oracle.band_matvec(p, bands, np.ones(p.N)), b, rtol=1e-12)
    assert np.isclose(b.sum(), p.W * p.H, rtol=1e-13)  # total area


def test_matrix_is_symmetric(oracle):
    p = oracle.Problem(nW=19, nH=11, bc_type=(2, 2, 0, 0), bc_value=(138.78, 18.78, 0, 0))
    rng = np.random.default_rng(0)
    p.d11, p.d22, p.d12 = rng.uniform(.5, 2, p.N), rng.uniform(.5, 2, p.N), rng.uniform(-.3, .3, p.N)
    bands, _ = oracle.assemble(p, None)
    A = oracle.bands_to_csr(p, bands)
    assert abs(A - A.T).max() < 1e-10


def test_neumann_cosine_mode_symbol(oracle):
    """One BE step multiplies cos(k pi x/W) cos(l pi y/H) by m/(m + tau*k) with the DISCRETE symbols of the
    P1 mass and stiffness on the 'right' mesh (the NE/SW mass coupling included)."""
    nW, nH, k, l = 33, 25, 3, 2
    p = oracle.Problem(nW=nW, nH=nH, bc_type=(0, 0, 0, 0))
    y, x = np.mgrid[0:nH, 0:nW]
    tx, ty = k * np.pi / (nW - 1), l * np.pi / (nH - 1)
    u0 = (np.cos(tx * x) * np.cos(ty * y)).ravel()
    u1 = oracle.solve_lu(p, u0)
    h2, tau = p.h ** 2, p.dt * p.D
    # cos modes are eigenvectors only of the symmetric part; the NE/SW term mixes (k,l) with sin*sin,
    # which is not in the Neumann space -- so check the Rayleigh quotient to discretisation accuracy instead
    m = h2 * (0.5 + (np.cos(tx) + np.cos(ty)) / 6 + np.cos(tx) * np.cos(ty) / 6)
    kk = 2 * (2 - np.cos(tx) - np.cos(ty))
    factor = m / (m + tau * kk)
    got = (u1 @ u0) / (u0 @ u0)
    assert abs(got - factor) < 2e-3 * factor
    # exact identity: the solve satisfies the assembled system
    bands, b = oracle.assemble(p, u0)
    assert np.linalg.norm(oracle.band_matvec(p, bands, u1) - b) < 1e-10 * np.linalg.norm(b)


def test_dirichlet_walls_hold_values_and_interior_decays(oracle):
    p = oracle.Problem(nW=41, nH=21, bc_type=(1, 1, 1, 1), bc_value=(1.0, 2.0, 3.0, 4.0))
    u = oracle.solve_lu(p, np.zeros(p.N)).reshape(p.nH, p.nW)
    assert np.allclose(u[1:-1, 0], 1.0, rtol=1e-13) and np.allclose(u[1:-1, -1], 2.0, rtol=1e-13)
    # corners: DirichletBC list order left,right,top,bottom -> the last applied wins
    assert np.allclose(u[-1, :], 3.0, rtol=1e-13) and np.allclose(u[0, :], 4.0, rtol=1e-13)
    assert u[1:-1, 1:-1].min() > 0.99 and u[1:-1, 1:-1].max() < 4.01  # discrete maximum principle


def test_symmetric_elimination_equals_identity_rows(oracle):
    p = oracle.Problem(nW=37, nH=19, bc_type=(2, 1, 1, 0), bc_value=(50.0, 2.0, 3.0, 0.0))
    u0 = np.random.default_rng(3).uniform(0, 5, p.N)
    a = oracle.solve_lu(p, u0, symmetric=False)
    b = oracle.solve_lu(p, u0, symmetric=True)
    assert np.linalg.norm(a - b) < 1e-12 * np.linalg.norm(a)
    c, it, rel = oracle.solve_cg(p, u0)
    assert it > 0 and np.linalg.norm(a - c) < 1e-10 * np.linalg.norm(a)


def test_robin_steady_profile_1d(oracle):
    """Robin walls with rate r = D/L (src/fHSL.cpp:356-357): the steady profile of a uniformly fed strip is
    the parabola with u'(0) = r u(0)/D; run to steady state and compare the wall/centre ratio."""
    nW, nH = 81, 5
    D, L = 1200.0, 20.0
    r = D / L
    p = oracle.Problem(nW=nW, nH=nH, D=D, dt=0.1, bc_type=(2, 2, 0, 0), bc_value=(r, r, 0, 0))
    src = 1.0
    u = np.zeros(p.N)
    for _ in range(400):
        u = oracle.solve_lu(p, u + p.dt * src)
    u = u.reshape(nH, nW)[2]
    W = p.W
    # D u'' = -src, D u'(0) = r u(0): u(0) = src*W/(2r), u(W/2) = u(0) + src*W^2/(8D)
    assert np.isclose(u[0], src * W / (2 * r), rtol=2e-3)
    assert np.isclose(u[nW // 2], src * W / (2 * r) + src * W * W / (8 * D), rtol=2e-3)


def test_mass_conservation_neumann(oracle):
    p = oracle.Problem(nW=65, nH=33, bc_type=(0, 0, 0, 0))
    u0 = np.random.default_rng(0).uniform(0, 1, p.N)
    u1 = oracle.solve_lu(p, u0)
    _, m0 = oracle.assemble(p, u0, want_matrix=False)
    _, m1 = oracle.assemble(p, u1, want_matrix=False)
    assert np.isclose(m0.sum(), m1.sum(), rtol=1e-12)


def test_boundary_functional_on_linear_field(oracle):
    # u = a x + b y: -oint grad(u).n ds = 0 ; u = x^2: -oint = -2 W H exactly for P1 one-sided differences? no:
    p = oracle.Problem(nW=21, nH=11, bc_type=(0, 0, 0, 0))
    y, x = np.mgrid[0:p.nH, 0:p.nW] * p.h
    assert abs(oracle.boundary_functional(p, (2 * x + 3 * y).ravel())) < 1e-10
    # u = x: left wall contributes +H (outward normal -x: -(-1)*1), right wall -H
    u = (x ** 2).ravel()
    W, H, h = p.W, p.H, p.h
    # one-sided P1 gradient of x^2 at the walls: left (h^2-0)/h = h, right (W^2-(W-h)^2)/h = 2W-h
    assert np.isclose(oracle.boundary_functional(p, u), h * H - (2 * W - h) * H, rtol=1e-12)


def test_channel_cn_conserves_mass_without_robin(oracle):
    # v = 0, r = 0: Crank-Nicolson diffusion with natural ends conserves sum(M u)
    p = oracle.Problem(nW=101, nH=5, channels=True, channel_v=0.0, channel_r=(0.0, 0.0), channel_iters=8)
    u0 = np.exp(-((np.arange(p.nW) - 50) / 6.0) ** 2)
    u1 = oracle.channel_substeps(p, np.zeros(p.nW), u0)
    w = np.full(p.nW, p.h); w[0] = w[-1] = p.h / 2
    assert np.isclose((w * u0).sum(), (w * u1).sum(), rtol=1e-12)
    assert u1.max() < u0.max()


def test_channel_advection_moves_pulse_downstream(oracle):
    p = oracle.Problem(nW=201, nH=5, D=10.0, channels=True, channel_v=120.0, channel_r=(0.0, 0.0), channel_iters=48)
    x = np.arange(p.nW) * p.h
    u0 = np.exp(-((x - 30.0) / 4.0) ** 2)
    u1 = oracle.channel_substeps(p, np.zeros(p.nW), u0)
    c0, c1 = (x * u0).sum() / u0.sum(), (x * u1).sum() / u1.sum()
    assert np.isclose(c1 - c0, 120.0 * p.dt, rtol=0.05)  # v*dt


def test_robin_rates_formula(oracle):
    rl, rr = oracle.robin_rates(120.0, 1200.0, 20.0, 20.0)
    assert np.isclose(rl, 120 / (1 - np.exp(-2.0))) and np.isclose(rr, 120 / (np.exp(2.0) - 1))
    rl, rr = oracle.robin_rates(0.0, 1200.0, 20.0, 40.0)
    assert rl == 60.0 and rr == 30.0


# ---- cells -------------------------------------------------------------------------------------------------
def _inside_numpy(rec, x, y):
    """Independent restatement of the rod predicate: strict interior of bodyA's rectangle."""
    dx, dy = x - rec[0], y - rec[1]
    lx = rec[2] * dx + rec[3] * dy
    ly = -rec[3] * dx + rec[2] * dy
    return (lx > -rec[4]) & (lx < rec[5]) & (abs(ly) < rec[6])


def test_raster_matches_independent_predicate(oracle):
    npm, nW, nH = 2.0, 201, 41
    W, H = 100.0, 20.0
    rng = np.random.default_rng(4)
    n = 300
    cells = oracle.make_cells(np.c_[rng.uniform(2, W - 2, n), rng.uniform(2, H - 2, n)], rng.uniform(0, 2 * np.pi, n),
                              (1 + rng.uniform(size=n)) * 2.1, W, H)
    cnt, nodes = oracle.raster(cells, npm, nH, nW, cap=256)
    yy, xx = np.mgrid[0:nH, 0:nW] / npm
    for k in range(n):
        ins = _inside_numpy(cells[k], xx, yy)
        want = np.flatnonzero(ins.ravel())
        # away from rounding-sensitive edges the two formulations agree exactly
        got = nodes[k, :cnt[k]]
        if len(want) == 0:
            assert cnt[k] == 1
        else:
            assert set(want.tolist()) == set(got.tolist())
            assert np.all(np.diff(got) > 0)  # the reference's row-major push_back order


def test_raster_fallback_and_clamping(oracle):
    npm, nW, nH = 2.0, 201, 41
    # a rod thinner than the node pitch between nodes -> no interior node -> the centre node (eQabm.cpp:299-303)
    rec = oracle.make_cells([(10.26, 5.26)], [0.0], [1.05], 100.0, 20.0)
    rec[0, 4] = rec[0, 5] = 0.02
    rec[0, 6] = 0.02
    cnt, nodes = oracle.raster(rec, npm, nH, nW, cap=16)
    assert cnt[0] == 1 and nodes[0, 0] == int(round(5.26 * 2)) * nW + int(round(10.26 * 2))
    # poles are clamped to the trap (Ecoli.cpp:47-50,57-60): a rod poking through the wall still rasterises in-grid
    rec = oracle.make_cells([(0.3, 0.2)], [np.pi / 4], [4.0], 100.0, 20.0)
    cnt, nodes = oracle.raster(rec, npm, nH, nW, cap=64)
    assert cnt[0] >= 1 and nodes[0, :cnt[0]].min() >= 0 and nodes[0, :cnt[0]].max() < nW * nH


def test_scatter_deposits_expected_molecule_count(oracle):
    # writeHSL: every cell adds n# * npm^2 / (1 - V/L) spread over its points (eQabm.cpp:338-359)
    npm, nW, nH = 2.0, 201, 41
    rec = oracle.make_cells([(50.0, 10.0)], [0.3], [3.0], 100.0, 20.0)
    u = oracle.scatter(rec, npm, nH, nW, np.array([100.0]), np.zeros(nW * nH))
    L = 3.0
    V = (L - 1) * np.pi / 4 + np.pi / 6
    want = 100.0 * 0.602 * V / (1 - V / L) * npm * npm
    assert np.isclose(u.sum(), want, rtol=1e-13)
    cnt, _ = oracle.raster(rec, npm, nH, nW)
    assert np.count_nonzero(u) == cnt[0]


def test_gather_is_mean_over_points(oracle):
    npm, nW, nH = 2.0, 201, 41
    rec = oracle.make_cells([(50.0, 10.0), (20.0, 5.0)], [0.3, 1.2], [3.0, 4.0], 100.0, 20.0)
    u = np.random.default_rng(0).uniform(0, 9, nW * nH)
    g = oracle.gather(rec, npm, nH, nW, u)
    cnt, nodes = oracle.raster(rec, npm, nH, nW)
    for k in range(2):
        assert np.isclose(g[k], u[nodes[k, :cnt[k]]].mean(), rtol=1e-14)


def test_sequential_order_differs_only_on_overlap(oracle):
    """Quirk 5: the reference reads and writes cell by cell.  For separated rods
    gather-all-then-scatter-all (the GPU order) is identical; for overlapping rods later cells see earlier deposits."""
    npm, nW, nH = 2.0, 201, 41
    sep = oracle.synthetic_colony(60, 100.0, 20.0, seed=2)
    a0 = np.full(len(sep), 100.0)
    u0 = np.random.default_rng(1).uniform(0, 5, nW * nH)
    useq, gseq = oracle.update_cells_sequential(sep, npm, nH, nW, a0, 0.5, u0)
    g = oracle.gather(sep, npm, nH, nW, u0)
    ubat = oracle.scatter(sep, npm, nH, nW, a0 + 0.5 * g, u0)
    assert np.array_equal(gseq, g) and np.array_equal(useq, ubat)
    over = oracle.make_cells([(50.0, 10.0), (50.4, 10.1)], [0.0, 0.0], [3.0, 3.0], 100.0, 20.0)
    _, gseq = oracle.update_cells_sequential(over, npm, nH, nW, np.full(2, 100.0), 0.5, u0)
    g = oracle.gather(over, npm, nH, nW, u0)
    assert gseq[0] == g[0] and gseq[1] > g[1]


def test_full_step_with_channels_runs_and_feeds_back(oracle):
    rl, rr = oracle.robin_rates(120.0, 1200.0, 20.0, 20.0)
    p = oracle.Problem(nW=101, nH=21, bc_type=(2, 2, 3, 3), bc_value=(rl, rr, 0, 0), channels=True,
                       channel_r=(rl, rr), well_scaling=12.5)
    s = oracle.new_state(p)
    s.u[:] = 1.0
    for _ in range(3):
        prev_bottom, prev_top = s.bottom.copy(), s.top.copy()
        s = oracle.step(p, s)
    # HSL leaves through top/bottom into the channels, which feed the Dirichlet rows of the NEXT step
    # (one-step lag, src/fHSL.cpp:100,514,531)
    assert s.top.max() > 0 and s.bottom.max() > 0
    u = s.u.reshape(p.nH, p.nW)
    assert s.total_boundary_flux > 0  # -oint grad(u).n > 0: net outflow
    assert np.allclose(u[0], prev_bottom, rtol=1e-12) and np.allclose(u[-1], prev_top, rtol=1e-12)


def test_cells_tensor_known_answers(oracle):
    """setDiffusionTensor (src/abm/eQabm.cpp:246-248,306-325): isotropic away from rods; on a horizontal rod
    (theta = 0) D11 = Dx, D22 = Dy, D12 = 0; on a vertical one the roles swap; trace and determinant of the
    nodal tensor are rotation invariants (Dx + Dy, Dx * Dy); a later rod overwrites an earlier one."""
    nW, nH, npm = 81, 61, 2.0
    W, H = (nW - 1) / npm, (nH - 1) / npm
    Dx, Dy = 1.5, 0.6
    cells = oracle.make_cells([(10.0, 10.0), (25.0, 12.0), (30.0, 22.0)], [0.0, np.pi / 2, 0.7], [4.0, 3.5, 4.2], W, H)
    d11, d22, d12 = oracle.cells_tensor(cells, npm, nH, nW, Dx, Dy)
    cnt, nodes = oracle.raster(cells, npm, nH, nW, cap=256)
    touched = np.zeros(nW * nH, dtype=bool)
    for k in range(3):
        touched[nodes[k, :cnt[k]]] = True
    assert np.all(d11[~touched] == 1.0) and np.all(d22[~touched] == 1.0) and np.all(d12[~touched] == 0.0)
    a = nodes[0, :cnt[0]]
    assert np.all(d11[a] == Dx) and np.all(d22[a] == Dy) and np.all(d12[a] == 0.0)
    b = nodes[1, :cnt[1]]
    assert np.allclose(d11[b], Dy, rtol=1e-15) and np.allclose(d22[b], Dx, rtol=1e-15) and np.all(np.abs(d12[b]) < 1e-16)
    c = nodes[2, :cnt[2]]
    assert np.allclose(d11[c] + d22[c], Dx + Dy, rtol=1e-14)
    assert np.allclose(d11[c] * d22[c] - d12[c] ** 2, Dx * Dy, rtol=1e-14)
    # list order: the same rod twice with different angles -> the later record's tensor stays
    two = oracle.make_cells([(10.0, 10.0), (10.0, 10.0)], [0.0, np.pi / 2], [4.0, 4.0], W, H)
    e11, _, _ = oracle.cells_tensor(two, npm, nH, nW, Dx, Dy)
    cnt2, nodes2 = oracle.raster(two, npm, nH, nW, cap=256)
    both = np.intersect1d(nodes2[0, :cnt2[0]], nodes2[1, :cnt2[1]])
    assert len(both) > 0 and np.allclose(e11[both], Dy, rtol=1e-15)
    # Dx = Dy = 1 (the shipped values, src/eQinit.h:64-65): identity up to the rounding of c^2 + s^2
    i11, i22, i12 = oracle.cells_tensor(cells, npm, nH, nW, 1.0, 1.0)
    assert np.abs(i11 - 1).max() < 3e-16 and np.abs(i22 - 1).max() < 3e-16 and np.all(i12 == 0.0)


def test_fd_oracle_known_answers(oracle):
    """diffusionPETSc restatement (diffuclass.cpp:191-275,786-862): under all-Neumann walls the ghost-node
    rows make cos(k pi x/W) cos(l pi y/H) an exact eigenvector with factor 1/(1 + F(4 - 2cos tx - 2cos ty));
    rows sum to one (a constant stays constant, sum of w*u is conserved); times the node's cell share the
    matrix is symmetric; Dirichlet walls hold their values with left/right winning the corners."""
    import scipy.sparse as sp
    nW, nH, k, l = 33, 25, 3, 2
    p = oracle.Problem(nW=nW, nH=nH, bc_type=(0, 0, 0, 0))
    y, x = np.mgrid[0:nH, 0:nW]
    tx, ty = k * np.pi / (nW - 1), l * np.pi / (nH - 1)
    u0 = (np.cos(tx * x) * np.cos(ty * y)).ravel()
    F = p.D * p.dt / p.h ** 2
    u1 = oracle.fd_solve(p, u0)
    assert np.allclose(u1, u0 / (1 + F * (4 - 2 * np.cos(tx) - 2 * np.cos(ty))), rtol=0, atol=1e-13)
    w = oracle.fd_node_weights(p)
    rng = np.random.default_rng(3)
    v0 = rng.uniform(0, 5, p.N)
    assert np.isclose(w @ oracle.fd_solve(p, v0), w @ v0, rtol=1e-12)
    pr = oracle.Problem(nW=19, nH=11, bc_type=(2, 2, 0, 0), bc_value=(138.78, 18.78, 0, 0), robin_s=(0.3, 0.1))
    S = sp.diags(oracle.fd_node_weights(pr)) @ oracle.fd_assemble(pr, oracle.fd_walls_from_problem(pr))
    assert abs(S - S.T).max() < 1e-12
    pd = oracle.Problem(nW=21, nH=15, bc_type=(1, 1, 1, 1), bc_value=(1.0, 2.0, 3.0, 4.0))
    u = oracle.fd_solve(pd, v0[:pd.N]).reshape(pd.nH, pd.nW)
    eq = lambda a, v: np.allclose(a, v, rtol=1e-10, atol=0)   # identity rows with kept columns: LU rounding
    assert eq(u[:, 0], 1.0) and eq(u[:, -1], 2.0)                      # corners included: left/right win
    assert eq(u[-1, 1:-1], 3.0) and eq(u[0, 1:-1], 4.0)
    # mixed: Dirichlet top/bottom, Robin sides (the consistent mixed case upstream): Dirichlet rows at the corners
    pm = oracle.Problem(nW=21, nH=15, bc_type=(2, 2, 1, 1), bc_value=(138.78, 18.78, 2.0, 0.5), robin_s=(0.3, 0.1))
    um = oracle.fd_solve(pm, v0[:pm.N]).reshape(pm.nH, pm.nW)
    assert eq(um[-1, :], 2.0) and eq(um[0, :], 0.5)


def test_fd_robin_steady_profile(oracle):
    """1-D steady state between a Dirichlet-like source and a Robin wall: with Neumann top/bottom and
    u0 == u the FD solution is constant in y and its left-wall ghost-node relation holds to rounding."""
    p = oracle.Problem(nW=41, nH=7, bc_type=(2, 1, 0, 0), bc_value=(60.0, 5.0, 0, 0), robin_s=(0.0, 0.0), dt=1e6)
    u = oracle.fd_solve(p, np.zeros(p.N)).reshape(p.nH, p.nW)   # dt -> infinity: the steady profile
    assert np.allclose(u, u[0], rtol=1e-10)
    # steady 1-D: linear profile with D u'(0) = r u(0)  (ghost node: (u1 - u_-1)/(2h) = (r/D) u0)
    r, D, h = 60.0, p.D, p.h
    slope = (u[0, 2] - u[0, 1]) / h
    assert np.isclose(slope, (r / D) * u[0, 0], rtol=1e-5)
    assert np.isclose(u[0, -1], 5.0)


@pytest.mark.parametrize("bc", [
    dict(bc_type=(1, 1, 1, 1), bc_value=(1.0, 2.0, 3.0, 4.0)),
    dict(bc_type=(0, 0, 0, 0)),
    dict(bc_type=(2, 2, 0, 0), bc_value=(138.78, 18.78, 0, 0), robin_s=(0.3, 0.1)),
    dict(bc_type=(2, 2, 1, 1), bc_value=(138.78, 18.78, 2.0, 0.5), robin_s=(0.3, 0.1)),
    dict(bc_type=(1, 1, 0, 0), bc_value=(1.5, 0.5, 0, 0)),
    dict(bc_type=(0, 2, 1, 0), bc_value=(0, 50.0, 1.0, 0)),
])
def test_fd_two_restatements_agree(oracle, bc):
    """MyMatMult / ApplyBoundaryConditions restated twice, independently: as a sparse matrix in numpy
    (oracle.fd_assemble / fd_rhs) and as the matrix-free C loops the reference runs (eqo_fd_matmult /
    eqo_fd_apply_bc); and the reference's own solver -- unpreconditioned BiCGStab to PETSc's default
    rtol 1e-5 (diffuclass.cpp:386-392) -- lands within that tolerance of the exact solve."""
    p = oracle.Problem(nW=61, nH=37, **bc)
    w = oracle.fd_walls_from_problem(p)
    rng = np.random.default_rng(1)
    x = rng.uniform(0, 5, p.N)
    A = oracle.fd_assemble(p, w)
    assert np.allclose(A @ x, oracle.fd_matmult(p, w, x), rtol=1e-14, atol=1e-11)
    assert np.array_equal(oracle.fd_rhs(p, w, x), oracle.fd_rhs_c(p, w, x))
    exact = oracle.fd_solve(p, x)
    u5, it5, rel5 = oracle.fd_step_krylov(p, x)
    assert it5 > 0 and rel5 <= 1e-5
    assert np.linalg.norm(u5 - exact) < 1e-4 * np.linalg.norm(exact)
    u13, it13, _ = oracle.fd_step_krylov(p, x, rtol=1e-13)
    assert it13 > it5 and np.linalg.norm(u13 - exact) < 1e-10 * np.linalg.norm(exact)
