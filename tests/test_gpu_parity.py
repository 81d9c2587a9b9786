"""GPU parity tests: the CUDA path (through the C-ABI, eq_b200.GpuHSL) against the
CPU oracle on identical meshes, fields, cells and dt.

Bars (BASELINE.json north_star): fields and per-cell sampled concentrations
rel-L2 <= 1e-8 against the direct (LU) solve; cell -> node lookup bit-exact.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import eq_b200 as E

TOL = 1e-8  # relative L2, north_star


def rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def field(p, seed=0, smooth=True):
    rng = np.random.default_rng(seed)
    if not smooth:
        return rng.uniform(0, 10, p.N)
    y, x = np.mgrid[0:p.nH, 0:p.nW]
    u = 5 + 3 * np.sin(2 * np.pi * x / p.nW) * np.cos(np.pi * y / p.nH) + rng.uniform(0, 1, (p.nH, p.nW))
    return u.ravel()


def make(oracle, nW, nH, **kw):
    p = oracle.Problem(nW=nW, nH=nH, **kw)
    g = E.GpuHSL(nW, nH, h=p.h, hy=p.hy, dt=p.dt, D=p.D, bc_type=p.bc_type, bc_value=p.bc_value,
                 robin_s=p.robin_s, channels=p.channels, channel_v=p.channel_v,
                 channel_r=p.channel_r, channel_iters=p.channel_iters, well_scaling=p.well_scaling)
    return p, g


BCS = {
    "dirichlet0": dict(bc_type=(1, 1, 1, 1), bc_value=(0, 0, 0, 0)),
    "neumann": dict(bc_type=(0, 0, 0, 0), bc_value=(0, 0, 0, 0)),
    "robin_lr": dict(bc_type=(2, 2, 0, 0), bc_value=(120.0, 120.0, 0, 0)),
    "robin_lr_dir_tb": dict(bc_type=(2, 2, 1, 1), bc_value=(138.78, 18.78, 2.0, 0.5), robin_s=(0.3, 0.1)),
    "threewall": dict(bc_type=(0, 0, 0, 1), bc_value=(0, 0, 0, 0)),
    "dir_values": dict(bc_type=(1, 1, 1, 1), bc_value=(1.0, 2.0, 3.0, 4.0)),
}


@pytest.mark.parametrize("bc", list(BCS))
@pytest.mark.parametrize("shape", [(201, 41), (64, 37), (17, 16)])
def test_operator_matches_assembled_matrix(oracle, bc, shape):
    """(1) matrix-free stencil == the Fenics-style assembled matrix (unconstrained and constrained)."""
    p, g = make(oracle, *shape, **BCS[bc])
    bands, _ = oracle.assemble(p, None)
    x = field(p, 1, smooth=False)
    y_ref = oracle.band_matvec(p, bands, x)
    y = g.apply_operator(x)
    assert rel(y, y_ref) < 1e-13
    # constrained operator vs symmetric elimination of the same matrix
    import ctypes as C
    mask, gv = oracle.dirichlet(p)
    b = np.zeros(p.N)
    oracle.lib().eqo_apply_dirichlet_sym(C.c_long(p.nW), C.c_long(p.nH), oracle._dp(bands), oracle._dp(b),
                                         mask.ctypes.data_as(oracle.c_u8p), oracle._dp(gv))
    assert rel(g.apply_operator(x, constrained=True), oracle.band_matvec(p, bands, x)) < 1e-13
    g.close()


@pytest.mark.parametrize("bc", ["dirichlet0", "robin_lr_dir_tb"])
def test_rhs_matches_assembled_load(oracle, bc):
    p, g = make(oracle, 201, 41, **BCS[bc])
    u0 = field(p, 2)
    _, b_ref = oracle.assemble(p, u0, want_matrix=False)
    assert rel(g.build_rhs(u0), b_ref) < 1e-13
    g.close()


@pytest.mark.parametrize("bc", list(BCS))
@pytest.mark.parametrize("shape", [(201, 41), (257, 257), (130, 75)])
def test_step_matches_direct_solve(oracle, bc, shape):
    """(2) one backward-Euler step == assemble + DirichletBC + sparse LU (src/fHSL.cpp:104-108)."""
    p, g = make(oracle, *shape, **BCS[bc])
    u0 = field(p, 3)
    ref = oracle.solve_lu(p, u0)
    g.solution_vector[:] = u0
    out = g.stepDiffusion()
    st = g.stats()
    assert rel(out, ref) < TOL, (st.iterations, st.relres)
    # flux functional (src/fHSL.cpp:156-160)
    want = p.D * p.dt * oracle.boundary_functional(p, ref)
    assert abs(g.totalBoundaryFlux - want) <= 1e-7 * max(abs(want), 1.0)
    g.close()


def test_multi_step_default_trap(oracle):
    """Config 1: default trap 201x41, Dirichlet-0, 32 cells depositing, 20 steps."""
    p, g = make(oracle, 201, 41, **BCS["dirichlet0"])
    rng = np.random.default_rng(5)
    n = 32
    centers = np.c_[rng.uniform(3, p.W - 3, n), rng.uniform(3, p.H - 3, n)]
    cells = oracle.make_cells(centers, rng.uniform(0, 2 * np.pi, n), (1 + rng.uniform(size=n)) * 2.1, p.W, p.H)
    npm = 1.0 / p.h
    g.upload_cells(cells, npm)
    s = oracle.new_state(p)
    for k in range(20):
        amount = 100.0 + 5.0 * oracle.gather(cells, npm, p.nH, p.nW, s.u)
        s.u = oracle.scatter(cells, npm, p.nH, p.nW, amount, s.u)
        s = oracle.step(p, s)
        ga = 100.0 + 5.0 * g.gather()
        g.scatter(ga)
        g.step()
    assert rel(g.get_field(), s.u) < TOL
    assert rel(g.gather(), oracle.gather(cells, npm, p.nH, p.nW, s.u)) < TOL
    g.close()


def test_channels_config4_small(oracle):
    """Microfluidic trap: Robin left/right, top/bottom Dirichlet from the 1-D channels,
    CN channel sub-steps driven by the FD wall flux (src/fHSL.cpp:110-152)."""
    rl, rr = oracle.robin_rates(120.0, 1200.0, 20.0, 20.0)
    kw = dict(bc_type=(2, 2, 3, 3), bc_value=(rl, rr, 0, 0), channels=True, channel_v=120.0,
              channel_r=(rl, rr), channel_iters=48, well_scaling=10.0 * (25.0 / 5.0) * 0.5)
    p, g = make(oracle, 201, 41, **kw)
    s = oracle.new_state(p)
    rng = np.random.default_rng(7)
    for k in range(6):
        dep = np.zeros(p.N)
        dep[rng.integers(0, p.N, 50)] += rng.uniform(10, 100, 50)
        s.u = s.u + dep
        s = oracle.step(p, s)
        g.set_field(g.get_field() + dep)
        g.step()
    t, b = g.channels()
    assert rel(g.get_field(), s.u) < TOL
    assert rel(t, s.top) < TOL and rel(b, s.bottom) < TOL
    assert abs(g.totalBoundaryFlux - s.total_boundary_flux) <= 1e-7 * abs(s.total_boundary_flux)
    g.close()


def test_tensor_operator_and_step(oracle):
    """Variable anisotropic tensor D*[[D11,D12],[D12,D22]] (fenics/hslD.ufl:34)."""
    p, g = make(oracle, 97, 65, **BCS["robin_lr_dir_tb"])
    rng = np.random.default_rng(11)
    th = rng.uniform(0, np.pi, p.N)
    dx, dy = 1.5, 0.6
    p.d11 = dx * np.cos(th) ** 2 + dy * np.sin(th) ** 2
    p.d22 = dx * np.sin(th) ** 2 + dy * np.cos(th) ** 2
    p.d12 = (dx - dy) * np.sin(th) * np.cos(th)
    g.set_tensor(p.d11, p.d22, p.d12)
    bands, _ = oracle.assemble(p, None)
    x = field(p, 1, smooth=False)
    assert rel(g.apply_operator(x), oracle.band_matvec(p, bands, x)) < 1e-12
    u0 = field(p, 4)
    ref = oracle.solve_lu(p, u0)
    g.solution_vector[:] = u0
    assert rel(g.stepDiffusion(), ref) < TOL
    g.close()


@pytest.mark.parametrize("npm,shape", [(2.0, (201, 41)), (4.0, (401, 81))])
def test_cells_tensor_feed_bit_exact_and_step(oracle, npm, shape):
    """eqgpu_cells_tensor == eQabm::updateCells' setDiffusionTensor (src/abm/eQabm.cpp:246-248,306-325,407):
    the three grids bit-exact (overlapping rods included: the later record wins), and a step with the fed
    tensor against the direct solve of the system assembled with the oracle's grids."""
    nW, nH = shape
    p, g = make(oracle, nW, nH, **BCS["robin_lr_dir_tb"], h=1.0 / npm)
    rng = np.random.default_rng(17)
    n = 160
    centers = np.c_[rng.uniform(1, p.W - 1, n), rng.uniform(0.5, p.H - 0.5, n)]
    cells = oracle.make_cells(centers, rng.uniform(0, 2 * np.pi, n), (1 + rng.uniform(size=n)) * 2.1, p.W, p.H)
    # the heading setDiffusionTensor sees is the mean body angle, not bodyA's: give some rods another one
    th = rng.uniform(0, 2 * np.pi, n)
    cells[::3, 14], cells[::3, 15] = np.cos(th[::3]), np.sin(th[::3])
    cells = np.vstack([cells, cells[:20]])          # duplicates: full overlap, later record must win
    cells[-20:, 14], cells[-20:, 15] = np.cos(th[:20] + 1.0), np.sin(th[:20] + 1.0)
    Dx, Dy = 1.5, 0.6
    g.upload_cells(cells, npm)
    g.cells_tensor(Dx, Dy)
    ref = oracle.cells_tensor(cells, npm, nH, nW, Dx, Dy)
    got = g.get_tensor()
    for a, b in zip(got, ref):
        assert np.array_equal(a, b)
    assert g.path()["tensor"]
    p.d11, p.d22, p.d12 = ref
    bands, _ = oracle.assemble(p, None)
    x = field(p, 1, smooth=False)
    assert rel(g.apply_operator(x), oracle.band_matvec(p, bands, x)) < 1e-12
    u0 = field(p, 4)
    g.solution_vector[:] = u0
    assert rel(g.stepDiffusion(), oracle.solve_lu(p, u0)) < TOL
    # the shipped scalings Dx = Dy = 1: grids still written (bit-exact), solve stays on the isotropic kernels
    g.cells_tensor(1.0, 1.0)
    for a, b in zip(g.get_tensor(), oracle.cells_tensor(cells, npm, nH, nW, 1.0, 1.0)):
        assert np.array_equal(a, b)
    assert not g.path()["tensor"]
    p.d11 = p.d22 = p.d12 = None
    g.solution_vector[:] = u0
    assert rel(g.stepDiffusion(), oracle.solve_lu(p, u0)) < TOL
    g.close()


def test_tensor_multi_step_with_warm_starts(oracle):
    """The anisotropic path over several steps (rods secrete, the tensor is re-rasterised every step as
    eQabm::updateCells does): every field against the direct solve of the oracle's system, and the history
    candidates (k_init_hist) take over from the cold start."""
    p, g = make(oracle, 161, 97, **BCS["robin_lr"])
    npm = 1.0 / p.h
    cells = oracle.synthetic_colony(120, p.W, p.H, seed=6)
    g.upload_cells(cells, npm)
    p.d11, p.d22, p.d12 = oracle.cells_tensor(cells, npm, p.nH, p.nW, 2.0, 0.5)
    u = np.zeros(p.N)
    its, guesses = [], []
    for k in range(10):
        amount = np.full(len(cells), 100.0 + 2.0 * k)
        u = oracle.scatter(cells, npm, p.nH, p.nW, amount, u)
        u = oracle.solve_lu(p, u)
        g.scatter(amount)
        g.cells_tensor(2.0, 0.5)
        g.step()
        its.append(g.stats().iterations)
        guesses.append(g.last_guess())
        assert g.path()["tensor"]
        assert rel(g.get_field(), u) < TOL, (k, its)
    assert guesses[0] in (0, 1) and all(q in (2, 3, 4, 5) for q in guesses[4:]), guesses
    assert max(its[5:]) < its[0], its
    g.close()


@pytest.mark.parametrize("npm,shape", [(2.0, (201, 41)), (1.0, (101, 21)), (4.0, (401, 81))])
def test_raster_bit_exact(oracle, npm, shape):
    """Cell -> node lookup must be bit-exact (north_star): identical node lists, order included."""
    nW, nH = shape
    h = 1.0 / npm
    W, H = (nW - 1) * h, (nH - 1) * h
    rng = np.random.default_rng(21)
    n = 4000
    # include rods poking through the walls (clamped poles) and tiny rods (empty set fallback)
    centers = np.c_[rng.uniform(-0.5, W + 0.5, n), rng.uniform(-0.5, H + 0.5, n)]
    centers = np.clip(centers, 0.0, [W, H])
    lengths = np.where(rng.uniform(size=n) < 0.1, 1.05, (1 + rng.uniform(size=n)) * 2.1)
    cells = oracle.make_cells(centers, rng.uniform(0, 2 * np.pi, n), lengths, W, H)
    # grown rods: newOffset ratcheted beyond offset (src/abm/cpmEcoli.cpp:407-415)
    cells[::3, 5] += 0.05 * rng.integers(0, 20, len(cells[::3]))
    # nodes exactly on the rectangle edge: axis-aligned rods centred on nodes
    cells[:50, 2], cells[:50, 3] = 1.0, 0.0
    cells[:50, 0] = np.round(cells[:50, 0] * npm) / npm
    cells[:50, 1] = np.round(cells[:50, 1] * npm) / npm
    g = E.GpuHSL(nW, nH, h=h)
    g.upload_cells(cells, npm)
    cnt, nodes = g.raster(cap=256)
    cnt_ref, nodes_ref = oracle.raster(cells, npm, nH, nW, cap=256)
    assert np.array_equal(cnt, cnt_ref)
    assert np.array_equal(nodes, nodes_ref)
    assert cnt.min() >= 1
    # gather is summed in the reference's order -> bit-exact; scatter adds the same per-node amount
    u = rng.uniform(0, 50, nW * nH)
    g.set_field(u)
    assert np.array_equal(g.gather(), oracle.gather(cells, npm, nH, nW, u))
    g.close()


def test_scatter_conserves_and_matches(oracle):
    nW = nH = 512
    p = oracle.Problem(nW=nW, nH=nH)
    cells = oracle.synthetic_colony(1500, p.W, p.H, seed=3)
    g = E.GpuHSL(nW, nH)
    npm = 2.0
    g.upload_cells(cells, npm)
    rng = np.random.default_rng(9)
    amount = rng.uniform(10, 200, len(cells))
    u0 = rng.uniform(0, 5, p.N)
    g.set_field(u0)
    g.scatter(amount)
    ref = oracle.scatter(cells, npm, nH, nW, amount, u0)
    out = g.get_field()
    # rods are separated: each node gets one add -> bit-exact
    assert np.array_equal(out, ref)
    g.close()


def test_full_size_2048_properties(oracle):
    """Config 3 at full size: properties that need no direct solve.
    (a) residual of the returned field in the ORACLE's assembled operator;
    (b) mass conservation under all-Neumann walls (sum M u is invariant);
    (c) oracle CG (independent code) agreement."""
    nW = nH = 2048
    p, g = make(oracle, nW, nH, **BCS["neumann"])
    cells = oracle.synthetic_colony(20000, p.W, p.H)
    npm = 2.0
    g.upload_cells(cells, npm)
    amount = np.full(len(cells), 100.0)
    g.scatter(amount)
    u0 = g.get_field()
    assert np.array_equal(u0, oracle.scatter(cells, npm, nH, nW, amount, np.zeros(p.N)))
    g.step()
    u1 = g.get_field()
    bands, b = oracle.assemble(p, u0)
    r = b - oracle.band_matvec(p, bands, u1)
    assert np.linalg.norm(r) / np.linalg.norm(b) < 1e-11
    # (b) 1^T M u1 = 1^T M u0  (K 1 = 0)
    _, m1 = oracle.assemble(p, u1, want_matrix=False)
    assert abs(m1.sum() - b.sum()) <= 1e-10 * abs(b.sum())
    # (c) Dirichlet-0 variant against the oracle's own CG
    p2, g2 = make(oracle, nW, nH, **BCS["dirichlet0"])
    g2.solution_vector[:] = u0
    out = g2.stepDiffusion()
    ref, it, relres = oracle.solve_cg(p2, u0, rtol=1e-13)
    assert it > 0
    assert rel(out, ref) < TOL
    g.close(); g2.close()


def test_config4_channels_midsize(oracle):
    """BASELINE configs[3] (channel-flow trap) at a size the LU oracle still handles: 769x385 nodes,
    4 steps, Robin left/right + channel Dirichlet top/bottom + 48 CN channel sub-steps per step."""
    rl, rr = oracle.robin_rates(120.0, 1200.0, 20.0, 20.0)
    kw = dict(bc_type=(2, 2, 3, 3), bc_value=(rl, rr, 0, 0), channels=True, channel_v=120.0,
              channel_r=(rl, rr), channel_iters=48, well_scaling=10.0 * (25.0 / 5.0) * 0.5)
    p, g = make(oracle, 769, 385, **kw)
    cells = oracle.synthetic_colony(800, p.W, p.H, seed=4)
    g.upload_cells(cells, 2.0)
    s = oracle.new_state(p)
    for k in range(4):
        amount = 100.0 + 2.0 * oracle.gather(cells, 2.0, p.nH, p.nW, s.u)
        s.u = oracle.scatter(cells, 2.0, p.nH, p.nW, amount, s.u)
        s = oracle.step(p, s)
        g.scatter(100.0 + 2.0 * g.gather())
        g.step()
    t, b = g.channels()
    assert rel(g.get_field(), s.u) < TOL
    assert rel(t, s.top) < TOL and rel(b, s.bottom) < TOL
    g.close()


def test_config4_channels_4096_properties(oracle):
    """configs[3] at the full 4096^2: no direct solve fits, so check the step against the oracle's assembled
    operator (residual), the channel recurrence against the oracle's channel solver fed with the GPU's own
    wall flux, and the one-step lag of the channel Dirichlet rows."""
    n = 4096
    rl, rr = oracle.robin_rates(120.0, 1200.0, 20.0, 20.0)
    kw = dict(bc_type=(2, 2, 3, 3), bc_value=(rl, rr, 0, 0), channels=True, channel_v=120.0,
              channel_r=(rl, rr), channel_iters=48, well_scaling=10.0 * (25.0 / 5.0) * 0.5)
    p, g = make(oracle, n, n, **kw)
    cells = oracle.synthetic_colony(20000, p.W, p.H, seed=6)
    g.upload_cells(cells, 2.0)
    g.scatter(np.full(len(cells), 100.0))
    g.step()
    t1, b1 = g.channels()
    g.scatter(np.full(len(cells), 100.0))
    u_in = g.get_field()
    g.step()
    u_out = g.get_field()
    t2, b2 = g.channels()
    ft, fb = g.channel_flux()
    # (a) the solve: residual of u_out in the oracle's operator with the lagged channel values as Dirichlet data
    bands, b = oracle.assemble(p, u_in)
    r = b - oracle.band_matvec(p, bands, u_out)
    U = u_out.reshape(n, n)
    assert np.array_equal(U[0], b1) and np.array_equal(U[-1], t1)          # rows = previous step's channels
    interior = np.zeros((n, n), bool); interior[1:-1, :] = True
    assert np.linalg.norm(r.reshape(n, n)[interior]) < 1e-10 * np.linalg.norm(b)
    # (b) wall flux (src/fHSL.cpp:54-96) and the 48 CN sub-steps of both channels
    fb_ref, ft_ref = oracle.compute_boundary_flux(p, u_out)
    assert np.allclose(ft, ft_ref, rtol=1e-12, atol=0) and np.allclose(fb, fb_ref, rtol=1e-12, atol=0)
    assert rel(t2, oracle.channel_substeps(p, ft, t1)) < 1e-10
    assert rel(b2, oracle.channel_substeps(p, fb, b1)) < 1e-10
    g.close()


@pytest.mark.parametrize("npm,n", [(2.0, 512), (4.0, 640)])
def test_binned_scatter_matches_direct_and_oracle(oracle, npm, n):
    """writeHSL through shared-memory bins (eqgpu_set_scatter_mode 1): bit-exact for separated rods, and equal
    to the oracle's sequential sum to rounding for a dense overlapping colony; oversize rods take the direct path."""
    h = 1.0 / npm
    W = (n - 1) * h
    rng = np.random.default_rng(31)
    sep = oracle.synthetic_colony(1500, W, W, seed=13)
    g = E.GpuHSL(n, n, h=h)
    g.set_scatter_mode(1)
    u0 = rng.uniform(0, 5, n * n)
    amount = rng.uniform(10, 200, len(sep))
    g.upload_cells(sep, npm)
    g.set_field(u0)
    g.scatter(amount)
    assert np.array_equal(g.get_field(), oracle.scatter(sep, npm, n, n, amount, u0))
    # dense: 6000 overlapping rods in a corner, plus a few rods far longer than a bin window
    m = 6000
    centers = np.c_[rng.uniform(2, W / 4, m), rng.uniform(2, W / 4, m)]
    dense = oracle.make_cells(centers, rng.uniform(0, 2 * np.pi, m), (1 + rng.uniform(size=m)) * 2.1, W, W)
    big = oracle.make_cells([(W / 2, W / 2), (W / 3, W / 2)], [0.3, 1.2], [60.0, 45.0], W, W)
    cells = np.vstack([dense, big])
    amount = rng.uniform(10, 200, len(cells))
    g.upload_cells(cells, npm)
    g.set_field(u0)
    g.scatter(amount)
    ref = oracle.scatter(cells, npm, n, n, amount, u0)
    out = g.get_field()
    assert np.allclose(out, ref, rtol=1e-12, atol=0)
    g.set_scatter_mode(0)
    g.set_field(u0)
    g.scatter(amount)
    assert np.allclose(g.get_field(), ref, rtol=1e-12, atol=0)
    g.close()


@pytest.mark.parametrize("bc", ["robin_lr_dir_tb", "neumann"])
def test_unequal_spacing(oracle, bc):
    """hx != hy (fenicsClassInit rounds the cell counts up separately, src/fHSL.cpp:242-243, so the two mesh
    spacings differ whenever W*npm or H*npm is not an integer)."""
    p, g = make(oracle, 151, 97, h=0.5, hy=0.4, **BCS[bc])
    bands, _ = oracle.assemble(p, None)
    x = field(p, 1, smooth=False)
    assert rel(g.apply_operator(x), oracle.band_matvec(p, bands, x)) < 1e-13
    u0 = field(p, 2)
    g.solution_vector[:] = u0
    assert rel(g.stepDiffusion(), oracle.solve_lu(p, u0)) < TOL
    want = p.D * p.dt * oracle.boundary_functional(p, oracle.solve_lu(p, u0))
    assert abs(g.totalBoundaryFlux - want) <= 1e-7 * max(abs(want), 1.0)
    g.close()


@pytest.mark.parametrize("D,dt", [(1.0, 0.1), (10.0, 0.01), (640.0, 0.1), (1.0e5, 0.1), (3.0e4, 1.0)])
@pytest.mark.parametrize("bc", ["dirichlet0", "neumann", "robin_lr"])
def test_parameter_extremes(oracle, D, dt, bc):
    """Fourier numbers from 0.04 (mass-dominated, single level) to 4e5 (deep hierarchy, high-degree coarse
    solve): the hierarchy depth, the coarse Chebyshev degree and the tail/tile choice all change."""
    p, g = make(oracle, 321, 193, D=D, dt=dt, **BCS[bc])
    u0 = field(p, 5)
    g.solution_vector[:] = u0
    out = g.stepDiffusion()
    st = g.stats()
    assert rel(out, oracle.solve_lu(p, u0)) < TOL, (st.iterations, st.relres, st.levels)
    assert st.iterations <= 40
    g.close()


@pytest.mark.parametrize("bc", ["dirichlet0", "robin_lr"])
def test_warm_start_changes_iterations_not_the_answer(oracle, bc):
    """eqgpu_set_warm_start: previous-solution / extrapolated starting guesses against the direct solve, and
    against the same run with warm starts off.  The stopping test is relative to the right-hand side in
    every mode, so the fields must agree far below TOL while the iteration count drops."""
    nW = nH = 321
    runs = {}
    for mode in (0, 3, 4, 5, 6):
        p, g = make(oracle, nW, nH, **BCS[bc])
        g.set_warm_start(mode)
        cells = oracle.synthetic_colony(400, p.W, p.H, seed=21)
        npm = 1.0 / p.h
        g.upload_cells(cells, npm)
        its, guesses = [], []
        s = oracle.new_state(p)
        for k in range(14):
            amount = np.full(len(cells), 100.0 + 3.0 * k)
            if mode == 0:
                s.u = oracle.scatter(cells, npm, p.nH, p.nW, amount, s.u)
                s = oracle.step(p, s)
            g.scatter(amount)
            g.step()
            its.append(g.stats().iterations)
            guesses.append(g.last_guess())
        runs[mode] = (g.get_field(), its, guesses)
        if mode == 0:
            ref = s.u.copy()
        g.close()
    assert rel(runs[0][0], ref) < TOL and rel(runs[3][0], ref) < TOL
    assert rel(runs[3][0], runs[0][0]) < 1e-10
    assert set(runs[0][2]) <= {0, 1}                      # no history used when warm starts are off
    assert runs[3][2][0] in (0, 1) and runs[3][2][1] in (1, 2)
    assert all(q in (2, 3, 4) for q in runs[3][2][4:]), runs[3][2]
    assert sum(runs[3][1][4:]) < sum(runs[0][1][4:]), (runs[0][1], runs[3][1])
    # mode 4: the least-squares combination of the last three solutions (guess code 5) once two solutions exist;
    # its span contains every mode-3 candidate, so its starting residual is the smallest; the answer is the same
    assert rel(runs[4][0], ref) < TOL and rel(runs[4][0], runs[0][0]) < 1e-10
    assert all(q == 5 for q in runs[4][2][4:]), runs[4][2]
    assert sum(runs[4][1][4:]) < sum(runs[0][1][4:]), (runs[0][1], runs[4][1])
    # mode 5: mode 3 plus the cubic extrapolation of the last four solutions (guess code 6)
    assert rel(runs[5][0], ref) < TOL and rel(runs[5][0], runs[0][0]) < 1e-10
    assert all(q in (2, 3, 4, 6) for q in runs[5][2][4:]) and 6 in runs[5][2], runs[5][2]
    assert sum(runs[5][1][4:]) <= sum(runs[3][1][4:]), (runs[3][1], runs[5][1])
    # mode 6: plus the quartic extrapolation of the last five (guess code 7; its A (h3 - h4) is the A (h2 - h3)
    # the previous step kept, so it becomes available one step after the cubic)
    assert rel(runs[6][0], ref) < TOL and rel(runs[6][0], runs[0][0]) < 1e-10
    assert all(q in (2, 3, 4, 6, 7) for q in runs[6][2][4:]) and 7 in runs[6][2], runs[6][2]
    assert sum(runs[6][1][4:]) <= sum(runs[3][1][4:]), (runs[3][1], runs[6][1])


def test_least_squares_guess_in_steady_state_and_after_wall_changes(oracle):
    """Mode 4 where its Gram matrix degenerates: constant sources until successive solutions coincide to
    solver tolerance (the difference columns vanish into noise), then a jump of the Dirichlet values
    (DIRICHLET_UPDATE: the history holds the old wall values).  Every step still matches the direct solve."""
    p, g = make(oracle, 161, 97, **BCS["dir_values"])
    g.set_warm_start(4)
    rng = np.random.default_rng(2)
    src = np.zeros(p.N)
    src[rng.integers(0, p.N, 40)] = 50.0
    s = oracle.new_state(p)
    its = []
    for k in range(40):
        if k == 30:
            p.bc_value = (4.0, 4.0, 4.0, 4.0)
            g.setBoundaryValues(4.0)
        # steady forcing: add the same source to a field that has stopped changing
        s.u = s.u + src
        g.set_field(g.get_field() + src)
        s = oracle.step(p, s)
        g.step()
        its.append(g.stats().iterations)
        assert rel(g.get_field(), s.u) < TOL, (k, its)
    assert max(its[5:]) <= its[0] + 2, its
    g.close()


def test_warm_start_survives_a_field_reset(oracle):
    """A field that has nothing to do with the history (set_field of noise): the zero guess or the stale
    history is picked on the device by residual norm; the answer still matches the direct solve."""
    p, g = make(oracle, 257, 129, **BCS["dir_values"])
    s = oracle.new_state(p)
    for k in range(3):
        u = field(p, seed=40 + k, smooth=(k != 1)) * (1.0 + 10.0 * k)
        g.set_field(u)
        s.u = u.copy()
        s = oracle.step(p, s)
        g.step()
        assert rel(g.get_field(), s.u) < TOL
    g.close()


@pytest.mark.gpu
def test_nonconvergence_policy_report_and_continue():
    """max_iters too small to reach rtol: policy 0 raises EQGPU_ENOCONV (the default), policy 1 reports and carries on
    with the best iterate -- what the reference does with its solver's diagnostics (src/fHSL.cpp:104-108 has no error
    path); the next, unconstrained solver reaches the tolerance again from that iterate."""
    import eq_b200 as E
    nW, nH = 257, 129
    rng = np.random.default_rng(3)
    u0 = rng.uniform(0.0, 10.0, nW * nH)
    g = E.GpuHSL(nW, nH, device=0, max_iters=2)
    g.set_warm_start(0)
    g.set_field(u0)
    with pytest.raises(E.EqGpuError):
        g.step()
    assert g.unconverged_steps() == 1
    g.set_nonconvergence_policy(1)
    g.set_field(u0)
    g.step()                                   # no exception
    assert g.unconverged_steps() == 2
    st = g.stats()
    assert st.iterations == 2 and st.relres > 1e-12
    assert np.all(np.isfinite(g.get_field()))
    g.close()
