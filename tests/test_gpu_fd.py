"""GPU parity tests of the finite-difference discretisation (EQGPU_DISC_FD): the drop-in for the reference's
second eQ::diffusionSolver, diffusionPETSc (diffuclass.cpp), against the oracle's restatement of MyMatMult /
ApplyBoundaryConditions solved exactly.  The reference iterates FBCGSR to PETSc's default rtol 1e-5 [ext];
the bar here is the north star's: rel-L2 <= 1e-8 against the exact solution of the same system."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import eq_b200 as E

TOL = 1e-8


def rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def field(p, seed=0, smooth=True):
    rng = np.random.default_rng(seed)
    if not smooth:
        return rng.uniform(0, 10, p.N)
    y, x = np.mgrid[0:p.nH, 0:p.nW]
    return (5 + 3 * np.sin(2 * np.pi * x / p.nW) * np.cos(np.pi * y / p.nH) + rng.uniform(0, 1, (p.nH, p.nW))).ravel()


def make(oracle, nW, nH, **kw):
    p = oracle.Problem(nW=nW, nH=nH, **kw)
    g = E.GpuHSL(nW, nH, h=p.h, dt=p.dt, D=p.D, bc_type=p.bc_type, bc_value=p.bc_value, robin_s=p.robin_s,
                 discretisation=E.DISC_FD)
    return p, g


BCS = {
    "dirichlet0": dict(bc_type=(1, 1, 1, 1), bc_value=(0, 0, 0, 0)),      # the one wiring diffusionPETSc ships
    "neumann": dict(bc_type=(0, 0, 0, 0), bc_value=(0, 0, 0, 0)),
    "robin_lr": dict(bc_type=(2, 2, 0, 0), bc_value=(120.0, 120.0, 0, 0), robin_s=(0.3, 0.1)),
    "robin_lr_dir_tb": dict(bc_type=(2, 2, 1, 1), bc_value=(138.78, 18.78, 2.0, 0.5), robin_s=(0.3, 0.1)),
    "dir_values": dict(bc_type=(1, 1, 1, 1), bc_value=(1.0, 2.0, 3.0, 4.0)),   # corners: left/right win
    "dir_lr_neumann_tb": dict(bc_type=(1, 1, 0, 0), bc_value=(1.5, 0.5, 0, 0)),  # Dirichlet wins the corners
}


@pytest.mark.parametrize("bc", ["neumann", "robin_lr"])
@pytest.mark.parametrize("shape", [(201, 41), (64, 37), (17, 16)])
def test_fd_operator_and_rhs(oracle, bc, shape):
    """The matrix-free lumped stencil == MyMatMult's rows times the node's cell share (the symmetric form
    PCG works on), and the load vector == ApplyBoundaryConditions(u0) times the same weights."""
    p, g = make(oracle, *shape, **BCS[bc])
    walls = oracle.fd_walls_from_problem(p)
    A = oracle.fd_assemble(p, walls)
    w = oracle.fd_node_weights(p)
    x = field(p, 1, smooth=False)
    assert rel(g.apply_operator(x), w * (A @ x)) < 1e-13
    u0 = field(p, 2)
    assert rel(g.build_rhs(u0), w * oracle.fd_rhs(p, walls, u0)) < 1e-13
    g.close()


@pytest.mark.parametrize("bc", list(BCS))
@pytest.mark.parametrize("shape", [(201, 41), (257, 257), (130, 75)])
def test_fd_step_matches_direct_solve(oracle, bc, shape):
    """diffusionPETSc::stepDiffusion (diffuclass.cpp:108-118) == exact solve of the restated system."""
    p, g = make(oracle, *shape, **BCS[bc])
    u0 = field(p, 3)
    ref = oracle.fd_solve(p, u0)
    g.solution_vector[:] = u0
    out = g.stepDiffusion()
    st = g.stats()
    assert rel(out, ref) < TOL, (st.iterations, st.relres)
    g.close()


def test_fd_multi_step_with_cells(oracle):
    """20 steps of the shipped wiring (DIRICHLET_0) with rods secreting and sampling between the steps
    (warm starts on): fields and per-cell samples against the oracle run."""
    p, g = make(oracle, 201, 41, **BCS["dirichlet0"])
    npm = 1.0 / p.h
    cells = oracle.synthetic_colony(60, p.W, p.H, seed=5)
    g.upload_cells(cells, npm)
    u = np.zeros(p.N)
    for k in range(20):
        amount = np.full(len(cells), 100.0 + k)
        s_ref = oracle.gather(cells, npm, p.nH, p.nW, u)
        s_gpu = g.gather()
        assert rel(s_gpu, s_ref) < TOL or np.linalg.norm(s_ref) == 0
        u = oracle.scatter(cells, npm, p.nH, p.nW, amount, u)
        g.scatter(amount)
        u = oracle.fd_solve(p, u)
        g.step()
        assert rel(g.get_field(), u) < TOL
    g.close()


def test_fd_parameter_extremes(oracle):
    for D, dt in [(1.0, 0.1), (640.0, 0.1), (1.0e5, 0.1)]:
        p, g = make(oracle, 321, 193, D=D, dt=dt, **BCS["robin_lr"])
        u0 = field(p, 5)
        g.solution_vector[:] = u0
        assert rel(g.stepDiffusion(), oracle.fd_solve(p, u0)) < TOL, (D, dt, g.stats().iterations)
        assert g.stats().iterations <= 40
        g.close()


def test_fd_full_size_2048_properties(oracle):
    """BASELINE's 2048^2 mesh: the solved field satisfies the oracle's own matrix to the solver tolerance,
    and under Neumann walls sum(w*u) is conserved (rows of MyMatMult's matrix sum to one)."""
    n = 2048
    for bc in ("dirichlet0", "neumann"):
        p, g = make(oracle, n, n, **BCS[bc])
        walls = oracle.fd_walls_from_problem(p)
        A = oracle.fd_assemble(p, walls)
        u0 = field(p, 9)
        g.set_field(u0)
        g.step()
        u1 = g.get_field()
        b = oracle.fd_rhs(p, walls, u0)
        assert np.linalg.norm(A @ u1 - b) < 1e-10 * np.linalg.norm(b)
        if bc == "neumann":
            w = oracle.fd_node_weights(p)
            assert abs(w @ u1 - w @ u0) < 1e-11 * abs(w @ u0)
        g.close()


def test_fd_rejects_what_the_reference_class_does_not_have(oracle):
    with pytest.raises(E.EqGpuError):
        E.GpuHSL(65, 65, channels=True, bc_type=(2, 2, 3, 3), discretisation=E.DISC_FD)
    g = E.GpuHSL(65, 65, discretisation=E.DISC_FD)
    with pytest.raises(E.EqGpuError):
        g.set_tensor(np.full(65 * 65, 1.5), np.ones(65 * 65), np.zeros(65 * 65))
    g.close()


def test_fd_gpu_against_the_reference_class_itself(oracle):
    """The GPU path against diffusionPETSc ITSELF (oracle/_ref/libeq_fd_ref.so: the reference's diffuclass.cpp
    compiled in place on the one-process PETSc shim, Krylov solve run to 1e-13): five steps of the default
    trap with deposits, as initDiffusion wires it (DIRICHLET_0) and with walls written through initData."""
    if oracle.fd_ref_lib() is None:
        pytest.skip("oracle/_ref/libeq_fd_ref.so not built (needs /root/reference at build time)")
    D = 1200.0
    cases = [
        (None, dict(bc_type=(1, 1, 1, 1), bc_value=(0, 0, 0, 0))),
        (oracle.FDWalls(Dc=(0.1, 0.02, 1, 1), Nc=(1, 1, 0, 0), BV=(0.03, 0.0, 2.0, 0.5)),
         dict(bc_type=(2, 2, 1, 1), bc_value=(D * 0.1, D * 0.02, 2.0, 0.5), robin_s=(0.3, 0.0))),
        (oracle.FDWalls(Dc=(0, 0, 0, 0), Nc=(1, 1, 1, 1), BV=(0, 0, 0, 0)), dict(bc_type=(0, 0, 0, 0), bc_value=(0, 0, 0, 0))),
    ]
    rng = np.random.default_rng(21)
    for walls, bc in cases:
        ref = oracle.FDReference(100, 20, 2.0, 0.1, D, walls)      # 201 x 41 nodes
        g = E.GpuHSL(201, 41, h=0.5, dt=0.1, D=D, discretisation=E.DISC_FD, **bc)
        deposit = np.zeros(201 * 41)
        deposit[rng.integers(0, deposit.size, 60)] = rng.uniform(10, 100, 60)
        u_ref = np.zeros(deposit.size)
        g.solution_vector[:] = 0.0
        for k in range(5):
            u_ref, _, _ = ref.step(u_ref + deposit, rtol=1e-13)
            g.solution_vector[:] = g.solution_vector + deposit
            out = g.stepDiffusion()
            assert rel(out, u_ref) < TOL, (k, bc["bc_type"])
        ref.close()
        g.close()
