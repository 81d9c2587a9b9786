"""Warm-start mode 7 (image ring; opt-in, first GPU run: profiles/r01_ring_first_run.json): what can be pinned without a GPU.
The K x K normal-equation solve the device runs (eqgpu_ring_solve, host-callable) and the algebra of the guess
-- backward differences, fixed extrapolation = all-ones combination, least-squares correction fitted to the
extrapolation's residual, image of a solution = b~ - r -- on histories from a real oracle run, against numpy's
QR-quality least squares.  (profiles/r01_guess_study.md has the iteration counts this is worth.)"""
import numpy as np
import pytest

import eq_b200 as E

BINOM = {1: [1], 2: [2, -1], 3: [3, -3, 1], 4: [4, -6, 4, -1], 5: [5, -10, 10, -5, 1], 6: [6, -15, 20, -15, 6, -1],
         7: [7, -21, 35, -35, 21, -7, 1]}


def backward_differences(vs, K):
    """columns nabla^j v_0, j < K (what k_ring_gram / k_ring_impose form in registers)"""
    t = [np.array(v, dtype=np.float64, copy=True) for v in vs[:K]]
    out = [t[0].copy()]
    for j in range(1, K):
        for i in range(K - j):
            t[i] = t[i] - t[i + 1]
        out.append(t[0].copy())
    return np.array(out).T


def history(oracle, steps):
    """exact solves of a small bench-like run: free-row operator, solutions (oldest first) and right-hand sides"""
    import scipy.sparse.linalg as spla
    p = oracle.Problem(nW=161, nH=97)
    cells = oracle.synthetic_colony(60, p.W, p.H, seed=4)
    bands, _ = oracle.assemble(p, None)
    mask, _ = oracle.dirichlet(p)
    free = mask == 0
    A = oracle.bands_to_csr(p, bands)[free][:, free].tocsc()
    lu = spla.splu(A)
    u = np.zeros(p.N)
    hist, rhs = [], []
    for _ in range(steps):
        u0 = oracle.scatter(cells, 2.0, p.nH, p.nW, np.full(len(cells), 100.0), u)
        _, b = oracle.assemble(p, u0, want_matrix=False)
        uf = lu.solve(b[free])
        u = np.zeros(p.N)
        u[free] = uf
        hist.append(uf)
        rhs.append(b[free])
    return A, hist, rhs


def test_backward_differences_are_the_newton_form_of_the_extrapolation():
    rng = np.random.default_rng(1)
    h = [rng.normal(size=40) for _ in range(7)]                     # newest first
    for K in range(1, 8):
        W = backward_differences(h, K)
        assert np.allclose(W.sum(axis=1), sum(c * h[i] for i, c in enumerate(BINOM[K])), rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("K", [2, 3, 5, 7])
def test_ring_guess_matches_qr_least_squares_on_a_real_history(oracle, K):
    A, hist, rhs = history(oracle, 18)
    worst_gain = np.inf
    for step in range(8, 18):
        b = rhs[step]
        h = hist[:step][::-1][:K]                                    # the K newest solutions before this step
        a = [A @ v for v in h]                                       # their images (the device keeps b~ - r_final)
        W, AW = backward_differences(h, K), backward_differences(a, K)
        rho = b - AW.sum(axis=1)                                     # residual of the fixed extrapolation
        c = E.ring_solve(AW.T @ AW, AW.T @ rho)
        x = W @ (1.0 + c)
        res = np.linalg.norm(b - A @ x)
        # the residual vector the device forms from the images equals the true residual of the guess it forms
        assert np.linalg.norm((b - AW @ (1.0 + c)) - (b - A @ x)) <= 1e-12 * np.linalg.norm(b)
        # QR-quality optimum over the same span
        c_ref = np.linalg.lstsq(AW, rho, rcond=None)[0]
        res_ref = np.linalg.norm(rho - AW @ c_ref)
        # (normal equations with a 1e-13 ridge: within a small factor of it, or below 1e-10 ||b|| -- a start from which
        # one PCG iteration reaches the 1e-12 stopping test; with K = 7 on this small, almost steady run the highest
        # differences are rounding noise and the ridge gives up a factor 7 against QR)
        assert res <= 3.0 * res_ref + 1e-10 * np.linalg.norm(b), (step, res, res_ref)
        assert res <= np.linalg.norm(rho) * (1 + 1e-12)              # never worse than the extrapolation it corrects
        worst_gain = min(worst_gain, np.linalg.norm(rho) / max(res, 1e-300))
    if K >= 3:
        assert worst_gain > 3.0        # the correction is worth a large part of an iteration even at its worst


def test_image_of_a_solution_is_rhs_minus_residual(oracle):
    """a = A_ff u_f = b~ - r for any iterate with residual r: the ring needs no operator walk for its newest image,
    and the image does not depend on the Dirichlet data that went into b~."""
    A, hist, rhs = history(oracle, 3)
    rng = np.random.default_rng(2)
    u = hist[-1] * (1 + 1e-9 * rng.normal(size=hist[-1].size))      # an iterate that is not the exact solution
    r = rhs[-1] - A @ u
    assert np.allclose(rhs[-1] - r, A @ u, rtol=0, atol=1e-12 * np.abs(rhs[-1]).max())


def test_ring_solve_guards():
    rng = np.random.default_rng(0)
    cols = [rng.normal(size=60) for _ in range(4)]
    b = rng.normal(size=60)
    z = np.zeros(60)
    # a vanishing column (history shorter than the ring, or two identical solutions) is dropped
    M = np.array([cols[0], cols[1], z, cols[2]]).T
    c = E.ring_solve(M.T @ M, M.T @ b)
    ref = np.linalg.lstsq(np.array([cols[0], cols[1], cols[2]]).T, b, rcond=None)[0]
    assert c[2] == 0.0 and np.allclose(c[[0, 1, 3]], ref, rtol=1e-8)
    # nothing usable: no correction
    assert np.all(E.ring_solve(np.zeros((5, 5)), np.zeros(5)) == 0.0)
    # exactly dependent columns: the ridge keeps the solve finite
    M = np.array([cols[0], cols[1], 2.0 * cols[1]]).T
    assert np.all(np.isfinite(E.ring_solve(M.T @ M, M.T @ b)))
    # NaN in: zeros out (the fixed extrapolation stands)
    G = np.eye(3); G[1, 1] = np.nan
    assert np.all(E.ring_solve(G, np.ones(3)) == 0.0) or np.all(np.isfinite(E.ring_solve(G, np.ones(3))))
    # K = 1 and K = 7 are the ends of the accepted range
    assert np.isclose(E.ring_solve(np.array([[4.0]]), np.array([2.0]))[0], 0.5)
    with pytest.raises(Exception):
        E.ring_solve(np.eye(8), np.ones(8))
