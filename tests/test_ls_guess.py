"""Warm-start mode 4 (least-squares combination of the last three solutions): the 3x3 solve the device runs
(eqgpu_ls_solve3, host-callable) against numpy's lstsq on Gram data from a real oracle run, and its guards
(missing history, identical solutions, zero field).  The device's pick/true-residual logic is covered by the
GPU tests; this pins the arithmetic that can be pinned without a GPU."""
import numpy as np
import pytest

import eq_b200 as E


def gram(a, b):
    G = [a[0] @ a[0], a[0] @ a[1], a[0] @ a[2], a[1] @ a[1], a[1] @ a[2], a[2] @ a[2]]
    f = [a[0] @ b, a[1] @ b, a[2] @ b]
    return np.array(G), np.array(f), float(b @ b)


def test_ls_solve3_matches_lstsq_on_a_real_history(oracle):
    """History of a small colony run (exact solves): the combination the device would form beats the
    quadratic extrapolation by orders of magnitude and equals numpy's least-squares optimum."""
    import scipy.sparse.linalg as spla
    p = oracle.Problem(nW=161, nH=97)
    cells = oracle.synthetic_colony(60, p.W, p.H, seed=4)
    bands, _ = oracle.assemble(p, None)
    mask, g = oracle.dirichlet(p)
    free = mask == 0
    A = oracle.bands_to_csr(p, bands)[free][:, free].tocsc()
    lu = spla.splu(A)
    u = np.zeros(p.N)
    hist = []
    for step in range(16):
        u0 = oracle.scatter(cells, 2.0, p.nH, p.nW, np.full(len(cells), 100.0), u)
        _, b = oracle.assemble(p, u0, want_matrix=False)
        bf = b[free]
        if step >= 4:
            h0, h1, h2 = hist[-1], hist[-2], hist[-3]
            a = [A @ h0, A @ (h0 - h1), A @ (h1 - h2)]
            G, f, bb = gram(a, bf)
            c, pred = E.ls_solve3(G, f, bb)
            x = c[0] * h0 + c[1] * (h0 - h1) + c[2] * (h1 - h2)
            res_ls = np.linalg.norm(bf - A @ x)
            res_q = np.linalg.norm(bf - A @ (3 * h0 - 3 * h1 + h2))
            c_ref = np.linalg.lstsq(np.array(a).T, bf, rcond=None)[0]
            res_ref = np.linalg.norm(bf - np.array(a).T @ c_ref)
            assert res_ls <= 1.0001 * res_ref + 1e-9 * np.sqrt(bb)
            assert res_ls < res_q
            assert abs(pred - res_ls ** 2) <= 1e-12 * bb + 0.05 * res_ls ** 2
        uf = lu.solve(bf)
        u = np.zeros(p.N)
        u[free] = uf
        hist.append(uf)
    assert res_ls < 1e-2 * res_q     # by step 15 the gap is about three orders of magnitude


def test_ls_solve3_guards():
    rng = np.random.default_rng(0)
    b = rng.normal(size=50)
    a0, a1 = rng.normal(size=50), rng.normal(size=50)
    z = np.zeros(50)
    # only two solutions: third column empty -> plain 2-column least squares, c2 = 0
    G, f, bb = gram([a0, a1, z], b)
    c, pred = E.ls_solve3(G, f, bb)
    ref = np.linalg.lstsq(np.c_[a0, a1], b, rcond=None)[0]
    assert c[2] == 0.0 and np.allclose(c[:2], ref, rtol=1e-9)
    # identical solutions (steady state): difference columns vanish
    G, f, bb = gram([a0, z, z], b)
    c, pred = E.ls_solve3(G, f, bb)
    assert c[1] == 0.0 and c[2] == 0.0 and np.isclose(c[0], (a0 @ b) / (a0 @ a0))
    # no usable column at all (zero field): the guess degenerates to zero with residual ||b||
    G, f, bb = gram([z, z, z], b)
    c, pred = E.ls_solve3(G, f, bb)
    assert np.all(c == 0.0) and np.isclose(pred, bb)
    # exactly dependent columns: the ridge keeps the solve finite and the residual optimal
    G, f, bb = gram([a0, a1, 2.0 * a1], b)
    c, pred = E.ls_solve3(G, f, bb)
    assert np.all(np.isfinite(c))
    r = b - c[0] * a0 - c[1] * a1 - c[2] * 2.0 * a1
    assert np.linalg.norm(r) <= 1.000001 * np.linalg.norm(b - np.c_[a0, a1] @ ref)
    # b in the span: predicted residual clamps at zero instead of going negative
    G, f, bb = gram([a0, a1, z], 2.0 * a0 - 0.5 * a1)
    c, pred = E.ls_solve3(G, f, bb)
    assert pred >= 0.0 and pred <= 1e-12 * bb and np.allclose(c[:2], [2.0, -0.5], rtol=1e-8)
    # NaN input: reported as unusable
    c, pred = E.ls_solve3(np.full(6, np.nan), np.zeros(3), 1.0)
    assert pred >= 1e299 and list(c) == [1.0, 0.0, 0.0]
