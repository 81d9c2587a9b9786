"""The streaming smoothers (eq_b200/csrc/mg_stream.cuh) against the shared-memory tile kernels they replace: one
application of the multigrid preconditioner z = B r to the same vector through both code paths (EQGPU_STREAM=1 / 0,
read when the solver is created).  A wrong smoother would not necessarily make PCG fail -- a different SPD
preconditioner still converges to the right answer -- so step parity alone cannot see it; this test can.  Also: B is
symmetric (what PCG needs) and the iteration counts of a step agree between the two paths."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu

D, N_, R, C = 1, 0, 2, 3
CASES = [
    # nW, nH, bc_type, bc_value, extra
    (512, 384, (D, D, D, D), (0, 0, 0, 0), {}),                       # even pitch: TMA staging
    (640, 200, (N_, N_, N_, N_), (0, 0, 0, 0), {}),                   # walls are free rows
    (257, 129, (D, D, D, D), (0, 0, 0, 0), {}),                       # odd pitch: cp.async staging
    (201, 41, (D, D, D, D), (0, 0, 0, 0), {}),                        # the default trap
    (333, 222, (R, R, D, N_), (120.0, 18.8, 0, 0), {}),               # Robin walls, mixed top/bottom
    (500, 301, (N_, D, N_, D), (0, 0, 0, 0), {}),
    (512, 512, (D, D, D, D), (0, 0, 0, 0), {"discretisation": 1}),    # diffusionPETSc's finite differences
    (1026, 770, (D, N_, D, N_), (0, 0, 0, 0), {}),                    # several strips and chunks, ragged last ones
]


def _solver(E, stream, nW, nH, bt, bv, extra):
    os.environ["EQGPU_STREAM"] = "1" if stream else "0"
    try:
        return E.GpuHSL(nW, nH, bc_type=bt, bc_value=bv, device=0, **extra)
    finally:
        os.environ.pop("EQGPU_STREAM", None)


def _free_mask(nW, nH, bt):
    f = np.ones((nH, nW), dtype=bool)
    if bt[0] == D: f[:, 0] = False
    if bt[1] == D: f[:, -1] = False
    if bt[2] == D: f[-1, :] = False
    if bt[3] == D: f[0, :] = False
    return f.ravel()


@pytest.mark.parametrize("case", CASES, ids=[f"{c[0]}x{c[1]}_{i}" for i, c in enumerate(CASES)])
def test_preconditioner_matches_the_tile_kernels(case):
    import eq_b200 as E
    nW, nH, bt, bv, extra = case
    rng = np.random.default_rng(nW + nH)
    free = _free_mask(nW, nH, bt)
    r = rng.standard_normal(nW * nH) * free
    a = _solver(E, True, nW, nH, bt, bv, extra)
    b = _solver(E, False, nW, nH, bt, bv, extra)
    za, zb = a.apply_preconditioner(r), b.apply_preconditioner(r)
    assert np.all(np.isfinite(za))
    assert np.linalg.norm(za - zb) <= 1e-12 * np.linalg.norm(zb), np.linalg.norm(za - zb) / np.linalg.norm(zb)
    assert np.all(za[~free] == 0.0)
    # symmetry: <B r, s> == <r, B s>
    s = rng.standard_normal(nW * nH) * free
    zs = a.apply_preconditioner(s)
    lhs, rhs = float(za @ s), float(r @ zs)
    assert abs(lhs - rhs) <= 1e-11 * max(abs(lhs), abs(rhs), 1e-300)
    # a smooth vector and a point source as well (the two extremes of the spectrum)
    jj, ii = np.meshgrid(np.arange(nW), np.arange(nH))
    for v in (np.sin(np.pi * jj / (nW - 1)) * np.sin(np.pi * ii / (nH - 1)), (jj == nW // 3) * (ii == nH // 2) * 1.0):
        v = v.ravel() * free
        za, zb = a.apply_preconditioner(v), b.apply_preconditioner(v)
        assert np.linalg.norm(za - zb) <= 1e-12 * np.linalg.norm(zb)
    a.close()
    b.close()


def test_step_iterations_agree_at_2048():
    import eq_b200 as E
    from eq_b200.colony import Colony
    nW = nH = 2048
    col = Colony(5000, (nW - 1) * 0.5, (nH - 1) * 0.5, mode="moving", seed=2)
    fields, iters = [], []
    for stream in (True, False):
        g = _solver(E, stream, nW, nH, (D, D, D, D), (0, 0, 0, 0), {})
        g.set_warm_start(0)
        g.upload_cells(col.records(), 2.0)
        g.set_amounts(np.full(col.n, 100.0))
        its = []
        for _ in range(3):
            g.gather_resident()
            g.scatter_resident()
            g.step()
            its.append(int(g.stats().iterations))
        fields.append(g.get_field())
        iters.append(its)
        g.close()
    assert iters[0] == iters[1], iters
    assert np.linalg.norm(fields[0] - fields[1]) <= 1e-10 * np.linalg.norm(fields[1])
