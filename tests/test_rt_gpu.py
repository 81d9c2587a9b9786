"""The register-tile smoothers (eq_b200/csrc/smooth_rt.cu: interior tiles in registers, TMA-staged inputs, perimeter tiles on
a tile list) against the shared-memory tile kernels alone: one application of the multigrid preconditioner z = B r to the
same vector through both code paths (EQGPU_RT=1 / 0, read when the solver is created).  A wrong smoother would not
necessarily make PCG fail -- a different SPD preconditioner still converges to the right answer -- so step parity alone
cannot see it; this test can.  Also: B is symmetric (what PCG needs), the shared r.z sum of the two kernels equals <r, B r>,
and a run of steps agrees between the two paths."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu

D, N_, R, C = 1, 0, 2, 3
CASES = [
    # nW, nH, bc_type, bc_value, extra, env
    (2048, 2048, (D, D, D, D), (0, 0, 0, 0), {}, {}),                           # the headline mesh: levels 0 and 1
    (1024, 768, (D, D, D, D), (0, 0, 0, 0), {}, {}),                            # level 0 only
    (1026, 770, (N_, N_, N_, N_), (0, 0, 0, 0), {}, {}),                        # walls are free rows; ragged last tiles
    (1000, 900, (R, R, D, N_), (120.0, 18.8, 0, 0), {}, {}),                    # Robin walls, mixed top/bottom
    (640, 512, (D, N_, D, N_), (0, 0, 0, 0), {}, {"EQGPU_RT_MIN_TILES": "8"}),  # few tiles: more CTAs than tiles
    (1024, 1024, (D, D, D, D), (0, 0, 0, 0), {"discretisation": 1}, {}),        # diffusionPETSc's finite differences (cD = 0)
    (1536, 1100, (D, D, D, D), (0, 0, 0, 0), {}, {"EQGPU_RT_CTAS": "40"}),      # many tiles per persistent CTA
    # the three-CTAs-per-SM instances (right-hand side in shared memory, no box prefetch; opt-in)
    (2048, 2048, (D, D, D, D), (0, 0, 0, 0), {}, {"EQGPU_RT_LEAN": "1"}),
    (1026, 770, (D, D, N_, N_), (0, 0, 0, 0), {}, {"EQGPU_RT_LEAN": "1", "EQGPU_RT_CTAS": "30"}),
]


def _solver(E, rt, nW, nH, bt, bv, extra, env):
    os.environ["EQGPU_RT"] = "1" if rt else "0"
    os.environ.update(env)
    try:
        return E.GpuHSL(nW, nH, bc_type=bt, bc_value=bv, device=0, **extra)
    finally:
        os.environ.pop("EQGPU_RT", None)
        for k in env:
            os.environ.pop(k, None)


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
    nW, nH, bt, bv, extra, env = case
    rng = np.random.default_rng(nW + nH)
    free = _free_mask(nW, nH, bt)
    r = rng.standard_normal(nW * nH) * free
    a = _solver(E, True, nW, nH, bt, bv, extra, env)
    b = _solver(E, False, nW, nH, bt, bv, extra, env)
    assert a.path()["register_tile_levels"] >= 1, a.path()
    assert b.path()["register_tile_levels"] == 0
    za, zb = a.apply_preconditioner(r), b.apply_preconditioner(r)
    assert np.all(np.isfinite(za))
    assert np.linalg.norm(za - zb) <= 1e-12 * np.linalg.norm(zb), np.linalg.norm(za - zb) / np.linalg.norm(zb)
    assert np.all(za[~free] == 0.0)
    # the same vector again: the persistent kernels' tile counters must have been reset
    za2 = a.apply_preconditioner(r)
    assert np.array_equal(za, za2)
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


@pytest.mark.parametrize("size", [(2048, 2048), (1200, 1000)])
def test_steps_agree_with_the_tile_kernels(size):
    import eq_b200 as E
    from eq_b200.colony import Colony
    nW, nH = size
    col = Colony(4000, (nW - 1) * 0.5, (nH - 1) * 0.5, mode="moving", seed=2)
    fields, iters, rel = [], [], []
    for rt in (True, False):
        g = _solver(E, rt, nW, nH, (D, D, D, D), (0, 0, 0, 0), {}, {})
        g.set_warm_start(1)
        g.upload_cells(col.records(), 2.0)
        g.set_amounts(np.full(col.n, 100.0))
        its = []
        for _ in range(4):
            g.gather_resident()
            g.scatter_resident()
            g.step()
            its.append(int(g.stats().iterations))
        rel.append(g.stats().relres)
        fields.append(g.get_field())
        iters.append(its)
        g.close()
    assert iters[0] == iters[1], iters
    assert max(rel) <= 1e-12
    assert np.linalg.norm(fields[0] - fields[1]) <= 1e-10 * np.linalg.norm(fields[1])


def test_true_residual_after_a_step_at_2048():
    """||b - A u|| / ||b|| through eqgpu_apply_operator and eqgpu_build_rhs: the stopping test's r.r is the recurrence's."""
    import eq_b200 as E
    nW = nH = 2048
    g = _solver(E, True, nW, nH, (D, D, D, D), (0, 0, 0, 0), {}, {})
    rng = np.random.default_rng(5)
    u0 = rng.uniform(0, 1, nW * nH) * _free_mask(nW, nH, (D, D, D, D))
    g.set_field(u0)
    b = g.build_rhs(u0)
    g.step()
    u = g.get_field()
    free = _free_mask(nW, nH, (D, D, D, D))
    res = (b - g.apply_operator(u)) * free
    assert np.linalg.norm(res) <= 2e-12 * np.linalg.norm(b * free)
    g.close()
