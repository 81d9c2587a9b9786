"""Parity at the sizes and over the run lengths the headline is measured on (the multi-step warm-start run
was only ever compared with another mode of the same library).  The checker is the oracle: Jacobi-CG on the assembled
P1 matrix to rtol 1e-13, started from the field BEFORE the step (never from the answer under test), and SuperLU at 1025^2."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu


def rel(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


@pytest.mark.parametrize("mode", [None, 7])   # None: the library default (mode 6 with the adaptive history depth)
def test_2048_moving_colony_40_steps_against_the_oracle(oracle, mode):
    import eq_b200 as E
    from eq_b200.colony import Colony
    nW = nH = 2048
    p = oracle.Problem(nW=nW, nH=nH, h=0.5, dt=0.1, D=1200.0)
    col = Colony(20000, p.W, p.H, mode="moving", seed=12345)
    g = E.GpuHSL(nW, nH, device=0)
    if mode is not None:
        g.set_warm_start(mode)
    amount = np.full(col.n, 100.0)
    g.upload_cells(col.records(), 2.0)
    g.set_amounts(amount)
    worst, its = 0.0, []
    for step in range(40):
        col.advance()
        g.upload_cells(col.records(), 2.0)
        g.gather_resident()
        g.scatter_resident()
        check = step % 10 == 9
        u0 = g.get_field() if check else None
        g.step()
        its.append(int(g.stats().iterations))
        if check:
            ref, _, relres = oracle.solve_cg(p, u0, rtol=1e-13)
            assert relres <= 1e-13
            worst = max(worst, rel(g.get_field(), ref))
    g.close()
    assert worst <= 1e-8, (worst, its)
    assert max(its[5:]) <= 9, its        # a moving colony costs about as much as a cold start, never more


def test_1025_against_superlu(oracle):
    import eq_b200 as E
    from eq_b200.colony import Colony
    nW = nH = 1025
    p = oracle.Problem(nW=nW, nH=nH, h=0.5, dt=0.1, D=1200.0)
    col = Colony(5000, p.W, p.H, mode="static", seed=5)
    g = E.GpuHSL(nW, nH, device=0)
    rng = np.random.default_rng(1)
    g.set_field(rng.uniform(0.0, 40.0, p.N))
    g.upload_cells(col.records(), 2.0)
    g.scatter(np.full(col.n, 100.0))
    u0 = g.get_field()
    g.step()
    ref = oracle.solve_lu(p, u0)
    assert rel(g.get_field(), ref) <= 1e-8
    g.close()
