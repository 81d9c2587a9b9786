"""Warm mode 7 (image ring) over a long run: the TRUE residual ||b - A u|| / ||b||, evaluated through the verification
hooks (eqgpu_build_rhs / eqgpu_apply_operator: an operator walk that shares nothing with the PCG recurrence), must stay
at the solver tolerance step after step.  Round 1 took each new image as b~ - r_final, which fed the images' own error
back through extrapolation weights of absolute sum 127: a CPU model of that recursion reported relres <= 1e-12 while
the true residual grew to 1e-8 after 300 steps and 1e-7 after 1200.  Images are now one operator walk
over the solution, every step."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu


def _true_relres(g, u0, u1, nW, nH):
    b = g.build_rhs(u0)
    Au = g.apply_operator(u1, constrained=False)
    free = np.ones((nH, nW), dtype=bool)
    free[0, :] = free[-1, :] = False
    free[:, 0] = free[:, -1] = False
    f = free.ravel()
    return float(np.linalg.norm((b - Au)[f]) / np.linalg.norm(b[f]))


@pytest.mark.parametrize("mode,colony", [(7, "static"), (7, "moving"), (6, "static")])
def test_true_residual_does_not_drift(mode, colony):
    import eq_b200 as E
    from eq_b200.colony import Colony
    nW = nH = 321
    W = (nW - 1) * 0.5
    col = Colony(300, W, W, mode=colony, seed=17)
    g = E.GpuHSL(nW, nH, device=0)
    g.set_warm_start(mode)
    g.upload_cells(col.records(), 2.0)
    g.set_amounts(np.full(col.n, 100.0))
    worst, zero_iter_steps, guesses = 0.0, 0, set()
    for step in range(1300):
        if colony != "static":
            col.advance()
            g.upload_cells(col.records(), 2.0)
        g.gather_resident()
        g.scatter_resident()
        check = step % 100 == 99 or step >= 1290
        u0 = g.get_field() if check else None
        g.step()
        zero_iter_steps += int(g.stats().iterations == 0)
        guesses.add(g.last_guess())
        if check:
            worst = max(worst, _true_relres(g, u0, g.get_field(), nW, nH))
    g.close()
    # rtol 1e-12 on the recurrence; the true residual may sit a rounding-level multiple above it, never orders above
    assert worst < 2e-11, (worst, zero_iter_steps, guesses)
    if mode == 7:
        assert 8 in guesses      # the ring guess was in use
