"""The identities the warm-start kernels rely on (eq_b200/csrc/solver.cu: k_init_tile, k_impose), checked with the
oracle's assembled operator on the CPU: the residual of every extrapolated starting guess is a fixed combination
of r1 = b - A h0 and the history-difference images d_j = A (h_{j-1} - h_j), and d_j of one step is d_{j-1} of the
step before (which is why the quartic candidate needs no operator walk of its own)."""
import subprocess
import sys
import json
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_extrapolation_residual_identities(oracle):
    import ctypes as C
    p = oracle.Problem(nW=41, nH=29, bc_type=(1, 2, 0, 1), bc_value=(1.5, 60.0, 0, 0.5), robin_s=(0.0, 0.2))
    bands, _ = oracle.assemble(p, None)
    mask, g = oracle.dirichlet(p)
    b0 = np.zeros(p.N)
    oracle.lib().eqo_apply_dirichlet_sym(C.c_long(p.nW), C.c_long(p.nH), oracle._dp(bands), oracle._dp(b0),
                                         mask.ctypes.data_as(oracle.c_u8p), oracle._dp(g))
    A = oracle.bands_to_csr(p, bands)
    free = mask == 0
    rng = np.random.default_rng(3)
    h = [np.where(free, rng.normal(size=p.N), 0.0) for _ in range(6)]       # h0 (newest) .. h5, free part only
    b = np.where(free, rng.normal(size=p.N), 0.0)
    r1 = b - A @ h[0]
    d = [None] + [A @ (h[j - 1] - h[j]) for j in range(1, 6)]
    cases = {   # guess coefficients on h0.., residual coefficients on d1.. (k_impose picks 3, 4, 6, 7)
        "linear": ([2, -1], [-1]),
        "quadratic": ([3, -3, 1], [-2, 1]),
        "cubic": ([4, -6, 4, -1], [-3, 3, -1]),
        "quartic": ([5, -10, 10, -5, 1], [-4, 6, -4, 1]),
    }
    for name, (cg, cr) in cases.items():
        x = sum(c * h[k] for k, c in enumerate(cg))
        res = r1 + sum(c * d[k + 1] for k, c in enumerate(cr))
        assert np.allclose(b - A @ x, res, rtol=1e-12, atol=1e-10), name
    # least-squares form 1: u = h0 + c0 h0 + c1 (h0-h1) + c2 (h0-2h1+h2), r = r1 - c0 A h0 - c1 d1 - c2 (d1-d2)
    c = rng.normal(size=3)
    x = h[0] + c[0] * h[0] + c[1] * (h[0] - h[1]) + c[2] * (h[0] - 2 * h[1] + h[2])
    assert np.allclose(b - A @ x, r1 - c[0] * (b - r1) - c[1] * d[1] - c[2] * (d[1] - d[2]), rtol=1e-12, atol=1e-10)
    # one step later the history has moved on by one solution: the new d_j is the old d_{j-1}
    h_next = [np.where(free, rng.normal(size=p.N), 0.0)] + h[:5]
    for j in range(2, 6):
        assert np.array_equal(A @ (h_next[j - 1] - h_next[j]), d[j - 1])


def test_reference_arm_prints_the_contract_line():
    """bench.py --impl reference runs on the host alone (oracle port of the reference's CPU path) and prints one
    JSON line with the keys the driver reads."""
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "3"],
                         capture_output=True, text=True, timeout=300, check=True).stdout.strip().splitlines()[-1]
    line = json.loads(out)
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["value"] > 0 and line["unit"] == "steps/s"
    assert line["cpu_baseline"]["kind"] in ("port", "reference") and line["e2e"]["h2d_bytes_per_step"] == 0
