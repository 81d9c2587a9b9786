"""The oracle's restated element kernels against the committed golden vectors, which were produced by the
reference's own FFC-generated tabulate_tensor bodies (tests/golden/make_golden.py).  Bit-exact."""
import ctypes as C
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def golden():
    with open(os.path.join(HERE, "golden", "ufc_kernels.json")) as f:
        return json.load(f)["cases"]


def test_golden_file_is_substantial(golden):
    assert len(golden) >= 64


def test_hsld_kernels_bit_exact(oracle, golden):
    L = oracle.lib()
    dp, cd = oracle._dp, C.c_double
    for c in golden:
        xy = np.array(c["xy"])
        d11, d22, d12, u0 = (np.array(c[k]) for k in ("d11", "d22", "d12", "u0"))
        A, b = np.zeros(9), np.zeros(3)
        L.eqo_hsld_cell_a(dp(A), dp(d11), dp(d22), dp(d12), cd(c["D"]), cd(c["dt"]), dp(xy))
        assert A.tolist() == c["cell_a"]                      # fenics/hslD.h:3123-3259
        L.eqo_hsld_cell_L(dp(b), dp(u0), cd(c["dt"]), cd(c["f"]), dp(xy))
        assert b.tolist() == c["cell_L"]                      # fenics/hslD.h:3466-3529
        for facet in range(3):
            L.eqo_hsld_facet_a(dp(A), cd(c["dt"]), cd(c["r"]), dp(xy), C.c_int(facet))
            assert A.tolist() == c["facet_a"][facet]          # fenics/hslD.h:3284-3441
            L.eqo_hsld_facet_L(dp(b), cd(c["dt"]), cd(c["r"]), cd(c["s"]), dp(xy), C.c_int(facet))
            assert b.tolist() == c["facet_L"][facet]          # fenics/hslD.h:3554-3689
            got = L.eqo_boundary_facet(dp(u0), dp(xy), C.c_int(facet))
            assert got == c["boundary"][facet]                # fenics/boundary.h:2652-2741


def test_advection_diffusion_kernels_bit_exact(oracle, golden):
    L = oracle.lib()
    dp, cd = oracle._dp, C.c_double
    for c in golden:
        xc, u2 = np.array(c["xc"]), np.array(c["u2"])
        A, b = np.zeros(4), np.zeros(2)
        L.eqo_ad_cell_a(dp(A), cd(c["dt"]), cd(c["D"]), cd(c["v"]), dp(xc))
        assert A.tolist() == c["ad_cell_a"]                   # fenics/AdvectionDiffusion.h:2246-2289
        L.eqo_ad_cell_L(dp(b), dp(u2), cd(c["dt"]), cd(c["D"]), cd(c["v"]), dp(xc))
        assert b.tolist() == c["ad_cell_L"]                   # fenics/AdvectionDiffusion.h:2444-2510
        for facet in range(2):
            L.eqo_ad_facet_a(dp(A), cd(c["dt"]), cd(c["r"]), C.c_int(facet))
            assert A.tolist() == c["ad_facet_a"][facet]       # :2314-2419
            L.eqo_ad_facet_L(dp(b), dp(u2), cd(c["dt"]), cd(c["r"]), cd(c["s"]), C.c_int(facet))
            assert b.tolist() == c["ad_facet_L"][facet]       # :2535-2648


def test_against_compiled_reference_when_present(oracle):
    """Where oracle/_ref exists (built from /root/reference), compare on fresh random inputs too."""
    R = oracle.ref_lib()
    if R is None:
        pytest.skip("oracle/_ref not built (reference tree absent)")
    L = oracle.lib()
    dp, cd = oracle._dp, C.c_double
    rng = np.random.default_rng(1)
    for _ in range(500):
        xy = rng.uniform(-3, 3, 6)
        d11, d22, d12 = rng.uniform(0.5, 2, 3), rng.uniform(0.5, 2, 3), rng.uniform(-.3, .3, 3)
        D, dt = float(rng.uniform(1, 2000)), float(rng.uniform(0.01, 1))
        A, B = np.zeros(9), np.zeros(9)
        L.eqo_hsld_cell_a(dp(A), dp(d11), dp(d22), dp(d12), cd(D), cd(dt), dp(xy))
        R.ref_hsld_cell_a(dp(B), dp(d11), dp(d22), dp(d12), cd(D), cd(dt), dp(xy))
        assert np.array_equal(A, B)
