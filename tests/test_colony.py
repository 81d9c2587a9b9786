"""Host-side workload and record plumbing (no GPU): eq_b200/colony.py against the oracle's record builder and the
reference's own cell classes, and eq_b200/host/cellRecords.h instantiated stand-alone."""
import math
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from eq_b200 import colony as Cn   # noqa: E402  (numpy only; does not load the CUDA library)


@pytest.fixture(scope="module")
def oracle():
    from oracle import oracle as O
    O.build()
    return O


def test_fresh_records_equal_the_oracles_bit_for_bit(oracle):
    rng = np.random.default_rng(5)
    n, W, H = 300, 60.0, 35.0
    c = np.c_[rng.uniform(-1, W + 1, n), rng.uniform(-1, H + 1, n)]     # some poles clamp at the trap walls
    a, L = rng.uniform(0, 2 * np.pi, n), (1 + rng.uniform(size=n)) * 2.1
    assert np.array_equal(Cn.make_cells(c, a, L, W, H), oracle.make_cells(c, a, L, W, H))
    # the generator is the oracle's, so the bench colony did not change when it moved out of oracle/
    assert np.array_equal(Cn.synthetic_colony(150, 80.0, 80.0, seed=7), oracle.synthetic_colony(150, 80.0, 80.0, seed=7))


def test_static_colony_never_changes_and_moving_changes_node_sets(oracle):
    W = H = 127.5
    nW = nH = 256
    st = Cn.Colony(400, W, H, mode="static", seed=3)
    r0 = st.records().copy()
    st.advance()
    assert np.array_equal(r0, st.records())
    mv = Cn.Colony(400, W, H, mode="moving", seed=3)
    assert np.array_equal(mv.records(), r0)                              # same layout at step 0
    changed = []
    prev = oracle.raster(mv.records(), 2.0, nH, nW, cap=64)
    for _ in range(20):
        mv.advance()
        cur = oracle.raster(mv.records(), 2.0, nH, nW, cap=64)
        changed.append(np.mean(np.any(cur[1] != prev[1], axis=1)))
        prev = cur
        r = mv.records()
        assert np.all(r[:, 11] >= 3.0 - 1e-12) and np.all(r[:, 11] <= W - 3.0 + 1e-12)
        assert np.array_equal(r[:, 4], r[:, 5])                          # no growth: never ratcheted
    # 0.05-0.2 node per step: a good share of the rods changes at least one node every step, never all of them
    assert 0.05 < np.mean(changed) < 0.9


def test_growing_colony_ratchets_and_divides():
    g = Cn.Colony(500, 200.0, 200.0, mode="growing", seed=9)
    n0, L_prev = g.n, g.L.copy()
    divided = 0
    for _ in range(260):                                                # more than one doubling time (200 steps)
        g.advance()
        r = g.records()
        div = g.L < 0.75 * L_prev
        divided += int(div.sum())
        assert np.allclose(g.L[~div], L_prev[~div] * (1 + g.growth))
        notch = (r[:, 5] - r[:, 4]) / Cn.RATCHET_QUANTUM
        assert np.allclose(notch, np.round(notch), atol=1e-9) and np.all(notch >= -1e-9)
        # the back-filled rectangle lags the true half length by at most quantum + gap (+ one step of growth)
        sep = g.L - g.L0
        assert np.all(sep - (r[:, 5] - r[:, 4]) <= Cn.RATCHET_QUANTUM + Cn.COMPRESSION_GAP + 0.02)
        assert np.all(g.L < Cn.DIVISION_LENGTH * (1 + g.growth))
        # bodyA sits half the separation behind the centre, along the axis
        assert np.allclose(r[:, 11] - r[:, 0], r[:, 2] * 0.5 * sep) and np.allclose(r[:, 12] - r[:, 1], r[:, 3] * 0.5 * sep)
        L_prev = g.L.copy()
    assert g.n == n0 and divided >= n0                                  # every rod divided at least once


def test_grown_rods_match_the_reference_classes(oracle):
    """The reference's own Ecoli / cpmEcoli objects, driven to the body positions the colony model implies and updated
    by cpmEcoli::updateModel + Ecoli::updatePoleCenters, give the records Colony.records() builds (cellRecords.h reads
    them through ref_abm_record)."""
    if oracle.cell_ref_lib() is None:
        pytest.skip("oracle/_ref/libeq_cell_ref.so not built (needs /root/reference)")
    W, H, npm, n = 60, 40, 2.0, 25
    g = Cn.Colony(n, float(W), float(H), npm=npm, mode="growing", seed=21)
    g.L[:] = np.minimum(g.L, 3.0)        # stay below the division length for the 60 steps compared
    g.L0[:] = g.L
    g.new_offset[:] = (g.L0 - g.width) * 0.5
    ref = oracle.ABMReference(W, H, npm)
    for k in range(n):
        ref.add_cell(g.c[k, 0], g.c[k, 1], g.angle[k], g.L[k])
    order = list(range(n))[::-1]         # the reference lists cells newest first
    assert np.allclose(ref.records(), g.records()[order], rtol=0, atol=1e-12)
    worst = 0.0
    for _ in range(60):
        g.advance()
        sep = g.L - g.L0
        for k in range(n):
            d = np.array([math.cos(g.angle[k]), math.sin(g.angle[k])])
            a, b = g.c[k] - d * 0.5 * sep[k], g.c[k] + d * 0.5 * sep[k]
            ref.move_cell(n - 1 - k, (a[0], a[1], g.angle[k]), (b[0], b[1], g.angle[k]), calls=1)
        rr, mine = ref.records(), g.records()[order]
        # everything but the ratchet notch agrees to rounding; a notch may fall one step apart when the separation
        # sits within an ulp of the threshold (the reference measures it as a distance of two points)
        cols = [c for c in range(16) if c != 5]
        worst = max(worst, float(np.max(np.abs(rr[:, cols] - mine[:, cols]))))
        assert np.all(np.abs(rr[:, 5] - mine[:, 5]) <= Cn.RATCHET_QUANTUM + 1e-12)
    assert worst < 1e-11
    assert np.mean(np.abs(ref.records()[:, 5] - g.records()[order][:, 5]) < 1e-12) > 0.9
    ref.close()


def test_cell_records_header_stands_alone(tmp_path):
    exe = tmp_path / "test_cellRecords"
    src = os.path.join(ROOT, "eq_b200", "host", "test_cellRecords.cpp")
    subprocess.run(["g++", "-std=c++14", "-Wall", "-Werror", "-o", str(exe), src], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    rec = np.array([float(v) for v in out])
    want = [3.0, 4.0, 0.6, 0.8, 1.05, 1.35, 0.5, 4.0, 5.0, 1.5, 2.5, 3.09, 4.12, 3.4,
            math.cos(0.9272952180016122), math.sin(0.9272952180016122)]
    assert np.array_equal(rec, np.array(want))
