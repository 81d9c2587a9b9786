"""The CUDA path against the REFERENCE's own P1 solver class: tests/golden/fenics_ref.json holds what
fenicsInterface::stepDiffusion (src/fHSL.cpp:98-161, compiled in place on the DOLFIN interface shim,
tests/golden/make_golden_fenics.py) returned for every boundary / trap type it decodes; here the same inputs go
through the C-ABI -- (1) eq_b200.GpuHSL with the oracle's decoding, (2) the C++ drop-in class gpuHSL with its own
decoding and the host-vector contract (solution_vector in/out, D11/D22/D12, setBoundaryValues) -- and must come back
within the north-star tolerance (relative L2 <= 1e-8 at rtol 1e-12).

Geometries: 21 x 9 and 29 x 13 nodes (the 7 x 6 one of the golden file pins the oracle and the decoding only).
(The file sorts last on purpose: it was written after the round's GPU budget was spent, so its first run is the
driver's round-end run.)"""
import json
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import eq_b200 as E
from test_host_decode import EXE, config_block

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOL = 1e-8

with open(os.path.join(ROOT, "tests", "golden", "fenics_ref.json")) as _f:
    GOLDEN = [c for c in json.load(_f)["cases"] if min(c["nW"], c["nH"]) >= 9]
IDS = [f'{c["name"]}-{c["nW"]}x{c["nH"]}' for c in GOLDEN]


def rel(a, b):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def check_step(c, p, st, u, top, bottom, flux):
    want = np.array(st["u_out"])
    assert rel(u, want) < TOL, c["name"]
    scale = p.D * p.dt * np.abs(want).sum() / min(p.h, p.hy or p.h)     # the functional sums with cancellation
    assert abs(flux - st["total_boundary_flux"]) <= 1e-8 * scale, c["name"]
    if p.channels:
        assert rel(top, st["top"]) < TOL and rel(bottom, st["bottom"]) < TOL, c["name"]


@pytest.mark.parametrize("c", GOLDEN, ids=IDS)
def test_GpuHSL_matches_the_reference_class(oracle, c):
    p = oracle.problem_from_parameters(c["parameters"], c["dt"], c["D"], float(c["width"]), float(c["height"]), c["npm"])
    g = E.GpuHSL(p.nW, p.nH, h=p.h, hy=p.hy, dt=p.dt, D=p.D, bc_type=p.bc_type, bc_value=p.bc_value,
                 robin_s=p.robin_s, channels=p.channels, channel_v=p.channel_v, channel_r=p.channel_r,
                 channel_iters=p.channel_iters, well_scaling=p.well_scaling)
    if "tensor" in c:
        g.set_tensor(*(np.array(t) for t in c["tensor"]))
    for st in c["steps"]:
        if "boundary_value" in st:
            g.setBoundaryValues(st["boundary_value"])
        g.set_field(np.array(st["u_in"]))
        g.step()
        t, b = g.channels() if p.channels else (None, None)
        check_step(c, p, st, g.get_field(), t, b, g.totalBoundaryFlux)
    g.close()


@pytest.mark.parametrize("c", [c for c in GOLDEN if c["nW"] == 21], ids=[i for i, c in zip(IDS, GOLDEN) if c["nW"] == 21])
def test_cpp_gpuHSL_matches_the_reference_class(oracle, tmp_path, c):
    """Simulation's view: the C++ class with the reference's member names, fed the reference's parameter values."""
    if not os.path.exists(EXE):
        subprocess.run(["make", "-C", os.path.dirname(EXE), "all"], check=True)
    p = oracle.problem_from_parameters(c["parameters"], c["dt"], c["D"], float(c["width"]), float(c["height"]), c["npm"])
    parts = [config_block(c, nsteps=len(c["steps"]))]
    if "tensor" in c:
        parts += [np.array(t, dtype=np.float64) for t in c["tensor"]]
    for st in c["steps"]:
        parts += [np.array([st.get("boundary_value", np.nan)]), np.array(st["u_in"], dtype=np.float64)]
    fin, fout = tmp_path / "in.bin", tmp_path / "out.bin"
    np.concatenate(parts).tofile(fin)
    subprocess.run([EXE, "golden", str(fin), str(fout)], check=True)
    out = np.fromfile(fout)
    assert (int(out[0]), int(out[1])) == (p.nW, p.nH)
    o = 2
    for st in c["steps"]:
        u = out[o:o + p.N]
        t, b = out[o + p.N:o + p.N + p.nW], out[o + p.N + p.nW:o + p.N + 2 * p.nW]
        flux = out[o + p.N + 2 * p.nW]
        o += p.N + 2 * p.nW + 1
        check_step(c, p, st, u, t, b, flux)


def test_coupled_default_trap_matches_the_reference_classes(oracle):
    """BASELINE configs[0]: the default trap as shipped (201 x 41 nodes, DIRICHLET_0, 32 rods), 20 coupled steps run by
    the reference's own eQabm::updateCells + fenicsInterface::stepDiffusion (tests/golden/coupled_ref.json), against
    the GPU's fused mode: cell records up, gather, deposit a0 + a1 * sample, resident step -- the field never leaves
    HBM.  Per-cell samples and fields within 1e-8 (the samples feed back into the deposits)."""
    with open(os.path.join(ROOT, "tests", "golden", "coupled_ref.json")) as f:
        c = json.load(f)
    npm, a1, a0 = c["npm"], c["a1"], np.array(c["a0"])
    p = oracle.problem_from_parameters(c["parameters"], c["dt"], c["D"], float(c["width"]), float(c["height"]), npm)
    g = E.GpuHSL(p.nW, p.nH, h=p.h, hy=p.hy, dt=p.dt, D=p.D, bc_type=p.bc_type, bc_value=p.bc_value)
    g.upload_cells(np.array(c["records0"]), npm)
    for k, st in enumerate(c["steps"], start=1):
        if k == 10:
            g.upload_cells(np.array(c["records10"]), npm)
        sampled = g.gather()
        assert rel(sampled, st["gathered"]) < TOL or not np.any(st["gathered"]), k
        g.scatter(a0 + a1 * sampled)
        g.step()
        scale = p.D * p.dt * max(abs(st["sum"]), 1e-300) / p.h
        assert abs(g.totalBoundaryFlux - st["total_boundary_flux"]) <= 1e-8 * scale, k
        if str(k) in c["fields"]:
            assert rel(g.get_field(), c["fields"][str(k)]) < TOL, k
    g.close()


# --------------------------------------------------------------------------------------------------------------
# Warm-start mode 7 (image ring): opt-in, written after this round's GPU budget was spent.  Its kernels ran once, in
# scripts/ring_quick.py (profiles/r01_ring_first_run.json: same field as mode 6, 1.46 against 2.66 iterations per step);
# these tests against the direct solve have not run yet.
# Kept at the very end of the suite so that nothing else hides behind them.
# --------------------------------------------------------------------------------------------------------------
@pytest.mark.timeout(300)
def test_ring_guess_changes_iterations_not_the_answer(oracle):
    """Mode 7 against mode 6 and the direct solve on a bench-like run: same field (the stopping test is relative to
    the right-hand side in every mode), ring guess (code 8) picked once two solutions are stored, no more iterations
    than mode 6 over the run."""
    nW = nH = 321
    runs = {}
    for mode in (6, 7):
        p = oracle.Problem(nW=nW, nH=nH)
        g = E.GpuHSL(nW, nH, h=p.h, dt=p.dt, D=p.D, bc_type=p.bc_type, bc_value=p.bc_value)
        g.set_warm_start(mode)
        cells = oracle.synthetic_colony(400, p.W, p.H, seed=21)
        g.upload_cells(cells, 2.0)
        its, guesses = [], []
        s = oracle.new_state(p)
        for k in range(24):
            amount = np.full(len(cells), 100.0 + 3.0 * k)
            if mode == 6:
                s.u = oracle.scatter(cells, 2.0, p.nH, p.nW, amount, s.u)
                s = oracle.step(p, s)
            g.scatter(amount)
            g.step()
            its.append(g.stats().iterations)
            guesses.append(g.last_guess())
        runs[mode] = (g.get_field(), its, guesses)
        if mode == 6:
            ref = s.u.copy()
        g.close()
    assert rel(runs[6][0], ref) < TOL
    assert rel(runs[7][0], ref) < TOL, runs[7][1:]
    assert rel(runs[7][0], runs[6][0]) < 1e-10
    assert all(q in (0, 1) for q in runs[7][2][:2]) and all(q == 8 for q in runs[7][2][4:]), runs[7][2]
    assert sum(runs[7][1][8:]) <= sum(runs[6][1][8:]), (runs[6][1], runs[7][1])


@pytest.mark.timeout(300)
def test_ring_guess_with_changing_wall_values_and_a_field_reset(oracle):
    """The images A_ff h do not depend on the Dirichlet data, so the ring survives setBoundaryValues; a field that
    has nothing to do with the history is answered from zero or from the field as given, picked on the device by
    residual norm.  Every step matches the direct solve."""
    p = oracle.Problem(nW=161, nH=97, bc_type=(1, 1, 1, 1), bc_value=(1.0, 1.0, 1.0, 1.0))
    g = E.GpuHSL(p.nW, p.nH, h=p.h, dt=p.dt, D=p.D, bc_type=p.bc_type, bc_value=p.bc_value)
    g.set_warm_start(7)
    rng = np.random.default_rng(2)
    src = np.zeros(p.N)
    src[rng.integers(0, p.N, 40)] = 50.0
    s = oracle.new_state(p)
    for k in range(30):
        if k in (12, 20):
            v = 4.0 if k == 12 else 0.5
            p.bc_value = (v, v, v, v)
            g.setBoundaryValues(v)
        if k == 25:
            u = rng.uniform(0, 1000, p.N)                     # a reset: nothing to do with the history
            s.u = u.copy()
            g.set_field(u)
        s.u = s.u + src
        g.set_field(g.get_field() + src)
        s = oracle.step(p, s)
        g.step()
        assert rel(g.get_field(), s.u) < TOL, (k, g.stats().iterations, g.last_guess())
    g.close()
