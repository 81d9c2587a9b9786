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


# --------------------------------------------------------------------------
# diffusionPETSc (diffuclass.cpp): golden vectors produced by the reference's OWN class, compiled in place on the
# interface shim (tests/golden/make_golden_fd.py).
# --------------------------------------------------------------------------
@pytest.fixture(scope="module")
def fd_golden():
    with open(os.path.join(HERE, "golden", "fd_ref.json")) as f:
        return json.load(f)


def _fd_problem(oracle, g, c):
    nW, nH = int(g["width"] * g["npm"]) + 1, int(g["height"] * g["npm"]) + 1      # diffuclass.cpp:358-359
    p = oracle.Problem(nW=nW, nH=nH, h=1.0 / g["npm"], dt=c["dt"], D=c["D"])
    return p, oracle.FDWalls(Dc=tuple(c["Dc"]), Nc=tuple(c["Nc"]), BV=tuple(c["BV"]))


def test_fd_golden_file_is_substantial(fd_golden):
    names = {c["name"] for c in fd_golden["cases"]}
    assert len(fd_golden["cases"]) >= 12 and {"dirichlet0_as_shipped", "neumann", "robin_lr", "robin_all_walls"} <= names


def test_fd_restatements_match_the_reference_class_bit_for_bit(oracle, fd_golden):
    """MyMatMult (diffuclass.cpp:637-872) and ApplyBoundaryConditions (:191-275): the C restatement equals the
    reference's compiled code bit for bit; the numpy sparse restatement gives the same right-hand side bit for
    bit and the same product to rounding (its row sums are ordered differently)."""
    for c in fd_golden["cases"]:
        p, w = _fd_problem(oracle, fd_golden, c)
        x, u0 = np.array(c["x"]), np.array(c["u0"])
        assert oracle.fd_matmult(p, w, x).tolist() == c["Ax"], c["name"]
        assert oracle.fd_rhs_c(p, w, u0).tolist() == c["rhs"], c["name"]
        assert oracle.fd_rhs(p, w, u0).tolist() == c["rhs"], c["name"]
        assert np.allclose(oracle.fd_assemble(p, w) @ x, np.array(c["Ax"]), rtol=1e-14, atol=1e-12)


def test_fd_step_matches_the_reference_class(oracle, fd_golden):
    """diffusionPETSc::stepDiffusion (diffuclass.cpp:108-118) run by the reference class itself (Krylov solve to
    1e-13) against the oracle's exact solve of the restated system, two consecutive steps."""
    for c in fd_golden["cases"]:
        p, w = _fd_problem(oracle, fd_golden, c)
        u0, u1, u2 = np.array(c["u0"]), np.array(c["u1"]), np.array(c["u2"])
        e1 = oracle.fd_solve(p, u0, w)
        assert np.linalg.norm(e1 - u1) <= 1e-10 * np.linalg.norm(u1), c["name"]
        e2 = oracle.fd_solve(p, u1 + 0.25 * u0, w)
        assert np.linalg.norm(e2 - u2) <= 1e-10 * np.linalg.norm(u2), c["name"]
        # ... and the oracle's own copy of the reference's solver at the reference's tolerance
        k5, its, rel5 = oracle.fd_step_krylov(p, u0, w)
        assert rel5 <= 1e-5 and np.linalg.norm(k5 - u1) <= 1e-4 * np.linalg.norm(u1)


def test_fd_live_against_the_compiled_reference_class(oracle):
    """Where oracle/_ref/libeq_fd_ref.so is present (build container; it travels to the GPU box): the default
    trap (100 x 20 um at 2 nodes/um = 201 x 41 nodes) with fresh random data, every wall set."""
    if oracle.fd_ref_lib() is None:
        pytest.skip("oracle/_ref/libeq_fd_ref.so not built (needs /root/reference)")
    rng = np.random.default_rng(7)
    for walls in (None, oracle.FDWalls(Dc=(0.1, 0.02, 1, 1), Nc=(1, 1, 0, 0), BV=(0.03, 0.0, 2.0, 0.5)),
                  oracle.FDWalls(Dc=(0, 0, 0, 0), Nc=(1, 1, 1, 1), BV=(0, 0, 0, 0))):
        ref = oracle.FDReference(100, 20, 2.0, 0.1, 1200.0, walls)
        assert ref.N == 201 * 41
        p = oracle.Problem(nW=201, nH=41)
        w = walls or oracle.FDWalls()
        u0 = rng.uniform(0, 5, p.N)
        u1, rhs, its = ref.step(u0, rtol=1e-13)
        x = rng.uniform(-1, 1, p.N)
        assert np.array_equal(oracle.fd_matmult(p, w, x), ref.matmult(x))
        assert np.array_equal(oracle.fd_rhs_c(p, w, u0), rhs)
        ex = oracle.fd_solve(p, u0, w)
        assert np.linalg.norm(ex - u1) <= 1e-10 * np.linalg.norm(u1)
        u5, _, its5 = ref.step(u0, rtol=1e-5)             # the tolerance the reference actually runs at
        assert 0 < its5 < its and np.linalg.norm(u5 - ex) <= 1e-4 * np.linalg.norm(ex)
        ref.close()


# --------------------------------------------------------------------------
# Cell <-> mesh coupling: golden vectors produced by the reference's OWN eQabm / Ecoli / cpmEcoli classes compiled
# in place on the Chipmunk interface shim (tests/golden/make_golden_cells.py).
# --------------------------------------------------------------------------
@pytest.fixture(scope="module")
def cells_golden():
    with open(os.path.join(HERE, "golden", "cells_ref.json")) as f:
        return json.load(f)


def test_cells_restatement_matches_the_reference_classes_bit_for_bit(oracle, cells_golden):
    """pointIsInCell (src/abm/cpmEcoli.cpp:313-327) on the reference's own rod geometry (fresh, grown, bent and
    ratcheted rods, poles clamped at the walls), and one eQabm::updateCells pass (src/abm/eQabm.cpp:234-425:
    findInteriorPoints, readHSL, writeHSL, setDiffusionTensor in list order, overlapping rods included): the
    oracle's predicate, sequential sample/deposit loop and tensor grids equal the reference's bit for bit."""
    W, H = cells_golden["width"], cells_golden["height"]
    L = oracle.lib()
    assert len(cells_golden["cases"]) >= 3
    for c in cells_golden["cases"]:
        npm = c["npm"]
        nW, nH = int(W * npm) + 1, int(H * npm) + 1
        rec = np.array(c["records"])
        assert np.sum(rec[:, 5] > rec[:, 4]) >= 3                    # ratcheted rods are in the set
        for k, win in enumerate(c["inside"]):
            for di, row in enumerate(win["mask"]):
                for dj, want in enumerate(row):
                    x, y = (win["j0"] + dj) / npm, (win["i0"] + di) / npm
                    got = L.eqo_point_in_cell(oracle._dp(rec[k]), C.c_double(x), C.c_double(y))
                    assert got == want, (npm, k, x, y)
        u1, g = oracle.update_cells_sequential(rec, npm, nH, nW, np.array(c["a0"]), c["a1"], np.array(c["u0"]))
        assert u1.tolist() == c["u1"] and g.tolist() == c["gathered"]
        d11, d22, d12 = oracle.cells_tensor(rec, npm, nH, nW, c["Dx"], c["Dy"])
        assert d11.tolist() == c["d11"] and d22.tolist() == c["d22"] and d12.tolist() == c["d12"]


def test_cells_live_against_the_compiled_reference_classes(oracle, capfd):
    """Where oracle/_ref/libeq_cell_ref.so is present: fresh random colonies through the reference's constructors;
    the oracle's make_cells record equals the one read from the reference objects bit for bit (pole clamps
    included), and so does a coupled updateCells pass."""
    if oracle.cell_ref_lib() is None:
        pytest.skip("oracle/_ref/libeq_cell_ref.so not built (needs /root/reference)")
    rng = np.random.default_rng(11)
    W, H, npm, n = 40, 20, 2.0, 40
    ref = oracle.ABMReference(W, H, npm, 1.5, 0.6)
    xs, ys = rng.uniform(0.5, W - 0.5, n), rng.uniform(0.5, H - 0.5, n)
    an, Ls = rng.uniform(0, 2 * np.pi, n), (1 + rng.uniform(size=n)) * 2.1
    a0 = rng.uniform(50, 150, n)
    for k in range(n):
        ref.add_cell(xs[k], ys[k], an[k], Ls[k], a0[k], 0.5)
    order = list(range(n))[::-1]
    rec = ref.records()
    mine = oracle.make_cells(np.c_[xs, ys][order], an[order], Ls[order], float(W), float(H))
    assert np.array_equal(rec, mine)
    u0 = rng.uniform(0, 5, ref.nW * ref.nH)
    u1, g, tens = ref.update_cells(u0)
    mu, mg = oracle.update_cells_sequential(rec, npm, ref.nH, ref.nW, a0[order], 0.5, u0)
    assert np.array_equal(u1, mu) and np.array_equal(g, mg)
    for a, b in zip(tens, oracle.cells_tensor(rec, npm, ref.nH, ref.nW, 1.5, 0.6)):
        assert np.array_equal(a, b)
    ref.close()


# --------------------------------------------------------------------------
# The P1 step: golden vectors produced by the reference's OWN class fenicsInterface (src/fHSL.cpp) compiled in place on
# the one-process DOLFIN interface shim (tests/golden/make_golden_fenics.py).
# --------------------------------------------------------------------------
@pytest.fixture(scope="module")
def fenics_golden():
    with open(os.path.join(HERE, "golden", "fenics_ref.json")) as f:
        return json.load(f)


def _fenics_problem(oracle, c):
    p = oracle.problem_from_parameters(c["parameters"], c["dt"], c["D"], float(c["width"]), float(c["height"]), c["npm"])
    if "tensor" in c:
        p.d11, p.d22, p.d12 = (np.array(t) for t in c["tensor"])
    return p


def _close(a, b, tol):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    return np.linalg.norm(a - b) <= tol * max(np.linalg.norm(b), 1e-300)


def test_fenics_golden_file_is_substantial(fenics_golden):
    names = {c["name"] for c in fenics_golden["cases"]}
    assert len(fenics_golden["cases"]) >= 28
    assert {"default_trap_nowalled", "dirichlet_update_well", "h_trap_robin", "microfluidic_channels",
            "microfluidic_channels_noflow", "microfluidic_mixed_walls", "anisotropic_tensor"} <= names
    assert {(c["width"], c["height"], c["npm"]) for c in fenics_golden["cases"]} >= {(10, 4, 2.0), (7, 3, 4.0), (6, 5, 1.0)}


def test_fenics_set_up_matches_the_reference_class(oracle, fenics_golden):
    """fenicsClassInit / createHSL / setRobinBoundaryConditions (src/fHSL.cpp:195-364,436-574): node counts, evenly
    spread vertices, identity vertex->dof map and row-major (iy,jx)->dof lookup (the 'bit-exact lookup' of SURVEY 8a
    a12), Robin rates as bound into the trap forms (after the H_TRAP override) and into the channel forms (before
    it), and the channel well volume."""
    for c in fenics_golden["cases"]:
        p = _fenics_problem(oracle, c)
        assert (p.nW, p.nH) == (c["nW"], c["nH"]), c["name"]
        assert c["dof_is_identity"] and c["lookup_is_row_major"]
        assert np.allclose(c["mesh_first_row_x"], np.arange(p.nW) * p.h, rtol=0, atol=1e-15 * c["width"])
        assert np.allclose(c["mesh_first_col_y"], np.arange(p.nH) * (p.hy or p.h), rtol=0, atol=1e-15 * c["height"])
        r = c["robin"]
        assert p.well_scaling == r["well"]
        for w, key in ((0, "trap_left"), (1, "trap_right")):
            if p.bc_type[w] == oracle.ROBIN:
                assert p.bc_value[w] == r[key], c["name"]
        if p.channels:
            assert p.channel_r == (r["chan_left"], r["chan_right"]), c["name"]


def test_fenics_step_matches_the_reference_class(oracle, fenics_golden):
    """fenicsInterface::stepDiffusion (src/fHSL.cpp:98-161) itself -- trap solve on the reference's own element
    kernels, wall flux, channel sub-steps, flux functional -- against oracle.step on the same inputs, every
    boundary / trap type the reference decodes, three geometries.  The only arithmetic that differs is the order of
    the sums in assembly and the elimination order of the direct solve: fields to 1e-11, channels to 1e-10."""
    worst = {"u": 0.0, "chan": 0.0, "flux": 0.0}
    for c in fenics_golden["cases"]:
        p = _fenics_problem(oracle, c)
        s = oracle.new_state(p)
        for st in c["steps"]:
            if "boundary_value" in st:
                p.bc_value = (st["boundary_value"],) * 4             # setBoundaryValues, src/fHSL.cpp:601-604
            s.u = np.array(st["u_in"])
            s = oracle.step(p, s)
            want = np.array(st["u_out"])
            assert _close(s.u, want, 1e-11), (c["name"], c["npm"])
            worst["u"] = max(worst["u"], np.linalg.norm(s.u - want) / np.linalg.norm(want))
            # the functional is a sum with cancellation: measure it against the sum of the magnitudes it adds up
            scale = p.D * p.dt * np.abs(want).sum() / min(p.h, p.hy or p.h)
            assert abs(s.total_boundary_flux - st["total_boundary_flux"]) <= 1e-12 * scale, c["name"]
            worst["flux"] = max(worst["flux"], abs(s.total_boundary_flux - st["total_boundary_flux"]) / scale)
            if p.channels:
                fb, ft = oracle.compute_boundary_flux(p, s.u)
                fscale = np.abs(want).max() * p.D * p.dt / p.well_scaling
                assert np.abs(ft - st["flux_top"]).max() <= 1e-11 * fscale and np.abs(fb - st["flux_bottom"]).max() <= 1e-11 * fscale
                assert _close(s.top, st["top"], 1e-10) and _close(s.bottom, st["bottom"], 1e-10), c["name"]
                worst["chan"] = max(worst["chan"], np.linalg.norm(s.top - st["top"]) / np.linalg.norm(st["top"]))
            else:
                assert not np.any(st["top"]) and not np.any(st["bottom"])
    assert worst["u"] > 0.0     # not a copy: two different eliminations


def test_fenics_live_against_the_compiled_reference_class(oracle):
    """Where oracle/_ref/libeq_fenics_ref.so is present (build container; it travels to the GPU box): the default
    trap as shipped (100 x 20 um at 2 nodes/um = 201 x 41 nodes, DIRICHLET_0 + NOWALLED, src/main.cpp:490,494,511-534)
    and the microfluidic trap with channels at that size, fresh random data, several steps."""
    if oracle.fenics_ref_lib() is None:
        pytest.skip("oracle/_ref/libeq_fenics_ref.so not built (needs /root/reference)")
    rng = np.random.default_rng(5)
    chan = {k: oracle.bc_entry("Dirichlet", -1.0) for k in ("top", "bottom")}
    chan.update({k: oracle.bc_entry("Robin") for k in ("left", "right")})
    for over, steps in ((dict(), 2), (dict(boundaryType="MICROFLUIDIC_TRAP", boundaries=chan,
                                           simulationChannelLengthLeft=20.0, simulationChannelLengthRight=20.0,
                                           channelSolverNumberIterations=6), 2)):
        P = oracle.default_parameters(100, 20, 2.0, **over)
        F = oracle.FenicsReference(P, 0.1, 1200.0, 100.0, 20.0, 2.0)
        assert (F.nW, F.nH) == (201, 41)
        p = oracle.problem_from_parameters(P, 0.1, 1200.0, 100.0, 20.0, 2.0)
        s = oracle.new_state(p)
        u = rng.uniform(0, 50, p.N)
        for _ in range(steps):
            F.set_field(u)
            s.u = u.copy()
            F.step()
            s = oracle.step(p, s)
            uf = F.field()
            assert _close(s.u, uf, 1e-11)
            if p.channels:
                t, b, _, _ = F.channels()
                assert _close(s.top, t, 1e-10) and _close(s.bottom, b, 1e-10)
            u = uf + rng.uniform(0, 5, p.N)
        F.close()


# --------------------------------------------------------------------------
# BASELINE configs[0], the default trap as shipped, run by the reference's own classes on both sides of the
# controller <-> HSL-rank exchange (tests/golden/make_golden_coupled.py): eQabm::updateCells + fenicsInterface::
# stepDiffusion, 20 steps, 32 rods.
# --------------------------------------------------------------------------
@pytest.fixture(scope="module")
def coupled_golden():
    with open(os.path.join(HERE, "golden", "coupled_ref.json")) as f:
        return json.load(f)


def test_coupled_default_trap_matches_the_reference_classes(oracle, coupled_golden):
    """The oracle's cell loop + P1 step against the reference's coupled run: per-cell samples and fields to 1e-10
    over 20 steps (the sample feeds back into the deposit, so an error anywhere would grow), rods grown and bent
    through the reference's ratchet at step 10.  Also the premise of the GPU's fused mode on this colony: sampling
    all rods before depositing equals the reference's sequential loop bit for bit when rods do not overlap."""
    c = coupled_golden
    npm, a1 = c["npm"], c["a1"]
    p = oracle.problem_from_parameters(c["parameters"], c["dt"], c["D"], float(c["width"]), float(c["height"]), npm)
    assert (p.nW, p.nH) == (201, 41) and p.bc_type == (1, 1, 1, 1)
    a0 = np.array(c["a0"])
    rec = np.array(c["records0"])
    centers, ang = rec[:, 11:13], np.arctan2(rec[:, 15], rec[:, 14])
    assert np.array_equal(oracle.make_cells(centers, ang, rec[:, 13], float(c["width"]), float(c["height"])), rec)
    s = oracle.new_state(p)
    for k, st in enumerate(c["steps"], start=1):
        if k == 10:
            rec = np.array(c["records10"])
            assert np.sum(rec[:, 5] > rec[:, 4]) >= 5           # ratcheted rods from here on
        u_seq, g = oracle.update_cells_sequential(rec, npm, p.nH, p.nW, a0, a1, s.u)
        g_all = oracle.gather(rec, npm, p.nH, p.nW, s.u)
        u_all = oracle.scatter(rec, npm, p.nH, p.nW, a0 + a1 * g_all, s.u)
        assert np.array_equal(g, g_all) and np.array_equal(u_seq, u_all)
        assert np.allclose(g, st["gathered"], rtol=1e-10, atol=1e-300), k
        s.u = u_seq
        s = oracle.step(p, s)
        assert abs(s.u.sum() - st["sum"]) <= 1e-10 * abs(st["sum"]) and abs(np.linalg.norm(s.u) - st["norm"]) <= 1e-10 * st["norm"]
        assert abs(s.total_boundary_flux - st["total_boundary_flux"]) <= 1e-12 * p.D * p.dt * np.abs(s.u).sum() / p.h
        if str(k) in c["fields"]:
            want = np.array(c["fields"][str(k)])
            assert np.linalg.norm(s.u - want) <= 1e-10 * np.linalg.norm(want), k
    assert s.u.max() > 1.0     # the colony did build up a signal


def test_three_trap_forms_live_against_the_generated_wrappers(oracle):
    """fenics/hsl.ufl, hslRobin.ufl and hslD.ufl (the north star names all three; fenicsInterface instantiates hslD)
    assembled by the reference's own generated wrappers + kernels on the DOLFIN shim, with the constants the shipped
    run leaves at zero switched ON (Robin external concentrations s, both rates different, a rough tensor):
    the oracle's assembly equals hslD's to rounding (1e-15 of the largest entry), and hsl / hslRobin are hslD with
    the identity tensor without / with the Robin terms -- so the one operator the GPU implements covers all three."""
    if oracle.fenics_ref_lib() is None:
        pytest.skip("oracle/_ref/libeq_fenics_ref.so not built (needs /root/reference)")
    rng = np.random.default_rng(2)
    for (nW, nH, W, H, D, dt) in ((13, 9, 6.0, 4.0, 640.0, 0.05), (21, 9, 10.0, 4.0, 1200.0, 0.1), (8, 15, 7.0, 7.0, 35.0, 0.1)):
        N = nW * nH
        u0 = rng.uniform(0, 5, N)
        rA, sA, rB, sB = 3.5, 0.3, 0.7, 0.1
        p = oracle.Problem(nW=nW, nH=nH, h=W / (nW - 1), hy=H / (nH - 1), dt=dt, D=D, bc_type=(2, 2, 0, 0),
                           bc_value=(rA, rB, 0, 0), robin_s=(sA, sB))
        A, b = oracle.fenics_form_assemble("hslD", nW, nH, W, H, D, dt, rA=rA, sA=sA, rB=rB, sB=sB, u0=u0)
        bands, bo = oracle.assemble(p, u0)
        assert abs(A - oracle.bands_to_csr(p, bands)).max() <= 1e-15 * abs(A).max()
        assert np.abs(b - bo).max() <= 1e-15 * np.abs(b).max()
        A1, b1 = oracle.fenics_form_assemble("hslRobin", nW, nH, W, H, D, dt, rA=rA, sA=sA, rB=rB, sB=sB, u0=u0)
        assert abs(A1 - A).max() <= 1e-15 * abs(A).max() and np.abs(b1 - b).max() <= 1e-15 * np.abs(b).max()
        A0, b0 = oracle.fenics_form_assemble("hsl", nW, nH, W, H, D, dt, u0=u0)
        A2, b2 = oracle.fenics_form_assemble("hslD", nW, nH, W, H, D, dt, u0=u0)
        assert abs(A0 - A2).max() <= 1e-15 * abs(A2).max() and np.abs(b0 - b2).max() <= 1e-15 * np.abs(b2).max()
        assert abs(A - A2).max() > 0         # the Robin terms are really there
        # rough tensor: hslD against the oracle's tensor assembly
        a = rng.uniform(0, np.pi, N)
        t = (1.0 * np.cos(a) ** 2 + 0.2 * np.sin(a) ** 2, 1.0 * np.sin(a) ** 2 + 0.2 * np.cos(a) ** 2, 0.8 * np.sin(a) * np.cos(a))
        p.d11, p.d22, p.d12 = t
        At, bt = oracle.fenics_form_assemble("hslD", nW, nH, W, H, D, dt, rA=rA, sA=sA, rB=rB, sB=sB, tensor=t, u0=u0)
        bands, bo = oracle.assemble(p, u0)
        assert abs(At - oracle.bands_to_csr(p, bands)).max() <= 1e-14 * abs(At).max()
        assert np.abs(bt - bo).max() <= 1e-15 * np.abs(bt).max()


@pytest.mark.parametrize("script,name", [("make_golden_fenics.py", "fenics_ref.json"), ("make_golden_coupled.py", "coupled_ref.json")])
def test_committed_golden_vectors_are_what_the_generators_produce(oracle, tmp_path, script, name):
    """The committed fixtures are reproducible from the reference tree: re-running the committed generator against
    oracle/_ref (the reference's classes compiled in place) gives the committed file, number for number."""
    import subprocess
    import sys
    if oracle.fenics_ref_lib() is None or oracle.cell_ref_lib() is None or not os.path.isdir("/root/reference/src"):
        pytest.skip("needs /root/reference and oracle/_ref")
    out = tmp_path / name
    subprocess.run([sys.executable, os.path.join(HERE, "golden", script), str(out)], check=True, capture_output=True,
                   cwd=os.path.join(HERE, "golden"))
    with open(out) as f, open(os.path.join(HERE, "golden", name)) as g:
        assert json.load(f) == json.load(g)


def test_fenics_fuzz_live_against_the_compiled_reference_class(oracle):
    """60 random configurations through the reference's own fenicsInterface and through the oracle: random integer trap
    sizes, 1 / 2 / 4 nodes per micron, dt, D, flow rate (both branches of the Robin-rate formula), channel lengths,
    sub-step counts, every boundary type with random wall kinds and values, a random tensor on a third of them,
    two steps each.  Needs oracle/_ref (build container)."""
    if oracle.fenics_ref_lib() is None:
        pytest.skip("oracle/_ref/libeq_fenics_ref.so not built (needs /root/reference)")
    rng = np.random.default_rng(20261018)
    kinds = [("Dirichlet", None), ("Neumann", None), ("Robin", None)]
    seen = set()
    for trial in range(60):
        W, H = int(rng.integers(3, 12)), int(rng.integers(2, 8))
        npm = float(rng.choice([1.0, 2.0, 4.0]))
        dt, D = float(rng.choice([0.02, 0.1, 0.5])), float(rng.choice([35.0, 640.0, 1200.0, 3.0e4]))
        btype = str(rng.choice(["DIRICHLET_0", "DIRICHLET_UPDATE", "MICROFLUIDIC_TRAP", "NEUMANN_3WALLED_TEST", "OTHER"]))
        ttype = str(rng.choice(["NOWALLED", "THREEWALLED", "TWOWALLED", "ONEWALLED", "H_TRAP"]))
        over = dict(boundaryType=btype, trapType=ttype, simulationFlowRate=float(rng.choice([0.0, 12.0, 120.0])),
                    simulationChannelLengthLeft=float(rng.uniform(5, 200)), simulationChannelLengthRight=float(rng.uniform(5, 200)),
                    channelSolverNumberIterations=int(rng.integers(1, 7)), lengthScaling=float(rng.choice([1.0, 5.0])))
        if btype == "MICROFLUIDIC_TRAP":
            walls = {}
            for name in ("left", "right", "top", "bottom"):
                k = kinds[int(rng.integers(0, 3))][0]
                v = float(rng.uniform(0, 3)) if k == "Dirichlet" else 0.0
                if k == "Dirichlet" and name in ("top", "bottom") and rng.uniform() < 0.4:
                    v = -1.0                                           # "the channel Function"
                walls[name] = oracle.bc_entry(k, v)
            over["boundaries"] = walls
        P = oracle.default_parameters(W, H, npm, **over)
        if btype == "DIRICHLET_0" and ttype == "H_TRAP":
            # a combination the reference does not define: its four `if`s leave the sub-domain null and DirichletBC is
            # built on it (src/fHSL.cpp:559-568) -- the shim reports it, DOLFIN would dereference it
            with pytest.raises(RuntimeError, match="null sub-domain"):
                oracle.FenicsReference(P, dt, D, float(W), float(H), npm)
            continue
        F = oracle.FenicsReference(P, dt, D, float(W), float(H), npm)
        p = oracle.problem_from_parameters(P, dt, D, float(W), float(H), npm)
        assert (p.nW, p.nH) == (F.nW, F.nH)
        if trial % 3 == 0:
            a = rng.uniform(0, np.pi, p.N)
            p.d11, p.d22, p.d12 = 1.3 * np.cos(a) ** 2 + 0.4 * np.sin(a) ** 2, 1.3 * np.sin(a) ** 2 + 0.4 * np.cos(a) ** 2, 0.9 * np.sin(a) * np.cos(a)
            F.set_tensor(p.d11, p.d22, p.d12)
        s = oracle.new_state(p)
        u = rng.uniform(0, 50, p.N)
        for step in range(2):
            if btype == "DIRICHLET_UPDATE":
                v = float(rng.uniform(0, 5))
                F.set_boundary_value(v)
                p.bc_value = (v,) * 4
            F.set_field(u)
            s.u = u.copy()
            F.step()
            s = oracle.step(p, s)
            uf = F.field()
            assert _close(s.u, uf, 1e-10), (trial, P, step)
            if p.channels:
                t, b, _, _ = F.channels()
                # the wall flux is a difference of neighbouring rows (tiny under a Neumann wall), so the channels inherit
                # the solve's rounding relative to the FIELD's scale: dt D / well * 1e-13 max|u| per step
                tol = 1e-10 * p.dt * p.D / p.well_scaling * np.abs(uf).max()
                assert np.abs(s.top - t).max() <= tol + 1e-9 * np.abs(t).max(), (trial, P, step)
                assert np.abs(s.bottom - b).max() <= tol + 1e-9 * np.abs(b).max(), (trial, P, step)
            scale = p.D * p.dt * np.abs(uf).sum() / min(p.h, p.hy or p.h)
            assert abs(s.total_boundary_flux - F.total_boundary_flux()) <= 1e-11 * scale, (trial, P)
            u = uf + rng.uniform(0, 5, p.N)
        seen.add((btype, p.bc_type, p.channels))
        F.close()
    assert len(seen) >= 15
