"""The C++ drop-in class gpuHSL (eq_b200/host), driven the way Simulation drives fenicsInterface
(src/simulation.cpp:207,244,466-476), against the oracle's stepDiffusion."""
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "eq_b200", "host", "test_gpuHSL")


def run_case(oracle, tmp_path, kase, W, H, npm, dt, D, nsteps, cells, deposit):
    if not os.path.exists(EXE):
        subprocess.run(["make", "-C", os.path.dirname(EXE), "all"], check=True)
    inp = np.concatenate([[W, H, npm, dt, D, nsteps, len(cells)], cells.ravel(), deposit])
    fin, fout = tmp_path / "in.bin", tmp_path / "out.bin"
    inp.astype(np.float64).tofile(fin)
    subprocess.run([EXE, kase, str(fin), str(fout)], check=True)
    out = np.fromfile(fout)
    nW, nH = int(out[0]), int(out[1])
    N = nW * nH
    o = 3
    res = {"nW": nW, "nH": nH, "iters": int(out[2]), "u": out[o:o + N]}
    o += N
    res["top"], res["bottom"] = out[o:o + nW], out[o + nW:o + 2 * nW]
    o += 2 * nW
    res["flux"] = out[o:o + nsteps]
    o += nsteps
    if len(cells):
        res["gathered"] = out[o:o + len(cells)]
        res["u_fused"] = out[o + len(cells):o + len(cells) + N]
    return res


def rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


CASES = {
    "default": dict(bc_type=(1, 1, 1, 1)),
    "threewall": dict(bc_type=(0, 0, 0, 1)),
    "htrap": dict(bc_type=(2, 2, 0, 0), bc_value=(120.0, 120.0, 0, 0)),
}


@pytest.mark.parametrize("kase", list(CASES))
def test_gpuHSL_matches_oracle(oracle, tmp_path, kase):
    W, H, npm, dt, D, nsteps = 100.0, 20.0, 2.0, 0.1, 1200.0, 4
    p = oracle.Problem(nW=201, nH=41, h=0.5, dt=dt, D=D, **CASES[kase])
    rng = np.random.default_rng(8)
    n = 40
    cells = oracle.make_cells(np.c_[rng.uniform(3, W - 3, n), rng.uniform(3, H - 3, n)], rng.uniform(0, 2 * np.pi, n),
                              (1 + rng.uniform(size=n)) * 2.1, W, H)
    deposit = oracle.scatter(cells, npm, p.nH, p.nW, np.full(n, 100.0), np.zeros(p.N))
    got = run_case(oracle, tmp_path, kase, W, H, npm, dt, D, nsteps, cells, deposit)
    assert (got["nW"], got["nH"]) == (201, 41)
    s = oracle.new_state(p)
    flux = []
    for _ in range(nsteps):
        s.u = s.u + deposit
        s = oracle.step(p, s)
        flux.append(s.total_boundary_flux)
    assert rel(got["u"], s.u) < 1e-8
    assert np.allclose(got["flux"], flux, rtol=1e-7)
    # fused tail of the driver: gather, scatter 100 nM, one resident step
    g = oracle.gather(cells, npm, p.nH, p.nW, s.u)
    assert rel(got["gathered"], g) < 1e-8
    s.u = oracle.scatter(cells, npm, p.nH, p.nW, np.full(n, 100.0), s.u)
    s = oracle.step(p, s)
    assert rel(got["u_fused"], s.u) < 1e-8


def test_gpuHSL_microfluidic_trap_with_channels(oracle, tmp_path):
    W, H, npm, dt, D, nsteps = 100.0, 20.0, 2.0, 0.1, 1200.0, 5
    rl, rr = oracle.robin_rates(120.0, D, 20.0, 20.0)
    p = oracle.Problem(nW=201, nH=41, h=0.5, dt=dt, D=D, bc_type=(2, 2, 3, 3), bc_value=(rl, rr, 0, 0), channels=True,
                       channel_v=120.0, channel_r=(rl, rr), channel_iters=48, well_scaling=10.0 * (25.0 / 5.0) * 0.5)
    rng = np.random.default_rng(9)
    deposit = np.zeros(p.N)
    deposit[rng.integers(0, p.N, 80)] = rng.uniform(10, 100, 80)
    got = run_case(oracle, tmp_path, "channels", W, H, npm, dt, D, nsteps, np.zeros((0, 16)), deposit)
    s = oracle.new_state(p)
    for _ in range(nsteps):
        s.u = s.u + deposit
        s = oracle.step(p, s)
    assert rel(got["u"], s.u) < 1e-8
    assert rel(got["top"], s.top) < 1e-8 and rel(got["bottom"], s.bottom) < 1e-8


@pytest.mark.parametrize("kase", ["fd", "fd_robin"])
def test_gpuFD_matches_fd_oracle(oracle, tmp_path, kase):
    """gpuFD, the drop-in for diffusionPETSc (diffuclass.h:34-97): DIRICHLET_0 as initDiffusion wires it
    (diffuclass.cpp:68-86), and walls rewritten through initData the way upstream intends (:123-124)."""
    W, H, npm, dt, D, nsteps = 100.0, 20.0, 2.0, 0.1, 1200.0, 4
    p = oracle.Problem(nW=201, nH=41, h=0.5, dt=dt, D=D)
    walls = oracle.FDWalls() if kase == "fd" else oracle.FDWalls(Dc=(0.1, 0.02, 1.0, 1.0), Nc=(1.0, 1.0, 0.0, 0.0),
                                                                 BV=(0.03, 0.0, 2.0, 0.5))
    rng = np.random.default_rng(12)
    deposit = np.zeros(p.N)
    deposit[rng.integers(0, p.N, 80)] = rng.uniform(10, 100, 80)
    got = run_case(oracle, tmp_path, kase, W, H, npm, dt, D, nsteps, np.zeros((0, 16)), deposit)
    assert (got["nW"], got["nH"]) == (201, 41)      # gridNodes = length/h + 1 (diffuclass.cpp:358-359)
    u = np.zeros(p.N)
    for _ in range(nsteps):
        u = oracle.fd_solve(p, u + deposit, walls)
    assert rel(got["u"], u) < 1e-8


def test_boundary_well_loop(oracle, tmp_path):
    """DIRICHLET_UPDATE with Simulation's boundary-well model (src/simulation.cpp:581-627): after every step
    the well takes flux/wellScaling, decays by dt*rate, is clipped at zero, and becomes the value of all
    four walls for the next step (setBoundaryValues, src/fHSL.cpp:601-604)."""
    W, H, npm, dt, D, nsteps = 100.0, 20.0, 2.0, 0.1, 1200.0, 6
    rng = np.random.default_rng(14)
    N = 201 * 41
    deposit = np.zeros(N)
    deposit[rng.integers(0, N, 200)] = rng.uniform(100, 1000, 200)
    got = run_case(oracle, tmp_path, "well", W, H, npm, dt, D, nsteps, np.zeros((0, 16)), deposit)
    well_scaling = 2.0 * 10.0 * (15.0 / 5.0) * (W + H)       # :617-618
    rate = 120.0 / W                                          # :609-611
    conc, u, trace = 0.0, np.zeros(N), []
    for _ in range(nsteps):
        p = oracle.Problem(nW=201, nH=41, h=0.5, dt=dt, D=D, bc_type=(1, 1, 1, 1), bc_value=(conc,) * 4)
        s = oracle.State(u=u + deposit, top=np.zeros(201), bottom=np.zeros(201))
        s = oracle.step(p, s)
        u = s.u
        conc += s.total_boundary_flux / well_scaling          # :585-587
        conc -= dt * rate * conc                              # :588-589
        if conc < 0.0:                                        # :591-595
            conc = 0.0
        trace.append(conc)
    assert rel(got["u"], u) < 1e-8
    assert np.allclose(got["flux"], trace, rtol=1e-6, atol=1e-12)
    assert trace[-1] > 0.0


def test_field_snapshot_reads_back_to_the_last_digit(oracle, tmp_path, monkeypatch):
    """gpuHSL::writeDiffusionFiles (the drop-in for fenicsInterface::writeDiffusionFiles, src/fHSL.cpp:630-636, called by
    the controller every recording interval, src/main.cpp:148-162): the legacy-VTK snapshot it writes after the last
    step holds the solution vector in natural node order, 17 significant digits, so reading it back gives the field
    bit for bit; the header carries the mesh."""
    W, H, npm, dt, D, nsteps = 100.0, 20.0, 2.0, 0.1, 1200.0, 3
    p = oracle.Problem(nW=201, nH=41, h=0.5, dt=dt, D=D)
    rng = np.random.default_rng(3)
    deposit = rng.uniform(0.0, 7.0, p.N) * (rng.uniform(size=p.N) < 0.02)
    prefix = str(tmp_path / "hsl_snapshot")
    monkeypatch.setenv("EQ_TEST_SNAPSHOT", prefix)
    got = run_case(oracle, tmp_path, "default", W, H, npm, dt, D, nsteps, np.zeros((0, 16)), deposit)
    path = "%s_%012.4f.vtk" % (prefix, nsteps * dt)
    assert os.path.exists(path), os.listdir(tmp_path)
    lines = open(path).read().split("\n")
    assert lines[0].startswith("# vtk DataFile") and lines[2] == "ASCII" and lines[3] == "DATASET STRUCTURED_POINTS"
    head = {l.split()[0]: l.split()[1:] for l in lines[4:10] if l}
    assert [int(v) for v in head["DIMENSIONS"]] == [201, 41, 1]
    assert [float(v) for v in head["SPACING"]][:2] == [0.5, 0.5]
    assert int(head["POINT_DATA"][0]) == p.N and head["SCALARS"][:2] == ["u", "double"]
    k = lines.index("LOOKUP_TABLE default")
    vals = np.array([float(v) for v in lines[k + 1:k + 1 + p.N]])
    assert np.array_equal(vals, got["u"])                     # bit for bit
    s = oracle.new_state(p)
    for _ in range(nsteps):
        s.u = s.u + deposit
        s = oracle.step(p, s)
    assert rel(vals, s.u) < 1e-8


def test_gpuHSL_row_slabs_two_ranks(oracle, tmp_path):
    """gpuHSL::config::slabRank / slabWorld / slabId: two processes, one GPU each, drive ONE mesh through the
    C++ class the way two MPI ranks of eQ would -- the id made by gpuHSL::makeSlabId on rank 0 reaches rank 1 through a file
    (the stand-in for the controller's MPI_Bcast).  Every rank holds the whole solution_vector, steps, and owns the rows
    eqgpu_slab_plan names; the rows assembled from both ranks equal the oracle's step, and so does the fused tail (rank-summed
    gather, scatter, resident step)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import eq_b200 as E
    if not os.path.exists(EXE):
        subprocess.run(["make", "-C", os.path.dirname(EXE), "all"], check=True)
    W, H, npm, dt, D, nsteps, world = 100.0, 40.0, 2.0, 0.1, 1200.0, 1, 2
    p = oracle.Problem(nW=201, nH=81, h=0.5, dt=dt, D=D, bc_type=(1, 1, 1, 1))
    rng = np.random.default_rng(18)
    n = 60
    cells = oracle.make_cells(np.c_[rng.uniform(3, W - 3, n), rng.uniform(3, H - 3, n)], rng.uniform(0, 2 * np.pi, n),
                              (1 + rng.uniform(size=n)) * 2.1, W, H)
    deposit = oracle.scatter(cells, npm, p.nH, p.nW, np.full(n, 100.0), np.zeros(p.N))
    inp = np.concatenate([[W, H, npm, dt, D, nsteps, len(cells)], cells.ravel(), deposit])
    fin = tmp_path / "in.bin"
    inp.astype(np.float64).tofile(fin)
    idfile = tmp_path / "slab.id"
    procs = []
    for r in range(world):
        env = dict(os.environ, EQ_TEST_SLAB=f"{r},{world},{idfile}", EQGPU_PEER_TIMEOUT_MS="5000")
        procs.append(subprocess.Popen([EXE, "default", str(fin), str(tmp_path / f"out{r}.bin")], env=env))
    for pr in procs:
        assert pr.wait(timeout=180) == 0
    N = p.N
    u, u_fused, gathered = np.zeros(N), np.zeros(N), []
    for r in range(world):
        out = np.fromfile(tmp_path / f"out{r}.bin")
        assert (int(out[0]), int(out[1])) == (201, 81)
        g0, g1, _ = E.slab_plan(p.nH, world, r)[0]
        o = 3
        u[g0 * p.nW:g1 * p.nW] = out[o:o + N][g0 * p.nW:g1 * p.nW]
        o += N + 2 * p.nW + nsteps
        gathered.append(out[o:o + n])
        u_fused[g0 * p.nW:g1 * p.nW] = out[o + n:o + n + N][g0 * p.nW:g1 * p.nW]
    s = oracle.new_state(p)
    s.u = s.u + deposit
    s = oracle.step(p, s)
    assert rel(u, s.u) < 1e-8
    g = oracle.gather(cells, npm, p.nH, p.nW, s.u)
    for r in range(world):
        assert rel(gathered[r], g) < 1e-8          # rank-summed: the same on both ranks
    s.u = oracle.scatter(cells, npm, p.nH, p.nW, np.full(n, 100.0), s.u)
    s = oracle.step(p, s)
    assert rel(u_fused, s.u) < 1e-8
