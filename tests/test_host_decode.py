"""The C++ drop-in class gpuHSL decodes eQ::data::parameters into the C-ABI parameter block exactly as the reference
class decodes them into its forms: `gpuHSL::decodeParameters` (eq_b200/host/gpuHSL.cpp, host arithmetic only -- no
device call, so this runs without a GPU) against the values the reference's OWN fenicsInterface bound
(tests/golden/fenics_ref.json, written by src/fHSL.cpp compiled in place: createHSL :436-574,
setRobinBoundaryConditions :331-364, fenicsClassInit :242-283, initDiffusion :45-47)."""
import json
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "eq_b200", "host", "test_gpuHSL")
BTYPES = ["DIRICHLET_0", "DIRICHLET_UPDATE", "MICROFLUIDIC_TRAP", "NEUMANN_3WALLED_TEST"]
TTYPES = ["NOWALLED", "THREEWALLED", "TWOWALLED", "ONEWALLED", "H_TRAP"]


def config_block(c, nsteps=0):
    """The doubles the driver's "decode" / "golden" cases read (eq_b200/host/test_gpuHSL.cpp header)."""
    P = c["parameters"]
    b = P.get("boundaries")
    walls = [b[w][1] if b else [0.0, 1.0, 0.0] for w in ("left", "right", "top", "bottom")]
    bt = BTYPES.index(P["boundaryType"]) if P["boundaryType"] in BTYPES else 4
    head = [c["width"], c["height"], c["npm"], c["dt"], c["D"], nsteps, 0]
    cfg = [bt, TTYPES.index(P["trapType"])] + [v for w in walls for v in w] + [
        P["lengthScaling"], P["simulationFlowRate"], P["simulationChannelLengthLeft"], P["simulationChannelLengthRight"],
        P["channelSolverNumberIterations"], 1.0 if "tensor" in c else 0.0]
    return np.array(head + cfg, dtype=np.float64)


@pytest.fixture(scope="module")
def exe():
    if not os.path.exists(EXE):
        subprocess.run(["make", "-C", os.path.dirname(EXE), "all"], check=True, stdout=subprocess.DEVNULL)
    return EXE


@pytest.fixture(scope="module")
def fenics_golden():
    with open(os.path.join(ROOT, "tests", "golden", "fenics_ref.json")) as f:
        return json.load(f)


def test_gpuHSL_decodes_parameters_like_the_reference_class(exe, fenics_golden, tmp_path, oracle):
    fin, fout = tmp_path / "in.bin", tmp_path / "out.bin"
    seen = set()
    for c in fenics_golden["cases"]:
        config_block(c).tofile(fin)
        subprocess.run([exe, "decode", str(fin), str(fout)], check=True)
        q = np.fromfile(fout)
        nW, nH, hx, hy, dt, D = q[:6]
        bc_type, bc_value = q[6:10].astype(int), q[10:14]
        channels, iters, chan_v, chan_rl, chan_rr, well, s1, s2 = q[14:22]
        r = c["robin"]
        name = (c["name"], c["npm"])
        # node counts and the evenly spread vertices (fenicsClassInit, RectangleMesh)
        assert (nW, nH) == (c["nW"], c["nH"]), name
        assert np.allclose(np.arange(c["nW"]) * hx, c["mesh_first_row_x"], rtol=0, atol=1e-15 * c["width"])
        assert np.allclose(np.arange(c["nH"]) * hy, c["mesh_first_col_y"], rtol=0, atol=1e-15 * c["height"])
        assert (dt, D) == (c["dt"], c["D"]) and (s1, s2) == (0.0, 0.0)
        # Robin rates: what the reference bound into the trap's forms (after the H_TRAP override) and the channels'
        for w, key in ((0, "trap_left"), (1, "trap_right")):
            if bc_type[w] == oracle.ROBIN:
                assert bc_value[w] == r[key], name
        # same wall types as the oracle's decoding, which the golden steps validate against the reference's solve
        p = oracle.problem_from_parameters(c["parameters"], c["dt"], c["D"], float(c["width"]), float(c["height"]), c["npm"])
        assert tuple(bc_type) == tuple(p.bc_type), name
        assert tuple(bc_value) == tuple(float(v) for v in p.bc_value), name
        assert bool(channels) == p.channels, name
        if p.channels:
            assert (chan_rl, chan_rr) == (r["chan_left"], r["chan_right"]), name
            assert well == r["well"], name       # channel node volume, src/fHSL.cpp:47
            assert int(iters) == c["parameters"]["channelSolverNumberIterations"] and chan_v == c["parameters"]["simulationFlowRate"]
        seen.add(tuple(bc_type))
    assert len(seen) >= 8      # the golden set walks through every branch of the decoding


def test_data_recorder_writes_the_grids(exe, tmp_path):
    """gpuHSL's data-recording constructor + writeDataFiles (the controller's use of the class,
    src/fHSL.cpp:27-34,223-234,637-654; src/simulation.cpp:556-580): node counts ceil(H)*npm + 1, every registered
    grid sampled at the mesh vertices, scalar and vector ranks, values round-trip exactly.  No device call."""
    W, H, npm, t = 12.0, 5.0, 1.0, 30.0
    fin = tmp_path / "in.bin"
    np.array([W, H, npm, 0.1, 1200.0, t, 0], dtype=np.float64).tofile(fin)
    prefix = str(tmp_path) + "/"
    r = subprocess.run([exe, "recorder", str(fin), prefix], check=True, capture_output=True, text=True)
    nW, nH = (int(v) for v in r.stdout.split())
    assert (nW, nH) == (13, 6)

    def read(name):
        lines = open(os.path.join(prefix, f"{name}_{t:012.4f}.vtk")).read().splitlines()
        assert lines[4] == f"DIMENSIONS {nW} {nH} 1" and lines[7] == f"POINT_DATA {nW * nH}"
        start = 9 if lines[8].startswith("VECTORS") else 10
        return lines[8], np.array([[float(v) for v in l.split()] for l in lines[start:start + nW * nH]])

    kind, sc = read("scalarGrid")
    assert kind.startswith("SCALARS")
    i, j = np.mgrid[0:nH, 0:nW]
    assert np.array_equal(sc[:, 0], (1000.0 * i + j + 1.0 / 3.0).ravel())
    kind, ve = read("vectorGrid")
    assert kind.startswith("VECTORS")
    assert np.array_equal(ve[:, 0], i.ravel().astype(float)) and np.array_equal(ve[:, 1], -j.ravel().astype(float))


def test_reference_controller_compiles_against_the_drop_in():
    """src/simulation.{h,cpp} of the reference with INTEGRATION.md's three edits, compiled (syntax and types) against
    gpuHSL on the reference's real src/eQ.h: every member the controller reaches into exists with a compatible type.
    Needs the reference tree (build container)."""
    if not os.path.isdir("/root/reference/src"):
        pytest.skip("/root/reference absent")
    r = subprocess.run(["python", os.path.join(ROOT, "scripts", "check_dropin_compiles.py")], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr


def test_gpuHSL_decoding_fuzz_against_the_oracle(exe, tmp_path, oracle):
    """Random parameter sets through gpuHSL::decodeParameters and through the oracle's decoding (which the live fuzz of
    tests/test_oracle_golden.py holds against the reference class itself): same node counts, wall types, wall values,
    Robin rates and channel wiring, exactly."""
    rng = np.random.default_rng(7)
    fin, fout = tmp_path / "in.bin", tmp_path / "out.bin"
    for trial in range(80):
        W, H = int(rng.integers(3, 200)), int(rng.integers(2, 60))
        npm = float(rng.choice([1.0, 2.0, 4.0]))
        btype = str(rng.choice(BTYPES + ["SOMETHING_ELSE"]))
        ttype = str(rng.choice(TTYPES))
        P = oracle.default_parameters(W, H, npm, boundaryType=btype, trapType=ttype,
                                      simulationFlowRate=float(rng.choice([0.0, 1e-7, 12.0, 120.0])),
                                      simulationChannelLengthLeft=float(rng.uniform(5, 200)),
                                      simulationChannelLengthRight=float(rng.uniform(5, 200)),
                                      channelSolverNumberIterations=int(rng.integers(1, 60)), lengthScaling=float(rng.choice([1.0, 5.0])))
        if btype == "MICROFLUIDIC_TRAP":
            walls = {}
            for name in ("left", "right", "top", "bottom"):
                k = str(rng.choice(["Dirichlet", "Neumann", "Robin"]))
                v = float(rng.uniform(0, 3)) if k == "Dirichlet" else 0.0
                if k == "Dirichlet" and name in ("top", "bottom") and rng.uniform() < 0.4:
                    v = -1.0
                walls[name] = oracle.bc_entry(k, v)
            P["boundaries"] = walls
        c = {"parameters": P, "width": W, "height": H, "npm": npm, "dt": float(rng.choice([0.02, 0.1])), "D": float(rng.choice([35.0, 1200.0]))}
        config_block(c).tofile(fin)
        subprocess.run([exe, "decode", str(fin), str(fout)], check=True)
        q = np.fromfile(fout)
        p = oracle.problem_from_parameters(P, c["dt"], c["D"], float(W), float(H), npm)
        assert (int(q[0]), int(q[1])) == (p.nW, p.nH) and (q[2], q[3]) == (p.h, p.hy), (trial, P)
        assert tuple(q[6:10].astype(int)) == tuple(p.bc_type), (trial, P)
        assert tuple(q[10:14]) == tuple(float(v) for v in p.bc_value), (trial, P)
        assert bool(q[14]) == p.channels, (trial, P)
        if p.channels:
            assert (int(q[15]), q[16], q[17], q[18], q[19]) == (p.channel_iters, p.channel_v, p.channel_r[0], p.channel_r[1], p.well_scaling), (trial, P)
