"""Generates tests/golden/coupled_ref.json: BASELINE configs[0], the default eQ trap as shipped, run by the
REFERENCE's own classes on both sides of the controller <-> HSL-rank exchange.

Run in the build container, where /root/reference exists:
    make -C oracle ref && python tests/golden/make_golden_coupled.py
The trap is src/main.cpp's default (100 x 20 simulation microns at 2 nodes/um = 201 x 41 nodes, dt = 0.1, D = 1200,
DIRICHLET_0 + NOWALLED, 32 seed cells: :457,490,494,511-534,582).  Per step, as Simulation::stepSimulation orders it
(src/simulation.cpp:413-505): the HSL rank's fenicsInterface::stepDiffusion (oracle/_ref/libeq_fenics_ref.so) solves,
its solution_vector goes to the controller, eQabm::updateCells (oracle/_ref/libeq_cell_ref.so) lets every cell sample
it and deposit into it, and the vector goes back.  The MPI Isend/Irecv of the vector (src/simulation.cpp:428-432,
455-463,491-505) is a copy here; the gene circuit is the linear stand-in of oracle/cell_ref.cpp (deposit
a0 + a1 * sample, a1 < 0: secretion falls as the signal builds up); rods do not move (Chipmunk's space is inert in
the shim) except that a third of them is grown and bent once, at step 10, through the reference's own ratchet.
Rods are placed pairwise separated, so the reference's sequential sample/deposit loop and a sample-all-then-deposit-all
pass (the GPU's fused mode) see the same values.
Recorded: the cell records at step 0 and after step 10, per step the per-cell samples and the field's sum and 2-norm,
the full field after steps 1, 10 and 20.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from make_golden_cells import quiet  # noqa: E402

W, H, NPM, DT, D, NCELLS, NSTEPS = 100, 20, 2.0, 0.1, 1200.0, 32, 20


def main(path=None):
    if O.cell_ref_lib() is None or O.fenics_ref_lib() is None:
        raise SystemExit("oracle/_ref libraries missing: run `make -C oracle ref` where /root/reference exists")
    rng = np.random.default_rng(20261017)
    colony = O.synthetic_colony(NCELLS, float(W), float(H), seed=4242, min_clear=1.5, margin=3.0)
    n = len(colony)
    assert n == NCELLS
    a0 = rng.uniform(50, 150, n)
    a1 = -0.3
    with quiet():
        abm = O.ABMReference(W, H, NPM, 1.0, 1.0)
        for k in range(n):   # centre, angle, length of the oracle's record -> the reference's own constructor
            c = colony[k]
            abm.add_cell(c[11], c[12], float(np.arctan2(c[15], c[14])), c[13], a0[k], a1)
    order = list(range(n))[::-1]                      # the reference walks its list newest first
    P = O.default_parameters(W, H, NPM)
    F = O.FenicsReference(P, DT, D, float(W), float(H), NPM)
    out = {"width": W, "height": H, "npm": NPM, "dt": DT, "D": D, "parameters": P, "a0": a0[order].tolist(), "a1": a1,
           "records0": abm.records().tolist(), "steps": [], "fields": {}}
    u = F.field()                                     # zeros: src/fHSL.cpp:583-584
    for k in range(1, NSTEPS + 1):
        if k == 10:
            with quiet():
                for c in range(0, n, 3):
                    r0 = abm.records()[c]
                    cx, cy, ang = r0[11], r0[12], np.arctan2(r0[3], r0[2])
                    sep = 0.35
                    abm.move_cell(c, (cx - 0.5 * sep * np.cos(ang), cy - 0.5 * sep * np.sin(ang), ang + 0.05),
                                  (cx + 0.5 * sep * np.cos(ang), cy + 0.5 * sep * np.sin(ang), ang - 0.04), calls=12)
            out["records10"] = abm.records().tolist()
        with quiet():
            u, g, _ = abm.update_cells(u)             # controller: every cell samples and deposits
        F.set_field(u)                                # MPI transfer controller -> HSL rank
        F.step()                                      # HSL rank: fenicsInterface::stepDiffusion
        u = F.field()                                 # MPI transfer HSL rank -> controller
        out["steps"].append({"gathered": g.tolist(), "sum": float(u.sum()), "norm": float(np.linalg.norm(u)),
                             "total_boundary_flux": F.total_boundary_flux()})
        if k in (1, 10, NSTEPS):
            out["fields"][str(k)] = u.tolist()
    F.close()
    abm.close()
    path = path or os.path.join(os.path.dirname(os.path.abspath(__file__)), "coupled_ref.json")
    with open(path, "w") as f:
        json.dump(out, f)
    print(f"wrote {path}: {NSTEPS} steps, {n} cells, {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else None)
