"""Generates tests/golden/fd_ref.json from the REFERENCE's own finite-difference solver class.

Run in the build container, where /root/reference exists:
    make -C oracle ref && python tests/golden/make_golden_fd.py
oracle/_ref/libeq_fd_ref.so is /root/reference/diffuclass.{h,cpp} (class diffusionPETSc) compiled in place on
the one-process PETSc/MPI/boost interface shim in oracle/shim_petsc/ (no PETSc installation, no copy of the
sources).  Each case drives the class through its public surface and records: a random x and MyMatMult(x)
(diffuclass.cpp:637-872), a random u0, the right-hand side ApplyBoundaryConditions builds from it (:191-275)
and the field after diffusionPETSc::stepDiffusion (:108-118) with the stand-in Krylov solve run to 1e-13.
The vectors pin oracle/eq_oracle.c (eqo_fd_*) and oracle/oracle.py (fd_*) on machines where the reference tree
is absent (the GPU box).
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

CASES = {
    "dirichlet0_as_shipped": None,      # initDiffusion's own DIRICHLET_0 wiring (diffuclass.cpp:68-86)
    "dirichlet_values": O.FDWalls(Dc=(1, 1, 1, 1), Nc=(0, 0, 0, 0), BV=(1.0, 2.0, 3.0, 4.0)),
    "neumann": O.FDWalls(Dc=(0, 0, 0, 0), Nc=(1, 1, 1, 1), BV=(0, 0, 0, 0)),
    "robin_lr": O.FDWalls(Dc=(0.1, 0.02, 0, 0), Nc=(1, 1, 1, 1), BV=(0.03, 0.0, 0, 0)),
    "robin_lr_dirichlet_tb": O.FDWalls(Dc=(0.1, 0.02, 1, 1), Nc=(1, 1, 0, 0), BV=(0.03, 0.0, 2.0, 0.5)),
    "robin_all_walls": O.FDWalls(Dc=(0.1, 0.02, 0.05, 0.2), Nc=(1, 1, 2, 0.5), BV=(0.03, 0.01, 0.02, 0.04)),
}


def main():
    if O.fd_ref_lib() is None:
        raise SystemExit("oracle/_ref/libeq_fd_ref.so missing: run `make -C oracle ref` where /root/reference exists")
    rng = np.random.default_rng(20261017)
    W, H, npm = 8, 5, 2.0
    out = {"width": W, "height": H, "npm": npm, "cases": []}
    for name, walls in CASES.items():
        for (dt, D) in ((0.1, 1200.0), (0.05, 35.0)):
            ref = O.FDReference(W, H, npm, dt, D, walls)
            u0 = rng.uniform(0, 5, ref.N)
            u1, rhs, its = ref.step(u0, rtol=1e-13)
            x = rng.uniform(-1, 1, ref.N)
            y = ref.matmult(x)
            u2, _, _ = ref.step(u1 + 0.25 * u0, rtol=1e-13)     # a second step: the class keeps no hidden state
            ref.close()
            w = walls or O.FDWalls()
            out["cases"].append({"name": name, "dt": dt, "D": D, "Dc": list(map(float, w.Dc)), "Nc": list(map(float, w.Nc)),
                                 "BV": list(map(float, w.BV)), "x": x.tolist(), "Ax": y.tolist(), "u0": u0.tolist(),
                                 "rhs": rhs.tolist(), "u1": u1.tolist(), "u2": u2.tolist(), "krylov_its": its})
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "fd_ref.json")
    with open(path, "w") as f:
        json.dump(out, f)
    print(f"wrote {path}: {len(out['cases'])} cases, {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
