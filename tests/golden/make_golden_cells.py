"""Generates tests/golden/cells_ref.json from the REFERENCE's own agent-based-model classes.

Run in the build container, where /root/reference exists:
    make -C oracle ref && python tests/golden/make_golden_cells.py
oracle/_ref/libeq_cell_ref.so is /root/reference/src/abm/{eQabm,Ecoli,cpmEcoli,cpmHabitat,cpmTrap}.cpp +
src/Strain.cpp compiled in place on the Chipmunk 7.0.1 interface shim in oracle/shim_cpm/ (Chipmunk is not
vendored upstream: the shim restates its rigid-body transform arithmetic, everything else is inert).  Each case
builds rods through the reference's constructors, moves/grows some of them and lets the reference's own post-step
code run (cpmEcoli::updateModel with its ratchet, Ecoli::updatePoleCenters), then records: the 16-double cell
records read from the reference objects, cpmEcoli::pointIsInCell over a window of nodes around every rod, and the
result of one eQabm::updateCells pass (field after the sequential sample/deposit loop, the per-cell samples, the
D11/D22/D12 grids).  The vectors pin oracle/eq_oracle.c's cell functions where the reference tree is absent.
"""
import contextlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402


@contextlib.contextmanager
def quiet():
    """the reference's constructors print to stdout"""
    fd = os.dup(1)
    with open(os.devnull, "w") as dn:
        os.dup2(dn.fileno(), 1)
        try:
            yield
        finally:
            os.dup2(fd, 1)
            os.close(fd)


def main():
    if O.cell_ref_lib() is None:
        raise SystemExit("oracle/_ref/libeq_cell_ref.so missing: run `make -C oracle ref` where /root/reference exists")
    rng = np.random.default_rng(20261017)
    W, H = 20, 10
    out = {"width": W, "height": H, "cases": []}
    for npm, Dx, Dy, a1 in ((2.0, 1.5, 0.6, 0.25), (4.0, 1.0, 1.0, 0.0), (1.0, 0.7, 2.0, 1.0)):
        with quiet():
            ref = O.ABMReference(W, H, npm, Dx, Dy)
            n = 14
            xs, ys = rng.uniform(1.5, W - 1.5, n), rng.uniform(1.0, H - 1.0, n)
            an, Ls = rng.uniform(0, 2 * np.pi, n), (1 + rng.uniform(size=n)) * 2.1
            xs[0], ys[0], an[0], Ls[0] = 0.7, 0.6, 0.3, 4.0                 # poles clamped at two walls
            xs[1], ys[1], an[1], Ls[1] = W - 0.6, H - 0.5, 2.0, 3.9
            xs[2], ys[2] = xs[3] + 0.4, ys[3] + 0.3                          # overlapping rods: order matters
            a0 = rng.uniform(50, 150, n)
            for k in range(n):
                ref.add_cell(xs[k], ys[k], an[k], Ls[k], a0[k], a1)
            for k in range(n):
                if k % 3 == 0:
                    continue
                r0 = ref.records()[k]
                cx, cy, ang = r0[11], r0[12], np.arctan2(r0[3], r0[2])
                sep, da, db = rng.uniform(0.05, 1.6), rng.uniform(-0.15, 0.15), rng.uniform(-0.15, 0.15)
                ref.move_cell(k, (cx - 0.5 * sep * np.cos(ang), cy - 0.5 * sep * np.sin(ang), ang + da),
                              (cx + 0.5 * sep * np.cos(ang), cy + 0.5 * sep * np.sin(ang), ang + db), calls=int(rng.integers(1, 40)))
            rec = ref.records()
            inside = []
            for k in range(n):
                c = rec[k]
                i0, i1 = int(max(0, (c[12] - 4) * npm)), int(min(H * npm, (c[12] + 4) * npm))
                j0, j1 = int(max(0, (c[11] - 4) * npm)), int(min(W * npm, (c[11] + 4) * npm))
                mask = [[int(ref.point_in_cell(k, j / npm, i / npm)) for j in range(j0, j1 + 1)] for i in range(i0, i1 + 1)]
                inside.append({"i0": i0, "j0": j0, "mask": mask})
            u0 = rng.uniform(0, 5, ref.nW * ref.nH)
            u1, g, (d11, d22, d12) = ref.update_cells(u0)
            ref.close()
        order = list(range(n))[::-1]                                        # list order: newest first
        out["cases"].append({"npm": npm, "Dx": Dx, "Dy": Dy, "a0": a0[order].tolist(), "a1": a1, "records": rec.tolist(),
                             "inside": inside, "u0": u0.tolist(), "u1": u1.tolist(), "gathered": g.tolist(),
                             "d11": d11.tolist(), "d22": d22.tolist(), "d12": d12.tolist()})
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cells_ref.json")
    with open(path, "w") as f:
        json.dump(out, f)
    print(f"wrote {path}: {len(out['cases'])} cases, {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
