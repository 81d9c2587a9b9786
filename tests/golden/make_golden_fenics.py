"""Generates tests/golden/fenics_ref.json from the REFERENCE's own P1 solver class, fenicsInterface.

Run in the build container, where /root/reference exists:
    make -C oracle ref && python tests/golden/make_golden_fenics.py
oracle/_ref/libeq_fenics_ref.so is /root/reference/src/fHSL.cpp (with src/fHSL.h, src/Expressions.h and the FFC-generated
form headers fenics/*.h it includes) compiled in place on the one-process DOLFIN interface shim in oracle/shim_dolfin/ (no
DOLFIN installation, no copy of the sources).  Each case builds eQ::data::parameters the way src/main.cpp does for that
boundary / trap type, calls fenicsInterface::initDiffusion, and drives stepDiffusion through its public members
(solution_vector, D11/D22/D12, setBoundaryValues), recording per step: the field handed in (previous solution plus
deposits), the field after the step, totalBoundaryFlux, the channel vectors and the wall fluxes; plus the Robin rates
the reference bound into its forms, its mesh coordinates, its vertex->dof map and its (iy,jx)->dof lookup table.
The vectors pin oracle/oracle.py (`problem_from_parameters`, `step`) and, through tests/test_gpu_parity.py, the CUDA path
on machines where the reference tree is absent (the GPU box).
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402


def walls(left, right, top, bottom):
    return {"left": O.bc_entry(*left), "right": O.bc_entry(*right), "top": O.bc_entry(*top), "bottom": O.bc_entry(*bottom)}


ROBIN, NEU = ("Robin",), ("Neumann",)
CHAN = ("Dirichlet", -1.0)
# (name, parameter overrides, extras)
CASES = [
    ("default_trap_nowalled", dict(boundaryType="DIRICHLET_0", trapType="NOWALLED"), {}),            # src/main.cpp:490,494
    ("dirichlet0_threewalled", dict(boundaryType="DIRICHLET_0", trapType="THREEWALLED"), {}),
    ("dirichlet0_twowalled", dict(boundaryType="DIRICHLET_0", trapType="TWOWALLED"), {}),
    ("dirichlet0_onewalled", dict(boundaryType="DIRICHLET_0", trapType="ONEWALLED"), {}),
    ("dirichlet_update_well", dict(boundaryType="DIRICHLET_UPDATE", trapType="NOWALLED"), {"bval": [0.0, 1.5, 0.4]}),
    ("neumann_3walled_test", dict(boundaryType="NEUMANN_3WALLED_TEST", trapType="NOWALLED"), {}),
    ("unknown_boundary_type", dict(boundaryType="SOMETHING_ELSE", trapType="NOWALLED"), {}),           # src/fHSL.cpp:572-573
    ("h_trap_robin", dict(boundaryType="MICROFLUIDIC_TRAP", trapType="H_TRAP", boundaries=walls(ROBIN, ROBIN, NEU, NEU)), {}),
    ("microfluidic_channels", dict(boundaryType="MICROFLUIDIC_TRAP", trapType="NOWALLED",
                                   boundaries=walls(ROBIN, ROBIN, CHAN, CHAN)), {"steps": 4}),
    ("microfluidic_channels_noflow", dict(boundaryType="MICROFLUIDIC_TRAP", trapType="NOWALLED", simulationFlowRate=0.0,
                                          boundaries=walls(ROBIN, ROBIN, CHAN, CHAN)), {"steps": 3}),  # r = D/L branch, :354-358
    ("microfluidic_mixed_walls", dict(boundaryType="MICROFLUIDIC_TRAP", trapType="NOWALLED",
                                      boundaries=walls(("Dirichlet", 2.0), NEU, ("Dirichlet", 0.5), CHAN)), {}),
    ("microfluidic_robin_left_dirichlet_right", dict(boundaryType="MICROFLUIDIC_TRAP", trapType="NOWALLED",
                                                     boundaries=walls(ROBIN, ("Dirichlet", 1.0), NEU, ("Dirichlet", 0.0))), {}),
    ("anisotropic_tensor", dict(boundaryType="DIRICHLET_0", trapType="NOWALLED"), {"tensor": True}),
    ("anisotropic_tensor_robin", dict(boundaryType="MICROFLUIDIC_TRAP", trapType="H_TRAP", boundaries=walls(ROBIN, ROBIN, NEU, NEU)),
     {"tensor": True}),
]
GEOMETRIES = [(10, 4, 2.0, 0.1, 1200.0), (7, 3, 4.0, 0.05, 640.0), (6, 5, 1.0, 0.1, 35.0)]


def main(path=None):
    if O.fenics_ref_lib() is None:
        raise SystemExit("oracle/_ref/libeq_fenics_ref.so missing: run `make -C oracle ref` where /root/reference exists")
    rng = np.random.default_rng(20261017)
    out = {"cases": []}
    for gi, (W, H, npm, dt, D) in enumerate(GEOMETRIES):
        for name, over, extra in CASES:
            if gi > 0 and name.startswith("dirichlet0_") is False and name not in (
                    "default_trap_nowalled", "microfluidic_channels", "h_trap_robin", "anisotropic_tensor"):
                continue                      # the full list on the first geometry, the main ones on the others
            P = O.default_parameters(W, H, npm, **over)
            F = O.FenicsReference(P, dt, D, float(W), float(H), npm)
            case = {"name": name, "parameters": P, "dt": dt, "D": D, "width": W, "height": H, "npm": npm,
                    "nW": F.nW, "nH": F.nH, "robin": F.robin(), "steps": []}
            xy, dof = F.mesh()
            case["mesh_first_row_x"] = xy[:F.nW, 0].tolist()
            case["mesh_first_col_y"] = xy[::F.nW, 1].tolist()
            case["dof_is_identity"] = bool((dof == np.arange(F.N)).all())
            case["lookup_is_row_major"] = bool((F.lookup() == np.arange(F.N).reshape(F.nH, F.nW)).all())
            if extra.get("tensor"):
                a = rng.uniform(0, np.pi, F.N)
                dx, dy = 1.0, 0.2
                mask = rng.uniform(size=F.N) < 0.5          # half the nodes under cells, the rest isotropic (1, 1, 0)
                d11 = np.where(mask, dx * np.cos(a) ** 2 + dy * np.sin(a) ** 2, 1.0)
                d22 = np.where(mask, dx * np.sin(a) ** 2 + dy * np.cos(a) ** 2, 1.0)
                d12 = np.where(mask, (dx - dy) * np.sin(a) * np.cos(a), 0.0)
                F.set_tensor(d11, d22, d12)
                case["tensor"] = [d11.tolist(), d22.tolist(), d12.tolist()]
            u = rng.uniform(0, 50, F.N)
            for k in range(extra.get("steps", 3)):
                st = {}
                if "bval" in extra:
                    F.set_boundary_value(extra["bval"][k])
                    st["boundary_value"] = extra["bval"][k]
                F.set_field(u)
                F.step()
                uf = F.field()
                t, b, ft, fb = F.channels()
                st.update({"u_in": u.tolist(), "u_out": uf.tolist(), "total_boundary_flux": F.total_boundary_flux(),
                           "top": t.tolist(), "bottom": b.tolist(), "flux_top": ft.tolist(), "flux_bottom": fb.tolist()})
                case["steps"].append(st)
                u = uf + rng.uniform(0, 5, F.N)
            F.close()
            out["cases"].append(case)
    path = path or os.path.join(os.path.dirname(os.path.abspath(__file__)), "fenics_ref.json")
    with open(path, "w") as f:
        json.dump(out, f)
    print(f"wrote {path}: {len(out['cases'])} cases, {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else None)
