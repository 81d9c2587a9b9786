"""Generates tests/golden/ufc_kernels.json from the REFERENCE's own FFC-generated element kernels.

Run in the build container, where /root/reference exists:
    make -C oracle ref && python tests/golden/make_golden.py
oracle/_ref/libeq_ufc_ref.so is /root/reference/fenics/{hslD,AdvectionDiffusion,boundary}.h compiled in
place under the interface shim in oracle/shim/ (no FEniCS installation, no copy of the sources).  The
vectors pin oracle/eq_oracle.c bit-for-bit on machines where the reference tree is absent (the GPU box).
"""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402


def main():
    R = O.ref_lib()
    if R is None:
        raise SystemExit("oracle/_ref/libeq_ufc_ref.so missing: run `make -C oracle ref` where /root/reference exists")
    dp = lambda a: a.ctypes.data_as(O.c_dp)
    cd = C.c_double
    rng = np.random.default_rng(20261017)
    cases = []
    for t in range(64):
        h = float(rng.choice([0.25, 0.5, 1.0, 0.37]))
        x0, y0 = (float(v) for v in rng.uniform(0, 50, 2))
        if t % 4 == 0:
            xy = np.array([x0, y0, x0 + h, y0, x0 + h, y0 + h])      # lower triangle (v0,v1,v3)
        elif t % 4 == 1:
            xy = np.array([x0, y0, x0, y0 + h, x0 + h, y0 + h])      # upper triangle (v0,v2,v3)
        elif t % 4 == 2:
            xy = np.array([x0, y0, x0 + 2 * h, y0, x0 + 2 * h, y0 + h])  # stretched cell
        else:
            xy = rng.uniform(-3, 3, 6)                                # generic triangle
        d11, d22, d12 = rng.uniform(0.5, 2, 3), rng.uniform(0.5, 2, 3), rng.uniform(-0.3, 0.3, 3)
        if t < 8:
            d11[:], d22[:], d12[:] = 1.0, 1.0, 0.0                    # the shipped isotropic tensor
        D, dt = float(rng.choice([1200.0, 640.0, 3.7])), float(rng.choice([0.1, 0.05]))
        u0, f = rng.normal(size=3), float(rng.normal())
        r, s = float(rng.uniform(0, 200)), float(rng.uniform(0, 5))
        case = {"xy": xy.tolist(), "d11": d11.tolist(), "d22": d22.tolist(), "d12": d12.tolist(), "D": D, "dt": dt,
                "u0": u0.tolist(), "f": f, "r": r, "s": s}
        A = np.zeros(9); b = np.zeros(3)
        R.ref_hsld_cell_a(dp(A), dp(d11), dp(d22), dp(d12), cd(D), cd(dt), dp(xy)); case["cell_a"] = A.tolist()
        R.ref_hsld_cell_L(dp(b), dp(u0), cd(dt), cd(f), dp(xy)); case["cell_L"] = b.tolist()
        case["facet_a"], case["facet_L"], case["boundary"] = [], [], []
        for facet in range(3):
            R.ref_hsld_facet_a(dp(A), cd(dt), cd(r), dp(xy), C.c_int(facet), C.c_int(1 + facet % 2))
            case["facet_a"].append(A.tolist())
            R.ref_hsld_facet_L(dp(b), cd(dt), cd(r), cd(s), dp(xy), C.c_int(facet), C.c_int(1 + facet % 2))
            case["facet_L"].append(b.tolist())
            case["boundary"].append(R.ref_boundary_facet(dp(u0), dp(xy), C.c_int(facet)))
        xc = np.sort(rng.uniform(0, 10, 2)); v = float(rng.choice([120.0, 0.0, 33.0])); u2 = rng.normal(size=2)
        A4 = np.zeros(4); b2 = np.zeros(2)
        case.update({"xc": xc.tolist(), "v": v, "u2": u2.tolist()})
        R.ref_ad_cell_a(dp(A4), cd(dt), cd(D), cd(v), dp(xc)); case["ad_cell_a"] = A4.tolist()
        R.ref_ad_cell_L(dp(b2), dp(u2), cd(dt), cd(D), cd(v), dp(xc)); case["ad_cell_L"] = b2.tolist()
        case["ad_facet_a"], case["ad_facet_L"] = [], []
        for facet in range(2):
            R.ref_ad_facet_a(dp(A4), cd(dt), cd(r), C.c_int(facet), C.c_int(1 + facet)); case["ad_facet_a"].append(A4.tolist())
            R.ref_ad_facet_L(dp(b2), dp(u2), cd(dt), cd(r), cd(s), C.c_int(facet), C.c_int(1 + facet))
            case["ad_facet_L"].append(b2.tolist())
        cases.append(case)
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ufc_kernels.json")
    with open(out, "w") as fh:
        json.dump({"source": "FFC-generated kernels of /root/reference/fenics/{hslD,AdvectionDiffusion,boundary}.h "
                             "(FFC 2019.1.0.post0 / UFC 2018.1.0) compiled under oracle/shim",
                   "cases": cases}, fh)
    print("wrote", out, len(cases), "cases")


if __name__ == "__main__":
    main()
