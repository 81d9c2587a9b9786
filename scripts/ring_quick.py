#!/usr/bin/env python
"""First run of warm-start mode 7 (image ring) on a GPU, sized for the last seconds of a GPU budget: (1) mode 7 against
mode 6 on a 321^2 bench-like run (same field? iterations? which guess?), (2) both modes on the 2048^2 bench workload,
10 + 100 steps, wall-clock around eqgpu_sync.  Prints one JSON line per part.  python scripts/ring_quick.py"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import eq_b200 as E  # noqa: E402

T0 = time.time()


def colony(name, n, w, seed):
    """cached beside the script (git-ignored; the cache travels with gpurun) so that a run does not spend its seconds here"""
    path = os.path.join(ROOT, "scripts", name)
    if os.path.exists(path):
        return np.load(path)
    from oracle import oracle as O
    c = O.synthetic_colony(n, w, w, seed=seed)
    np.save(path, c)
    return c


def run(nW, cells, steps, mode, warm=0):
    g = E.GpuHSL(nW, nW)
    g.set_warm_start(mode)
    g.upload_cells(cells, 2.0)
    g.set_amounts(np.full(len(cells), 100.0))
    its, guesses = [], []
    t = None
    for k in range(steps):
        if k == warm:
            g.sync()
            t = time.perf_counter()
        g.gather_resident()
        g.scatter_resident()
        g.step()
        its.append(int(g.stats().iterations))
        guesses.append(int(g.last_guess()))
    g.sync()
    dt = time.perf_counter() - t
    u = g.get_field()
    g.close()
    return u, its, guesses, (steps - warm) / dt


def main():
    out = {}
    try:
        cells = colony("_colony400.npy", 400, 160.0, 21)
        u6, i6, g6, _ = run(321, cells, 24, 6)
        u7, i7, g7, _ = run(321, cells, 24, 7)
        out["small"] = {"rel_diff_7_vs_6": float(np.linalg.norm(u7 - u6) / np.linalg.norm(u6)), "its6": i6, "its7": i7, "guess7": g7}
        print(json.dumps(out), flush=True)
        cells = colony("_colony20k.npy", 20000, 1023.5, 12345)
        for mode in (6, 7):
            u, its, gs, rate = run(2048, cells, 110, mode, warm=10)
            out[f"bench_mode{mode}"] = {"steps_per_s_wallclock": rate, "mean_iterations": float(np.mean(its[10:])),
                                        "its_every_10th": its[::10], "last_guess": gs[-1], "norm": float(np.linalg.norm(u))}
            print(json.dumps({f"bench_mode{mode}": out[f"bench_mode{mode}"], "t": time.time() - T0}), flush=True)
    except Exception as e:
        print(json.dumps({"error": repr(e), "partial": out, "t": time.time() - T0}), flush=True)


if __name__ == "__main__":
    main()
