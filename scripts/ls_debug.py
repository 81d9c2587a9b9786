"""Debug aid: per-step starting-guess diagnostics (EQGPU_LS_DEBUG) on the bench workload or a small mesh."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import eq_b200 as E
from oracle import oracle as O
n = int(sys.argv[1]); steps = int(sys.argv[2]); ncells = int(20000 * (n / 2048.0) ** 2)
g = E.GpuHSL(n, n)
W = (n - 1) * 0.5
cells = O.synthetic_colony(ncells, W, W, seed=12345)
g.upload_cells(cells, 2.0)
g.set_amounts(np.full(len(cells), 100.0))
its = []
for k in range(steps):
    g.gather_resident(); g.scatter_resident(); g.step()
    its.append(g.stats().iterations)
print("iterations", its, "mean", np.mean(its[10:]) if steps > 10 else np.mean(its), file=sys.stderr)
if n <= 400:
    p = O.Problem(nW=n, nH=n); s = O.new_state(p)
    for k in range(steps):
        s.u = O.scatter(cells, 2.0, n, n, np.full(len(cells), 100.0), s.u); s = O.step(p, s)
    print("rel err vs oracle", np.linalg.norm(g.get_field() - s.u) / np.linalg.norm(s.u), file=sys.stderr)
