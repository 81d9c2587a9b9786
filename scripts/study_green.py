#!/usr/bin/env python
"""Round-2 study, CPU only (development aid): can the warm start survive a MOVING colony if the response to the deposits is
taken out of the history?  u_{k+1} = T (u_k + s_k), T = A^-1 M.  T u_k is smooth in time (one step diffuses over
sqrt(tau) ~ 11 um); T s_k is not when rods move.  With g = T e (response to a unit point deposit, computed once,
translation-invariant away from the walls) the guess is   G s_k + [ring guess from the history of u_j - G s_{j-1}],
G s = s convolved with g.   python scripts/study_green.py [n] [policy: plain|green] [radius of g in nodes, 0 = full]"""
import os
import sys
import time

import numpy as np
import scipy.sparse.linalg as spla
from scipy.signal import fftconvolve

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import study_guess as S  # noqa: E402
from study_policy import backward_differences  # noqa: E402

O = S.O


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 641
    policy = sys.argv[2] if len(sys.argv) > 2 else "green"
    radius = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    p = O.Problem(nW=n, nH=n)
    mg = S.MG(p)
    A, free = mg.levels[0]["A"], mg.levels[0]["free"]
    # g = T e_centre
    e = np.zeros(p.N); c = n // 2; e[c * n + c] = 1.0
    _, be = O.assemble(p, e, want_matrix=False)
    g = spla.splu(A.tocsc()).solve(be * free).reshape(n, n)
    if radius:
        gm = np.zeros_like(g); gm[c - radius:c + radius + 1, c - radius:c + radius + 1] = g[c - radius:c + radius + 1, c - radius:c + radius + 1]; g = gm

    def G(s):
        return (fftconvolve(s.reshape(n, n), g, mode="full")[c:c + n, c:c + n]).ravel() * free

    cells = O.synthetic_colony(int(20000 * (n / 2048.0) ** 2), p.W, p.H, margin=40.0)
    ctr0 = cells[:, 11:13].copy(); ang0, len0 = np.arctan2(cells[:, 15], cells[:, 14]), cells[:, 13].copy()
    mid = np.array([p.W / 2, p.H / 2]); dirn = (ctr0 - mid) / np.maximum(np.linalg.norm(ctr0 - mid, axis=1, keepdims=True), 1e-9)
    u = np.zeros(p.N); hist, imgs, its = [], [], []
    t0 = time.time()
    for k in range(80):
        cells = O.make_cells(ctr0 + 0.02 * k * dirn, ang0, len0, p.W, p.H)
        s = O.scatter(cells, 2.0, p.nH, p.nW, np.full(len(cells), 100.0), np.zeros(p.N))
        _, b = O.assemble(p, u + s, want_matrix=False); b = b * free
        Gs = G(s) if policy == "green" else np.zeros(p.N)
        rb = b - A @ Gs                                   # what the history-based part has to explain
        Kp = min(7, len(hist))
        cands = {"zero": (Gs, rb)}
        if Kp >= 1:
            W_, AW = backward_differences(hist, Kp), backward_differences(imgs, Kp)
            x, r = W_.sum(axis=1), rb - AW.sum(axis=1)
            cands["ext"] = (Gs + x, r)
            if Kp >= 2:
                sw = 1.0 / np.maximum(np.linalg.norm(AW, axis=0), 1e-300)
                cc = np.linalg.solve((AW * sw).T @ (AW * sw) + 1e-13 * np.eye(Kp), (AW * sw).T @ r) * sw
                cands["corrX"] = (Gs + x + W_ @ cc, r - AW @ cc)
        best = min(cands, key=lambda q: np.linalg.norm(cands[q][1]))
        if k % 10 == 9:
            print(f"  step {k}: start residual / ||b|| = {np.linalg.norm(cands[best][1]) / np.linalg.norm(b):.2e} ({best})", flush=True)
        u, it = mg.pcg(b, cands[best][0] * free)
        its.append(it)
        ut = u - Gs                                       # history without the deposit response
        hist.insert(0, ut); imgs.insert(0, A @ ut); hist, imgs = hist[:8], imgs[:8]
    print(f"{n}x{n} moving colony, policy {policy}, radius {radius}: mean iterations over steps 10..79 = {np.mean(its[10:]):.2f} "
          f"(every 10th: {its[::10]}; {time.time() - t0:.0f} s)")


if __name__ == "__main__":
    main()
