#!/usr/bin/env python
"""Round-2 study, CPU only (development aid; imports the oracle): the BENCH metric under different starting-guess
policies -- mean MG-PCG iterations per step over steps 10..109 of the bench run (zero field, 100 nM per rod per step),
transient included, with the model solver of scripts/study_guess.py.

policies:  mode6     best of {zero, fixed extrapolations through the last 1..5 solutions} by residual norm (today)
           ext7      the same up to 7 solutions (what the image ring allows)
           corrX K   fixed extrapolation through the last K' = min(K, history) solutions plus a least-squares correction
                     in the backward-difference basis (K' x K' normal equations on the images, fitted to the
                     extrapolation's residual); the better of that and today's choice by residual norm

    python scripts/study_policy.py n policy [K]
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import study_guess as S  # noqa: E402

O = S.O


def backward_differences(vs, K):
    D = [list(vs[:K])]
    for j in range(1, K):
        D.append([D[-1][i] - D[-1][i + 1] for i in range(K - j)])
    return np.array([d[0] for d in D]).T


def main():
    n, policy = int(sys.argv[1]), sys.argv[2]
    K = int(sys.argv[3]) if len(sys.argv) > 3 else 5
    p = O.Problem(nW=n, nH=n)
    mg = S.MG(p, nu=int(os.environ.get("STUDY_NU", "3")))   # sweeps per smoothing pass on every level (GPU: 3 on level 0, 4 below)
    A, free = mg.levels[0]["A"], mg.levels[0]["free"]
    cells = O.synthetic_colony(int(20000 * (n / 2048.0) ** 2), p.W, p.H)
    u = np.zeros(p.N)
    hist, imgs, its, picks = [], [], [], []
    t0 = time.time()
    for k in range(110):
        # STUDY_SOURCE=osc: every rod's secretion oscillates (period 40 steps = 4 min, random phase per rod) the way a
        # synthetic gene oscillator would drive it, instead of the bench's constant 100 nM per step
        if os.environ.get("STUDY_SOURCE") == "move":
            # the colony expands: every rod drifts away from the trap centre by 0.02 um per step (a tenth of eQ's growth
            # speed scale; node spacing 0.5 um), so the set of nodes a rod deposits into changes every few steps
            if k == 0:
                ctr0 = cells[:, 11:13].copy()
                ang0, len0 = np.arctan2(cells[:, 15], cells[:, 14]), cells[:, 13].copy()
                mid = np.array([p.W / 2, p.H / 2])
                dirn = (ctr0 - mid) / np.maximum(np.linalg.norm(ctr0 - mid, axis=1, keepdims=True), 1e-9)
            cells = O.make_cells(ctr0 + 0.02 * k * dirn, ang0, len0, p.W, p.H)
            amount = np.full(len(cells), 100.0)
        elif os.environ.get("STUDY_SOURCE") == "osc":
            if k == 0:
                phase = np.random.default_rng(5).uniform(0, 2 * np.pi, len(cells))
            amount = 100.0 * (1.0 + 0.8 * np.sin(2 * np.pi * k / 40.0 + phase))
        else:
            amount = np.full(len(cells), 100.0)
        u0 = O.scatter(cells, 2.0, p.nH, p.nW, amount, u)
        _, b = O.assemble(p, u0, want_matrix=False)
        b = b * free
        depth = {"mode6": 5, "ext7": 7}.get(policy, 5 if policy == "corrX" and K < 5 else max(5, K))
        cands = {"zero": (np.zeros(p.N), b)}
        for kk in range(1, min(len(hist), depth) + 1):
            x = sum(c * hist[i] for i, c in enumerate(S.BINOM[kk]))
            cands[f"ext{kk}"] = (x, b - sum(c * imgs[i] for i, c in enumerate(S.BINOM[kk])))
        if policy == "corrX" and len(hist) >= 2:
            Kp = min(K, len(hist))
            W, AW = backward_differences(hist, Kp), backward_differences(imgs, Kp)
            x, r = cands[f"ext{Kp}"]
            sw = 1.0 / np.maximum(np.linalg.norm(AW, axis=0), 1e-300)
            c = np.linalg.solve((AW * sw).T @ (AW * sw) + 1e-13 * np.eye(Kp), (AW * sw).T @ r) * sw
            cands["corrX"] = (x + W @ c, r - AW @ c)
        best = min(cands, key=lambda name: np.linalg.norm(cands[name][1]))
        u, it = mg.pcg(b, cands[best][0] * free)
        its.append(it); picks.append(best)
        # image of the new solution: one more operator walk, or -- free -- b minus the residual vector the PCG recurrence
        # left behind (STUDY_RECURRENCE=1): they differ by the recurrence drift
        hist.insert(0, u.copy()); imgs.insert(0, (b - mg.last_r) * free if os.environ.get("STUDY_RECURRENCE") else A @ u)
        hist, imgs = hist[:8], imgs[:8]
    print(f"{n}x{n} source={os.environ.get('STUDY_SOURCE', 'const')} nu={os.environ.get('STUDY_NU', '3')} {policy} {K if policy == 'corrX' else ''}: mean iterations over steps 10..109 = {np.mean(its[10:]):.2f} "
          f"(first 10: {its[:10]}, every 10th after: {its[10::10]}, picks at 10/30/60/109: "
          f"{picks[10]}, {picks[30]}, {picks[60]}, {picks[109]}; {time.time() - t0:.0f} s)", flush=True)


if __name__ == "__main__":
    main()
