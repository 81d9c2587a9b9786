#!/usr/bin/env python
"""Round-2 study, CPU only (imports the oracle: development aid, not product): how good a starting guess can the PCG
get from the stored history, and what does it buy in ITERATIONS of a multigrid-preconditioned CG like the GPU's?

Workload: the bench colony cut to n x n nodes (default 641), Dirichlet-0, 100 nM per rod per step, the run
itself is solved by the model PCG to rtol 1e-12 from the quartic extrapolation (so the stored history carries real PCG
errors); the exact solution (one SuperLU factorisation) is only used to measure the A-norm error of a guess.
Candidates at every step once the history is full:
  ext K   fixed extrapolation through the last K solutions (what modes 1-6 do, K = 1..5; the image ring would allow 6, 7)
  res K   residual-minimising combination of the last K solutions (mode 4 is res 3)
  gal K   Galerkin (energy-norm) projection onto the span of the last K solutions: min ||x* - sum c_i v_i||_A,
          i.e. (v_i, A v_j) c = (v_i, b) -- needs the same images A v_j the ring keeps, and K(K+3)/2 dot products
Reported per candidate, averaged over the last steps: relative starting residual, and iterations of a V(3,3)
Chebyshev-Jacobi MG-PCG (P1 interpolation on the "right" mesh, rediscretised coarse operators, the GPU's design) to
||r|| <= 1e-12 ||b||.

    python scripts/study_guess.py [n] [steps] [evaluated steps]
"""
import os
import sys
import time

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
import ctypes as C  # noqa: E402


def operator(p):
    """SPD operator on the free nodes (Dirichlet-0 rows/cols removed by zeroing) + mask"""
    bands, _ = O.assemble(p, None)
    mask, g = O.dirichlet(p)
    b = np.zeros(p.N)
    O.lib().eqo_apply_dirichlet_sym(C.c_long(p.nW), C.c_long(p.nH), O._dp(bands), O._dp(b), mask.ctypes.data_as(O.c_u8p), O._dp(g))
    return O.bands_to_csr(p, bands).tocsr(), mask != 0


def prolongation(nWc, nHc):
    """P1 interpolation coarse -> fine on the "right" mesh (fine = 2*coarse - 1 nodes per direction)"""
    nWf, nHf = 2 * nWc - 1, 2 * nHc - 1
    rows, cols, vals = [], [], []
    I, J = np.mgrid[0:nHf, 0:nWf]
    fi = (I * nWf + J).ravel()
    i2, j2 = (I // 2).ravel(), (J // 2).ravel()
    oi, oj = (I % 2).ravel(), (J % 2).ravel()
    c00 = i2 * nWc + j2

    def add(sel, c, w):
        rows.append(fi[sel]); cols.append(c[sel]); vals.append(np.full(sel.sum(), w))
    add((oi == 0) & (oj == 0), c00, 1.0)
    e = (oi == 0) & (oj == 1); add(e, c00, 0.5); add(e, c00 + 1, 0.5)                 # E-W edge midpoint
    e = (oi == 1) & (oj == 0); add(e, c00, 0.5); add(e, c00 + nWc, 0.5)               # N-S edge midpoint
    e = (oi == 1) & (oj == 1); add(e, c00, 0.5); add(e, c00 + nWc + 1, 0.5)           # SW-NE diagonal midpoint
    return sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(nWf * nHf, nWc * nHc))


class MG:
    def __init__(self, p, nu=3, coarsest=41):
        self.levels = []
        q = p
        while True:
            A, dmask = operator(q)
            self.levels.append({"A": A, "dinv": 1.0 / A.diagonal(), "free": ~dmask})
            if q.nW <= coarsest or q.nW % 2 == 0:
                break
            q = O.Problem(nW=(q.nW + 1) // 2, nH=(q.nH + 1) // 2, h=2 * q.h, dt=q.dt, D=q.D, bc_type=q.bc_type, bc_value=q.bc_value)
        for l in range(len(self.levels) - 1):
            nWc = int(round(np.sqrt(self.levels[l + 1]["A"].shape[0])))
            P = prolongation(nWc, nWc)
            P = sp.diags(self.levels[l]["free"].astype(float)) @ P @ sp.diags(self.levels[l + 1]["free"].astype(float))
            self.levels[l]["P"] = P.tocsr()
            self.levels[l]["R"] = P.T.tocsr()
        self.lu = spla.splu(self.levels[-1]["A"].tocsc())
        k = np.arange(nu)
        self.omega = 1.0 / (1.25 + 0.75 * np.cos(np.pi * (2 * k + 1) / (2 * nu)))    # Chebyshev roots on [0.5, 2]

    def smooth(self, l, x, b, reverse=False):
        L = self.levels[l]
        for w in (self.omega[::-1] if reverse else self.omega):
            x = x + w * L["dinv"] * (b - L["A"] @ x)
        return x

    def vcycle(self, l, b):
        if l == len(self.levels) - 1:
            return self.lu.solve(b)
        L = self.levels[l]
        x = self.smooth(l, np.zeros_like(b), b)
        x = x + L["P"] @ self.vcycle(l + 1, L["R"] @ (b - L["A"] @ x))
        return self.smooth(l, x, b, reverse=True)

    def pcg(self, b, x0, rtol=1e-12, maxit=30):
        A = self.levels[0]["A"]
        x = x0.copy()
        r = b - A @ x
        stop = rtol * np.linalg.norm(b)
        it = 0
        self.last_r = r              # the RECURRENCE residual at exit (what the GPU has in HBM when the step ends)
        if np.linalg.norm(r) <= stop:
            return x, 0
        z = self.vcycle(0, r)
        pv = z.copy()
        rz = r @ z
        while it < maxit:
            Ap = A @ pv
            a = rz / (pv @ Ap)
            x += a * pv
            r -= a * Ap
            it += 1
            if np.linalg.norm(r) <= stop:
                break
            z = self.vcycle(0, r)
            rz2 = r @ z
            pv = z + (rz2 / rz) * pv
            rz = rz2
        return x, it


BINOM = {1: [1], 2: [2, -1], 3: [3, -3, 1], 4: [4, -6, 4, -1], 5: [5, -10, 10, -5, 1], 6: [6, -15, 20, -15, 6, -1],
         7: [7, -21, 35, -35, 21, -7, 1]}


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 641
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 60
    n_eval = int(sys.argv[3]) if len(sys.argv) > 3 else 12
    big = n > 1100            # no exact solution (no LU at that size), only the candidates a GPU kernel could form
    p = O.Problem(nW=n, nH=n)
    t0 = time.time()
    mg = MG(p)
    A, free = mg.levels[0]["A"], mg.levels[0]["free"]
    lu = None if big else spla.splu(A.tocsc())
    cells = O.synthetic_colony(int(20000 * (n / 2048.0) ** 2), p.W, p.H)
    print(f"# {n}x{n}, {len(cells)} rods, {len(mg.levels)} levels, set-up {time.time() - t0:.1f} s", flush=True)
    rng = np.random.default_rng(0)
    u = np.zeros(p.N)
    hist, imgs = [], []          # newest first: solutions and their images A h
    stats = {}
    run_its = []
    for k in range(steps):
        amount = np.full(len(cells), 100.0)
        u0 = O.scatter(cells, 2.0, p.nH, p.nW, amount, u)
        _, b = O.assemble(p, u0, want_matrix=False)
        b = b * free
        exact = None if big else lu.solve(b)
        if k >= steps - n_eval and len(hist) >= 7:
            cands = {"zero": np.zeros(p.N)}
            for K in range(1, 8):
                cands[f"ext {K}"] = sum(c * hist[i] for i, c in enumerate(BINOM[K]))
            for K in (3, 5, 7):
                # difference basis: v0 = h0, v_i = h_{i-1} - h_i
                V = [hist[0]] + [hist[i - 1] - hist[i] for i in range(1, K)]
                AV = [imgs[0]] + [imgs[i - 1] - imgs[i] for i in range(1, K)]
                Vm, AVm = np.array(V).T, np.array(AV).T
                sc = 1.0 / np.linalg.norm(AVm, axis=0)
                c_res = np.linalg.lstsq(AVm * sc, b, rcond=1e-13)[0] * sc
                cands[f"res {K}"] = Vm @ c_res
                # the same fit through its K x K normal equations (what a GPU kernel can form in one pass: Gram matrix of
                # the images, column scaling, ridge 1e-13 -- k_ls_gram / ls_solve3 today for K = 3)
                N_ = (AVm * sc).T @ (AVm * sc)
                c_n = np.linalg.solve(N_ + 1e-13 * np.eye(K), (AVm * sc).T @ b) * sc
                cands[f"resN {K}"] = Vm @ c_n
                # backward-difference basis v_j = nabla^j h0 (the Newton form: ext K is the combination with all
                # coefficients 1): far less collinear than successive first differences, so the normal equations hold up
                Dh, Da = [list(hist[:K])], [list(imgs[:K])]
                for j in range(1, K):
                    Dh.append([Dh[-1][i] - Dh[-1][i + 1] for i in range(K - j)])
                    Da.append([Da[-1][i] - Da[-1][i + 1] for i in range(K - j)])
                Wm, AWm = np.array([d[0] for d in Dh]).T, np.array([d[0] for d in Da]).T
                sw = 1.0 / np.linalg.norm(AWm, axis=0)
                Nw = (AWm * sw).T @ (AWm * sw)
                cands[f"resD {K}"] = Wm @ (np.linalg.solve(Nw + 1e-13 * np.eye(K), (AWm * sw).T @ b) * sw)
                if K == 7 and k == steps - 1:
                    print(f"#   step {k}: cond of the scaled Gram matrix, first differences {np.linalg.cond(N_):.1e}, "
                          f"backward differences {np.linalg.cond(Nw):.1e}", flush=True)
                # the fit as a CORRECTION to a known good guess (k_ls_gram's "form 1" for K = 3): x = g + W c with the
                # normal equations fitted to the residual of g, so that no coefficient near 1 multiplies ||A h0|| ~ ||b||;
                # g = previous solution (corr) or the fixed extrapolation of the same depth (corrX)
                for tag, g0, ag0 in (("corr", hist[0], imgs[0]), ("corrX", Wm.sum(axis=1), AWm.sum(axis=1))):
                    rg0 = b - ag0
                    cands[f"{tag} {K}"] = g0 + Wm @ (np.linalg.solve(Nw + 1e-13 * np.eye(K), (AWm * sw).T @ rg0) * sw)
                # ... plus one step of iterative refinement: the true residual of the fitted guess is fitted again
                c1 = np.linalg.solve(Nw + 1e-13 * np.eye(K), (AWm * sw).T @ b) * sw
                r1 = b - AWm @ c1
                c2 = c1 + np.linalg.solve(Nw + 1e-13 * np.eye(K), (AWm * sw).T @ r1) * sw
                cands[f"resD+ {K}"] = Wm @ c2
                # Galerkin in the same basis, scaled to unit energy norm
                Gw = Wm.T @ AWm
                Gw = 0.5 * (Gw + Gw.T)
                sg = 1.0 / np.sqrt(np.abs(np.diag(Gw)))
                cg = np.linalg.solve(Gw * np.outer(sg, sg) + 1e-13 * np.eye(K), (Wm * sg).T @ b) * sg
                cands[f"galD {K}"] = Wm @ cg
                rg = b - AWm @ cg
                cands[f"galD+ {K}"] = Wm @ (cg + np.linalg.solve(Gw * np.outer(sg, sg) + 1e-13 * np.eye(K), (Wm * sg).T @ rg) * sg)
                G = (Vm * sc).T @ (AVm * sc)
                c_gal = np.linalg.lstsq(0.5 * (G + G.T), (Vm * sc).T @ b, rcond=1e-13)[0] * sc
                cands[f"gal {K}"] = Vm @ c_gal
            for name, x0 in cands.items():
                r0 = np.linalg.norm(b - A @ x0) / np.linalg.norm(b)
                if name.split()[0] in ("res", "resN", "gal", "galD", "galD+") and big:
                    continue
                eA = 0.0
                if not big:
                    e = exact - x0
                    eA = np.sqrt(max(e @ (A @ e), 0.0) / (exact @ (A @ exact)))
                _, it = mg.pcg(b, x0 * free)
                s = stats.setdefault(name, {"r0": [], "eA": [], "it": []})
                s["r0"].append(r0); s["eA"].append(eA); s["it"].append(it)
        # what the GPU would store: the PCG iterate that met the stopping test, started from the best fixed extrapolation
        # the history allows (its error has the structure of a PCG error -- white noise of that size in u would have a
        # residual kappa times larger and hide everything below 1e-9)
        Kh = min(len(hist), 5)
        x0 = sum(c * hist[i] for i, c in enumerate(BINOM[Kh])) if Kh else np.zeros(p.N)
        u, its_run = mg.pcg(b, x0 * free)
        run_its.append(its_run)
        hist.insert(0, u.copy()); imgs.insert(0, A @ u)
        hist, imgs = hist[:8], imgs[:8]
    print(f"# the run itself (ext 5 once the history is full): iterations per step, last 12: {run_its[-12:]}")
    print(f"{'guess':8s} {'start residual':>15s} {'A-norm error':>13s} {'iterations':>11s}")
    for name, s in stats.items():
        print(f"{name:9s} {np.exp(np.mean(np.log(s['r0']))):15.2e} {np.exp(np.mean(np.log(np.maximum(s['eA'], 1e-300)))):13.2e} {np.mean(s['it']):11.2f}")


if __name__ == "__main__":
    main()
