"""Synthetic rod colonies for the bench and the tests (host side, numpy; no device code, no oracle).

The records are the 16-double cell records of include/eqgpu.h, i.e. what eq_b200/host/cellRecords.h reads from
the reference's Ecoli / cpmEcoli objects:
  [0,1] bodyA position   [2,3] bodyA rotation (cos, sin)   [4] offset = (L0 - W)/2 (src/abm/cpmEcoli.cpp:121)
  [5] vertsA[1].x = ratcheted newOffset (:407-415)   [6] radius = W/2 (:122)   [7..10] polePositionA, polePositionB
  (src/abm/Ecoli.cpp:36-63, clamped to the trap)   [11,12] centre   [13] length   [14,15] cos, sin of the cell angle.

`Colony.advance()` moves, turns and grows the rods the way a Chipmunk step + cpmEcoli::updateModel would leave
them (two body halves drifting apart as the cell grows exponentially, src/abm/Ecoli.cpp:71; the back-filled
rectangle of bodyA following in RATCHET_QUANTUM steps once the gap exceeds COMPRESSION_GAP, src/abm/cpmEcoli.cpp:
14-15,407-422; division at the division length, one daughter kept so that the rod count stays fixed).  It is a
workload generator, not a mechanics engine: rods do not collide.
"""
from __future__ import annotations

import math

import numpy as np

CELL_STRIDE = 16
RATCHET_QUANTUM = 0.05      # src/abm/cpmEcoli.cpp:14
COMPRESSION_GAP = 0.1       # src/abm/cpmEcoli.cpp:15
DIVISION_LENGTH = 4.2       # src/eQcell.h:43
DOUBLING_MINUTES = 20.0     # src/eQcell.h (default doubling period)


def _cos_sin(a):
    # the host's libm, element by element: the record carries exactly what C's cos()/sin() return
    a = np.asarray(a, dtype=np.float64)
    return (np.array([math.cos(v) for v in a.ravel()]).reshape(a.shape),
            np.array([math.sin(v) for v in a.ravel()]).reshape(a.shape))


def poles(cx, cy, ca, sa, length, width, trapW, trapH):
    """Ecoli::updatePoleCenters (src/abm/Ecoli.cpp:36-63), including its asymmetric second pole."""
    out = []
    for r in (1.0, -1.0):
        s = r * 0.5 * length - width / 2.0
        out.append(np.clip(cx + ca * s, 0.0, trapW))
        out.append(np.clip(cy + sa * s, 0.0, trapH))
    return out


def make_cells(centers, angles, lengths, trapW, trapH, width=1.0):
    """Fresh (un-ratcheted) rods: both body halves at the centre, newOffset == offset."""
    c = np.asarray(centers, dtype=np.float64).reshape(-1, 2)
    a = np.asarray(angles, dtype=np.float64)
    L = np.asarray(lengths, dtype=np.float64)
    ca, sa = _cos_sin(a)
    rec = np.zeros((len(a), CELL_STRIDE))
    rec[:, 0], rec[:, 1], rec[:, 2], rec[:, 3] = c[:, 0], c[:, 1], ca, sa
    rec[:, 4] = (L - width) * 0.5
    rec[:, 5] = rec[:, 4]
    rec[:, 6] = width * 0.5
    rec[:, 7], rec[:, 8], rec[:, 9], rec[:, 10] = poles(c[:, 0], c[:, 1], ca, sa, L, width, trapW, trapH)
    rec[:, 11], rec[:, 12], rec[:, 13] = c[:, 0], c[:, 1], L
    rec[:, 14], rec[:, 15] = ca, sa
    return rec


def synthetic_layout(n, trapW, trapH, seed=12345, min_clear=1.2, margin=3.0):
    """centres uniform in [margin, W-margin] x [margin, H-margin], angle U[0, 2 pi),
    length (1+U) * 0.5 * 4.2 (src/abm/eQabm.cpp:115), rejection-sampled to pairwise separated rods.
    Returns (centers[n,2], angles[n], lengths[n])."""
    rng = np.random.default_rng(seed)
    cell = 6.0   # hash-grid pitch > longest rod + clearance
    gx = int(np.ceil(trapW / cell)) + 1
    grid: dict = {}
    centers, angles, lengths = [], [], []
    tries = 0
    while len(angles) < n and tries < 200 * n:
        tries += 1
        x = rng.uniform(margin, trapW - margin)
        y = rng.uniform(margin, trapH - margin)
        a = rng.uniform(0.0, 2 * np.pi)
        L = (1.0 + rng.uniform()) * 0.5 * DIVISION_LENGTH
        ix, iy = int(x / cell), int(y / cell)
        ok = True
        for dx in (-1, 0, 1):
            for dy in (-1, 0, 1):
                for (ox, oy, oL) in grid.get((ix + dx) + gx * (iy + dy), ()):
                    if (ox - x) ** 2 + (oy - y) ** 2 < (0.5 * (L + oL) + 1.0 + min_clear) ** 2:
                        ok = False
                        break
                if not ok:
                    break
            if not ok:
                break
        if not ok:
            continue
        grid.setdefault(ix + gx * iy, []).append((x, y, L))
        centers.append((x, y))
        angles.append(a)
        lengths.append(L)
    return np.array(centers).reshape(-1, 2), np.array(angles), np.array(lengths)


def synthetic_colony(n, trapW, trapH, seed=12345, min_clear=1.2, margin=3.0):
    c, a, L = synthetic_layout(n, trapW, trapH, seed, min_clear, margin)
    return make_cells(c, a, L, trapW, trapH)


class Colony:
    """A colony that changes: mode 'static' (records never change), 'moving' (rods translate along their axis
    by `speed` nodes per step, turn slowly, bounce off the trap margin) or 'growing' (moving + exponential growth
    with the ratchet and division)."""

    def __init__(self, n, trapW, trapH, npm=2.0, mode="moving", seed=12345, dt=0.1, width=1.0,
                 speed_nodes=(0.05, 0.2), margin=3.0):
        assert mode in ("static", "moving", "growing")
        self.mode, self.W, self.H, self.npm, self.dt, self.width, self.margin = mode, trapW, trapH, npm, dt, width, margin
        self.c, self.angle, self.L = synthetic_layout(n, trapW, trapH, seed, margin=margin)
        self.n = len(self.angle)
        rng = np.random.default_rng(seed + 1)
        lo, hi = speed_nodes
        self.v = rng.uniform(lo, hi, self.n) / npm * rng.choice([-1.0, 1.0], self.n)   # um per step along the axis
        self.omega = rng.uniform(-2e-3, 2e-3, self.n)                                    # rad per step
        self.L0 = self.L.copy()                 # birth length: the two halves coincide at birth
        self.next_ratchet = np.full(self.n, RATCHET_QUANTUM)    # src/abm/cpmEcoli.cpp:124
        self.new_offset = (self.L0 - width) * 0.5
        self.growth = math.log(2.0) / DOUBLING_MINUTES * dt     # src/abm/Ecoli.cpp:71
        self.steps = 0

    def advance(self):
        if self.mode == "static":
            return
        ca, sa = np.cos(self.angle), np.sin(self.angle)
        self.c[:, 0] += self.v * ca
        self.c[:, 1] += self.v * sa
        # bounce: a rod that would leave the margin turns round
        out = ((self.c[:, 0] < self.margin) | (self.c[:, 0] > self.W - self.margin) |
               (self.c[:, 1] < self.margin) | (self.c[:, 1] > self.H - self.margin))
        self.v[out] = -self.v[out]
        self.c[:, 0] = np.clip(self.c[:, 0], self.margin, self.W - self.margin)
        self.c[:, 1] = np.clip(self.c[:, 1], self.margin, self.H - self.margin)
        self.angle += self.omega
        if self.mode == "growing":
            self.L *= 1.0 + self.growth
            sep = self.L - self.L0
            # one ratchet notch per updateModel call (src/abm/cpmEcoli.cpp:407-411)
            hit = sep > self.next_ratchet + COMPRESSION_GAP
            self.new_offset[hit] = (self.L0[hit] - self.width) * 0.5 + self.next_ratchet[hit]
            self.next_ratchet[hit] += RATCHET_QUANTUM
            div = self.L >= DIVISION_LENGTH
            if div.any():   # keep the daughter on the -axis side; it is born un-ratcheted
                d = np.stack([np.cos(self.angle[div]), np.sin(self.angle[div])], axis=1)
                self.c[div] -= d * (0.25 * self.L[div])[:, None]
                self.L[div] *= 0.5
                self.L0[div] = self.L[div]
                self.next_ratchet[div] = RATCHET_QUANTUM
                self.new_offset[div] = (self.L0[div] - self.width) * 0.5
        self.steps += 1

    def records(self):
        """The records cellRecords.h would build from this state: bodyA sits half the separation behind the
        centre and its back-filled rectangle reaches (-offset, newOffset) in its own frame."""
        ca, sa = _cos_sin(self.angle)
        sep = self.L - self.L0
        rec = np.zeros((self.n, CELL_STRIDE))
        rec[:, 0] = self.c[:, 0] - ca * (0.5 * sep)
        rec[:, 1] = self.c[:, 1] - sa * (0.5 * sep)
        rec[:, 2], rec[:, 3] = ca, sa
        rec[:, 4] = (self.L0 - self.width) * 0.5
        rec[:, 5] = self.new_offset
        rec[:, 6] = self.width * 0.5
        rec[:, 7], rec[:, 8], rec[:, 9], rec[:, 10] = poles(self.c[:, 0], self.c[:, 1], ca, sa, self.L, self.width,
                                                            self.W, self.H)
        rec[:, 11], rec[:, 12], rec[:, 13] = self.c[:, 0], self.c[:, 1], self.L
        rec[:, 14], rec[:, 15] = ca, sa
        return rec
