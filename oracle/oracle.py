"""CPU oracle for eQ's HSL diffusion hot path -- numpy/scipy front end.

TEST INFRASTRUCTURE ONLY.  Imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / ``--impl reference`` legs; never by eq_b200/.

It drives oracle/eq_oracle.c (the C restatement; every function there cites the
reference file:line it follows) and adds the pieces that live in third-party
code upstream: the sparse direct solve behind ``LVS->solve()``
(src/fHSL.cpp:106; DOLFIN 2019.1.0 default "lu" [ext]) is SciPy's SuperLU.

Parity pinning: see the header of eq_oracle.c.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_REF = None

c_dp = C.POINTER(C.c_double)
c_lp = C.POINTER(C.c_long)
c_u8p = C.POINTER(C.c_uint8)

CELL_STRIDE = 16
NBAND = 7


def build(quiet: bool = True) -> None:
    """Compile libeq_oracle.so (and oracle/_ref when /root/reference exists)."""
    subprocess.run(["make", "-C", _HERE, "all"], check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


def _dp(a):
    return a.ctypes.data_as(c_dp) if a is not None else None


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libeq_oracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        L.eqo_cg.restype = C.c_long
        L.eqo_raster_cell.restype = C.c_long
        L.eqo_boundary_functional.restype = C.c_double
        L.eqo_boundary_facet.restype = C.c_double
        L.eqo_deposit_per_point.restype = C.c_double
        L.eqo_point_in_cell.restype = C.c_int
        L.eqo_num_threads.restype = C.c_int
        L.eqo_set_num_threads.argtypes = [C.c_int]
        L.eqo_set_num_threads.restype = None
        L.eqo_assemble.restype = C.c_int
        L.eqo_fd_bicgstab.restype = C.c_long
        _LIB = L
    return _LIB


def ref_lib():
    """The reference's own generated kernels (oracle/_ref), or None."""
    global _REF
    if _REF is None:
        path = os.path.join(_HERE, "_ref", "libeq_ufc_ref.so")
        if not os.path.exists(path):
            return None
        R = C.CDLL(path)
        R.ref_boundary_facet.restype = C.c_double
        _REF = R
    return _REF


# --------------------------------------------------------------------------
# problem description (mirrors the parameters fenicsInterface reads:
# src/eQ.h:305-321 and the eQ::data::parameters keys of SURVEY.md 8b)
# --------------------------------------------------------------------------
NEUMANN, DIRICHLET, ROBIN, DIRICHLET_CHANNEL = 0, 1, 2, 3
LEFT, RIGHT, TOP, BOTTOM = 0, 1, 2, 3


@dataclass
class Problem:
    nW: int
    nH: int
    h: float = 0.5
    hy: float | None = None   # row spacing when it differs from h (W/ceil(W*npm) != H/ceil(H*npm) upstream)
    dt: float = 0.1
    D: float = 1200.0
    # per wall (left,right,top,bottom): type and value (Dirichlet value / Robin rate)
    bc_type: tuple = (DIRICHLET, DIRICHLET, DIRICHLET, DIRICHLET)
    bc_value: tuple = (0.0, 0.0, 0.0, 0.0)
    robin_s: tuple = (0.0, 0.0)
    # channels (src/fHSL.cpp:110-152)
    channels: bool = False
    channel_v: float = 120.0
    channel_r: tuple = (0.0, 0.0)      # Robin rates at channel ends (left,right)
    channel_iters: int = 48
    well_scaling: float = 25.0
    d11: np.ndarray | None = None
    d22: np.ndarray | None = None
    d12: np.ndarray | None = None

    @property
    def N(self):
        return self.nW * self.nH

    @property
    def W(self):
        return (self.nW - 1) * self.h

    @property
    def H(self):
        return (self.nH - 1) * (self.hy if self.hy else self.h)

    @property
    def use_robin(self):
        return self.bc_type[LEFT] == ROBIN or self.bc_type[RIGHT] == ROBIN


def assemble(p: Problem, u0: np.ndarray | None, want_matrix=True):
    """Unconstrained (bands, b) as DOLFIN would assemble hslD's a and L."""
    L = lib()
    N = p.N
    bands = np.zeros((NBAND, N)) if want_matrix else None
    b = np.zeros(N)
    r1 = p.bc_value[LEFT] if p.bc_type[LEFT] == ROBIN else 0.0
    r2 = p.bc_value[RIGHT] if p.bc_type[RIGHT] == ROBIN else 0.0
    rc = L.eqo_assemble(C.c_long(p.nW), C.c_long(p.nH), C.c_double(p.W), C.c_double(p.H),
                        C.c_double(p.D), C.c_double(p.dt), _dp(p.d11), _dp(p.d22), _dp(p.d12),
                        C.c_int(1 if p.use_robin else 0), C.c_double(r1), C.c_double(p.robin_s[0]),
                        C.c_double(r2), C.c_double(p.robin_s[1]),
                        _dp(u0), C.c_double(0.0), _dp(bands), _dp(b))
    assert rc == 0
    return bands, b


def dirichlet(p: Problem, top_vals=None, bottom_vals=None):
    L = lib()
    is_dir = (C.c_int * 4)(*[1 if t in (DIRICHLET, DIRICHLET_CHANNEL) else 0 for t in p.bc_type])
    val = (C.c_double * 4)(*[float(v) for v in p.bc_value])
    mask = np.zeros(p.N, dtype=np.uint8)
    g = np.zeros(p.N)
    tv = np.ascontiguousarray(top_vals, dtype=np.float64) if top_vals is not None else None
    bv = np.ascontiguousarray(bottom_vals, dtype=np.float64) if bottom_vals is not None else None
    L.eqo_dirichlet_mask(C.c_long(p.nW), C.c_long(p.nH), is_dir, val, _dp(tv), _dp(bv),
                         mask.ctypes.data_as(c_u8p), _dp(g))
    return mask, g


def bands_to_csr(p: Problem, bands):
    import scipy.sparse as sp
    off = (C.c_long * NBAND)()
    lib().eqo_band_offsets(C.c_long(p.nW), off)
    N = p.N
    diags = []
    offs = []
    for k in range(NBAND):
        o = off[k]
        d = bands[k]
        # scipy dia: data[k, j] = A[j - o, j]; we hold row-indexed A[i, i+o]
        arr = np.zeros(N)
        if o >= 0:
            arr[o:] = d[:N - o]
        else:
            arr[:N + o] = d[-o:]
        diags.append(arr)
        offs.append(o)
    return sp.dia_matrix((np.array(diags), offs), shape=(N, N)).tocsc()


def solve_lu(p: Problem, u0, top_vals=None, bottom_vals=None, symmetric=False):
    """One backward-Euler solve the way src/fHSL.cpp:104-108 does it:
    assemble, apply DirichletBC (identity rows), sparse direct LU."""
    import scipy.sparse.linalg as spla
    bands, b = assemble(p, u0)
    mask, g = dirichlet(p, top_vals, bottom_vals)
    L = lib()
    if mask.any():
        if symmetric:
            L.eqo_apply_dirichlet_sym(C.c_long(p.nW), C.c_long(p.nH), _dp(bands), _dp(b),
                                      mask.ctypes.data_as(c_u8p), _dp(g))
        else:
            L.eqo_apply_dirichlet_rows(C.c_long(p.N), _dp(bands), _dp(b),
                                       mask.ctypes.data_as(c_u8p), _dp(g))
    A = bands_to_csr(p, bands)
    lu = spla.splu(A)
    return lu.solve(b)


def solve_cg(p: Problem, u0, top_vals=None, bottom_vals=None, rtol=1e-13, maxit=200000, x0=None):
    """Same system, symmetric elimination + Jacobi-CG (for meshes too large for LU)."""
    bands, b = assemble(p, u0)
    mask, g = dirichlet(p, top_vals, bottom_vals)
    L = lib()
    L.eqo_apply_dirichlet_sym(C.c_long(p.nW), C.c_long(p.nH), _dp(bands), _dp(b),
                              mask.ctypes.data_as(c_u8p), _dp(g))
    x = np.array(u0 if x0 is None else x0, dtype=np.float64, copy=True)
    x[mask != 0] = g[mask != 0]
    rel = C.c_double(0.0)
    it = L.eqo_cg(C.c_long(p.nW), C.c_long(p.nH), _dp(bands), _dp(b), _dp(x),
                  C.c_double(rtol), C.c_long(maxit), C.byref(rel))
    return x, it, rel.value


def band_matvec(p: Problem, bands, x):
    y = np.zeros(p.N)
    lib().eqo_band_matvec(C.c_long(p.nW), C.c_long(p.nH), _dp(bands), _dp(np.ascontiguousarray(x)), _dp(y))
    return y


def boundary_functional(p: Problem, u):
    return lib().eqo_boundary_functional(C.c_long(p.nW), C.c_long(p.nH), C.c_double(p.W),
                                         C.c_double(p.H), _dp(np.ascontiguousarray(u)))


def compute_boundary_flux(p: Problem, u):
    fb = np.zeros(p.nW)
    ft = np.zeros(p.nW)
    lib().eqo_compute_boundary_flux(C.c_long(p.nW), C.c_long(p.nH), _dp(np.ascontiguousarray(u)),
                                    C.c_double(p.h), C.c_double(p.dt), C.c_double(p.D),
                                    C.c_double(p.well_scaling), _dp(fb), _dp(ft))
    return fb, ft


def channel_substeps(p: Problem, flux, u):
    u = np.array(u, dtype=np.float64, copy=True)
    lib().eqo_channel_substeps(C.c_long(p.nW), C.c_double(p.W), C.c_double(p.dt),
                               C.c_long(p.channel_iters), C.c_double(p.D), C.c_double(p.channel_v),
                               C.c_double(p.channel_r[0]), C.c_double(0.0),
                               C.c_double(p.channel_r[1]), C.c_double(0.0),
                               _dp(np.ascontiguousarray(flux)), _dp(u))
    return u


def robin_rates(v, D, L_left, L_right):
    a = C.c_double()
    b = C.c_double()
    lib().eqo_robin_rates(C.c_double(v), C.c_double(D), C.c_double(L_left), C.c_double(L_right),
                          C.byref(a), C.byref(b))
    return a.value, b.value


@dataclass
class State:
    """Mutable solver state of one layer (what fenicsInterface holds)."""
    u: np.ndarray
    top: np.ndarray
    bottom: np.ndarray
    total_boundary_flux: float = 0.0
    info: dict = field(default_factory=dict)


def new_state(p: Problem) -> State:
    return State(u=np.zeros(p.N), top=np.zeros(p.nW), bottom=np.zeros(p.nW))


def step(p: Problem, s: State, solver="lu") -> State:
    """fenicsInterface::stepDiffusion, src/fHSL.cpp:98-161.  s.u holds the
    previous solution plus the cell deposits (what the controller sent)."""
    top_vals = s.top if p.bc_type[TOP] == DIRICHLET_CHANNEL else None
    bot_vals = s.bottom if p.bc_type[BOTTOM] == DIRICHLET_CHANNEL else None
    if solver == "lu":
        u = solve_lu(p, s.u, top_vals, bot_vals)
    else:
        u, it, rel = solve_cg(p, s.u, top_vals, bot_vals)
        s.info["cg_iters"] = it
        s.info["cg_relres"] = rel
    s.u = u
    if p.channels:
        fb, ft = compute_boundary_flux(p, u)
        s.top = channel_substeps(p, ft, s.top)
        s.bottom = channel_substeps(p, fb, s.bottom)
    s.total_boundary_flux = p.D * p.dt * boundary_functional(p, u)
    return s


# --------------------------------------------------------------------------
# cells
# --------------------------------------------------------------------------
def make_cells(centers, angles, lengths, trapW, trapH, width=1.0):
    n = len(angles)
    rec = np.zeros((n, CELL_STRIDE))
    L = lib()
    for k in range(n):
        L.eqo_make_cell(C.c_double(centers[k][0]), C.c_double(centers[k][1]), C.c_double(angles[k]),
                        C.c_double(lengths[k]), C.c_double(width), C.c_double(trapW), C.c_double(trapH),
                        rec[k].ctypes.data_as(c_dp))
    return rec


def nodes_to_edge(npm):
    # src/abm/eQabm.cpp:75
    import math
    return int(math.floor(npm * 1.0 / 2.0 + 0.5))  # C round(): half away from zero


def raster(cells, npm, nH, nW, cap=512):
    n = cells.shape[0]
    counts = np.zeros(n, dtype=np.int64)
    nodes = np.full((n, cap), -1, dtype=np.int64)
    lib().eqo_raster(_dp(cells), C.c_long(n), C.c_double(npm), C.c_long(nH), C.c_long(nW),
                     C.c_long(nodes_to_edge(npm)), counts.ctypes.data_as(c_lp),
                     nodes.ctypes.data_as(c_lp), C.c_long(cap))
    return counts, nodes


def gather(cells, npm, nH, nW, u):
    out = np.zeros(cells.shape[0])
    lib().eqo_gather(_dp(cells), C.c_long(cells.shape[0]), C.c_double(npm), C.c_long(nH), C.c_long(nW),
                     C.c_long(nodes_to_edge(npm)), _dp(np.ascontiguousarray(u)), _dp(out))
    return out


def scatter(cells, npm, nH, nW, amount_nM, u):
    u = np.array(u, dtype=np.float64, copy=True)
    lib().eqo_scatter(_dp(cells), C.c_long(cells.shape[0]), C.c_double(npm), C.c_long(nH), C.c_long(nW),
                      C.c_long(nodes_to_edge(npm)), _dp(np.ascontiguousarray(amount_nM)), _dp(u))
    return u


def cells_tensor(cells, npm, nH, nW, Dx, Dy):
    """eQabm::updateCells' D11/D22/D12 grids (src/abm/eQabm.cpp:246-248,306-325,407)."""
    d11, d22, d12 = np.empty(nH * nW), np.empty(nH * nW), np.empty(nH * nW)
    lib().eqo_cells_tensor(_dp(cells), C.c_long(cells.shape[0]), C.c_double(npm), C.c_long(nH), C.c_long(nW),
                           C.c_long(nodes_to_edge(npm)), C.c_double(Dx), C.c_double(Dy), _dp(d11), _dp(d22), _dp(d12))
    return d11, d22, d12


def update_cells_sequential(cells, npm, nH, nW, a0, a1, u):
    u = np.array(u, dtype=np.float64, copy=True)
    g = np.zeros(cells.shape[0])
    lib().eqo_update_cells_sequential(_dp(cells), C.c_long(cells.shape[0]), C.c_double(npm),
                                      C.c_long(nH), C.c_long(nW), C.c_long(nodes_to_edge(npm)),
                                      _dp(np.ascontiguousarray(a0)), C.c_double(a1), _dp(u), _dp(g))
    return u, g


def synthetic_colony(n, trapW, trapH, seed=12345, min_clear=1.2, margin=3.0):
    """Config-3 style colony: centres uniform in [margin, W-margin] x
    [margin, H-margin], angle U[0,2pi), length (1+U)*0.5*4.2 (src/abm/eQabm.cpp:115),
    rejection-sampled so rods are pairwise separated (scatter order-independent)."""
    rng = np.random.default_rng(seed)
    cell = 6.0  # hash-grid pitch > max rod length + clearance
    gx = int(np.ceil(trapW / cell)) + 1
    grid: dict = {}
    centers, angles, lengths = [], [], []
    tries = 0
    while len(angles) < n and tries < 200 * n:
        tries += 1
        x = rng.uniform(margin, trapW - margin)
        y = rng.uniform(margin, trapH - margin)
        a = rng.uniform(0.0, 2 * np.pi)
        L = (1.0 + rng.uniform()) * 0.5 * 4.2
        ix, iy = int(x / cell), int(y / cell)
        ok = True
        for dx in (-1, 0, 1):
            for dy in (-1, 0, 1):
                for (ox, oy, oL) in grid.get((ix + dx) + gx * (iy + dy), ()):
                    # conservative: treat rods as discs of radius L/2 + 0.5
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
    return make_cells(centers, angles, lengths, trapW, trapH)


# --------------------------------------------------------------------------
# diffusionPETSc (diffuclass.cpp): the reference's second eQ::diffusionSolver,
# 5-point finite differences with ghost-node Neumann/Robin rows.  PETSc's
# matrix-free FBCGSR solve at its default rtol 1e-5 [ext] is replaced here by a
# sparse direct solve of the same matrix (scipy SuperLU), so the oracle is the
# exact solution of the system the reference iterates on.
# --------------------------------------------------------------------------
@dataclass
class FDWalls:
    """DiffusionData's boundary coefficients (diffuclass.h:16-32), wall order left,right,top,bottom:
    Dc*u + Nc*du/dn = BV.  Nc == 0 is a Dirichlet wall ("if N==0, then D must be 1", diffuclass.cpp:232)."""
    Dc: tuple = (1.0, 1.0, 1.0, 1.0)
    Nc: tuple = (0.0, 0.0, 0.0, 0.0)
    BV: tuple = (0.0, 0.0, 0.0, 0.0)


def fd_walls_from_problem(p: Problem) -> FDWalls:
    """The FD coefficients that describe the same walls as a Problem's (type, value, s): Dirichlet value v ->
    (Dc 1, Nc 0, BV v); Neumann -> (0, 1, 0); Robin rate r, s -> Nc 1, Dc = r/D, BV = Dc*s."""
    Dc, Nc, BV = [], [], []
    for w in range(4):
        t = p.bc_type[w]
        if t == DIRICHLET:
            Dc.append(1.0); Nc.append(0.0); BV.append(float(p.bc_value[w]))
        elif t == ROBIN:
            d = float(p.bc_value[w]) / p.D
            Dc.append(d); Nc.append(1.0); BV.append(d * float(p.robin_s[w]))
        elif t == NEUMANN:
            Dc.append(0.0); Nc.append(1.0); BV.append(0.0)
        else:
            raise ValueError("diffusionPETSc has no channel-coupled walls")
    return FDWalls(tuple(Dc), tuple(Nc), tuple(BV))


def fd_assemble(p: Problem, walls: FDWalls):
    """The matrix MyMatMult applies (diffuclass.cpp:786-862, the active loop), natural order i + j*gridNodesX.
    One deviation: at a corner between a Neumann/Robin top/bottom wall and a Dirichlet side wall the
    reference falls into its generic top/bottom-wall row and reads xarray[j][i-1] outside the grid, while
    ApplyBoundaryConditions writes the side wall's Dirichlet value there (:251-271); the row is the identity
    here, i.e. the Dirichlet wall wins."""
    import scipy.sparse as sp
    nX, nY = p.nW, p.nH
    h = p.h
    assert p.hy is None or p.hy == p.h, "diffusionPETSc has one grid spacing h"
    F = (p.D * p.dt) / (h * h)                 # diffuclass.cpp:361
    lD, rD, tD, bD = walls.Dc
    lN, rN, tN, bN = walls.Nc
    top = nY - 1
    idx = lambda i, j: i + j * nX
    rows, cols, vals = [], [], []
    irows, icols, ivals = [], [], []

    def put(r, entries):
        for (c, v) in entries:
            rows.append(r); cols.append(c); vals.append(v)

    # interior: -F on the four neighbours, 1 + 4F on the diagonal (:857-860)
    I, J = np.meshgrid(np.arange(1, nX - 1), np.arange(1, nY - 1))
    g = (I + J * nX).ravel()
    for off, v in ((0, 1.0 + 4.0 * F), (1, -F), (-1, -F), (nX, -F), (-nX, -F)):
        irows.append(g); icols.append(g + off); ivals.append(np.full(g.size, v))
    for j in range(nY):
        for i in range(nX):
            if not (i == 0 or j == 0 or i == nX - 1 or j == top):
                continue
            r = idx(i, j)
            if (tN != 0 and j == top) or (bN != 0 and j == 0 and not (tN != 0 and j == top)):
                is_top = (tN != 0 and j == top)
                wD, wN = (tD, tN) if is_top else (bD, bN)
                jin = j - 1 if is_top else j + 1
                if i == 0 and lN != 0:
                    put(r, [(idx(i + 1, j), -2 * F), (idx(i, jin), -2 * F),
                            (r, 1 + (4 + (2 * h * wD / wN) + (2 * h * lD / lN)) * F)])
                elif i == nX - 1 and rN != 0:
                    put(r, [(idx(i - 1, j), -2 * F), (idx(i, jin), -2 * F),
                            (r, 1 + (4 + (2 * h * wD / wN) + (2 * h * rD / rN)) * F)])
                elif i == 0 or i == nX - 1:
                    put(r, [(r, 1.0)])          # Dirichlet side wall wins the corner (see docstring)
                else:
                    put(r, [(idx(i, jin), -2 * F), (idx(i - 1, j), -F), (idx(i + 1, j), -F),
                            (r, 1 + (4 + 2 * h * wD / wN) * F)])
            elif lN != 0 and i == 0 and j != 0 and j != top:
                put(r, [(idx(i, j - 1), -F), (idx(i, j + 1), -F), (idx(i + 1, j), -2 * F),
                        (r, 1 + (4 + (2 * h * lD / lN)) * F)])
            elif rN != 0 and i == nX - 1 and j != 0 and j != top:
                put(r, [(idx(i, j - 1), -F), (idx(i, j + 1), -F), (idx(i - 1, j), -2 * F),
                        (r, 1 + (4 + (2 * h * rD / rN)) * F)])
            else:
                put(r, [(r, 1.0)])              # Dirichlet: yarray = xarray (:850-852)
    rows = np.concatenate(irows + [np.asarray(rows, dtype=np.int64)])
    cols = np.concatenate(icols + [np.asarray(cols, dtype=np.int64)])
    vals = np.concatenate(ivals + [np.asarray(vals, dtype=np.float64)])
    return sp.csc_matrix((vals, (rows, cols)), shape=(p.N, p.N))


def fd_rhs(p: Problem, walls: FDWalls, u0):
    """ApplyBoundaryConditions on b = u0 (diffuclass.cpp:191-275, TimeStep :405-413)."""
    nX, nY = p.nW, p.nH
    h = p.h
    F = (p.D * p.dt) / (h * h)
    twoFh = 2 * F * h
    lN, rN, tN, bN = walls.Nc
    lBV, rBV, tBV, bBV = walls.BV
    top = nY - 1
    b = np.array(u0, dtype=np.float64, copy=True).reshape(nY, nX)
    tBVa = np.broadcast_to(np.asarray(tBV, dtype=np.float64), (nX,))
    bBVa = np.broadcast_to(np.asarray(bBV, dtype=np.float64), (nX,))
    for j in range(nY):
        for i in range(nX):
            if j == top:
                if tN != 0:
                    if (i != 0 or lN != 0) and (i != nX - 1 or rN != 0):
                        b[j, i] += (twoFh * tBVa[i]) / tN
                else:
                    b[j, i] = tBVa[i]
            elif j == 0:
                if bN != 0:
                    if (i != 0 or lN != 0) and (i != nX - 1 or rN != 0):
                        b[j, i] += (twoFh * bBVa[i]) / bN
                else:
                    b[j, i] = bBVa[i]
            if i == nX - 1:
                if rN != 0:
                    if (j != 0 or bN != 0) and (j != top or tN != 0):
                        b[j, i] += (twoFh * rBV) / rN
                else:
                    b[j, i] = rBV
            elif i == 0:
                if lN != 0:
                    if (j != 0 or bN != 0) and (j != top or tN != 0):
                        b[j, i] += (twoFh * lBV) / lN
                else:
                    b[j, i] = lBV
    return b.ravel()


def fd_solve(p: Problem, u0, walls: FDWalls | None = None):
    """One diffusionPETSc::stepDiffusion (diffuclass.cpp:108-118): KSPSolve(A, b) -- solved exactly."""
    import scipy.sparse.linalg as spla
    walls = walls or fd_walls_from_problem(p)
    return spla.splu(fd_assemble(p, walls)).solve(fd_rhs(p, walls, u0))


def fd_node_weights(p: Problem):
    """w*h^2: the cell share of a node (h^2 inside, h^2/2 on a wall, h^2/4 in a corner).  Multiplying the
    ghost-node rows by it makes MyMatMult's matrix symmetric -- the form the GPU solver's PCG works on."""
    wx = np.ones(p.nW); wx[0] = wx[-1] = 0.5
    wy = np.ones(p.nH); wy[0] = wy[-1] = 0.5
    return (np.outer(wy, wx) * p.h * p.h).ravel()


def _fd_c_args(p: Problem, walls: FDWalls):
    Dc = (C.c_double * 4)(*[float(v) for v in walls.Dc])
    Nc = (C.c_double * 4)(*[float(v) for v in walls.Nc])
    F = (p.D * p.dt) / (p.h * p.h)
    return Dc, Nc, F


def fd_matmult(p: Problem, walls: FDWalls, x):
    """MyMatMult restated in C (eq_oracle.c: eqo_fd_matmult) -- independent of fd_assemble's sparse matrix."""
    Dc, Nc, F = _fd_c_args(p, walls)
    y = np.empty(p.N)
    lib().eqo_fd_matmult(C.c_long(p.nW), C.c_long(p.nH), C.c_double(p.h), C.c_double(F), Dc, Nc,
                         _dp(np.ascontiguousarray(x, dtype=np.float64)), _dp(y))
    return y


def fd_rhs_c(p: Problem, walls: FDWalls, u0):
    Dc, Nc, F = _fd_c_args(p, walls)
    b = np.array(u0, dtype=np.float64, copy=True)
    tBV = np.ascontiguousarray(np.broadcast_to(np.asarray(walls.BV[2], dtype=np.float64), (p.nW,)))
    bBV = np.ascontiguousarray(np.broadcast_to(np.asarray(walls.BV[3], dtype=np.float64), (p.nW,)))
    lib().eqo_fd_apply_bc(C.c_long(p.nW), C.c_long(p.nH), C.c_double(p.h), C.c_double(F), Nc, _dp(tBV), _dp(bBV),
                          C.c_double(float(walls.BV[0])), C.c_double(float(walls.BV[1])), _dp(b))
    return b


def fd_step_krylov(p: Problem, u0, walls: FDWalls | None = None, rtol=1e-5, maxit=100000):
    """diffusionPETSc::stepDiffusion the way the reference runs it: ApplyBoundaryConditions, then the
    unpreconditioned (F)BiCGStab of diffuclass.cpp:386-392 at PETSc's default rtol.  -> (u, iterations, relres)"""
    walls = walls or fd_walls_from_problem(p)
    Dc, Nc, F = _fd_c_args(p, walls)
    b = fd_rhs_c(p, walls, u0)
    x = np.empty(p.N)
    rel = C.c_double(0.0)
    it = lib().eqo_fd_bicgstab(C.c_long(p.nW), C.c_long(p.nH), C.c_double(p.h), C.c_double(F), Dc, Nc, _dp(b), _dp(x),
                               C.c_double(rtol), C.c_long(maxit), C.byref(rel))
    return x, it, rel.value


# --------------------------------------------------------------------------
# The reference's OWN diffusionPETSc class (oracle/_ref/libeq_fd_ref.so, built by `make -C oracle ref`
# from /root/reference/diffuclass.{h,cpp} on the interface shim in oracle/shim_petsc/): used to pin the
# restatements above and to generate tests/golden/fd_ref.json.  None when the library is absent.
# --------------------------------------------------------------------------
_FDREF = None


def fd_ref_lib():
    global _FDREF
    if _FDREF is None:
        path = os.path.join(_HERE, "_ref", "libeq_fd_ref.so")
        if not os.path.exists(path):
            return None
        R = C.CDLL(path)
        R.ref_fd_create.restype = C.c_void_p
        R.ref_fd_create_dirichlet0.restype = C.c_void_p
        R.ref_fd_size.restype = C.c_long
        R.ref_fd_size.argtypes = [C.c_void_p]
        R.ref_fd_step.argtypes = [C.c_void_p, c_dp, C.c_double, c_dp]
        R.ref_fd_matmult.argtypes = [C.c_void_p, c_dp, c_dp]
        R.ref_fd_destroy.argtypes = [C.c_void_p]
        _FDREF = R
    return _FDREF


class FDReference:
    """diffusionPETSc itself, driven through its public surface (initData, solution_vector, stepDiffusion).
    width/height are the integer microns upstream passes; nodes = length*npm + 1 (diffuclass.cpp:358-359)."""

    def __init__(self, width, height, npm, dt, D, walls: FDWalls | None = None):
        R = fd_ref_lib()
        if R is None:
            raise RuntimeError("oracle/_ref/libeq_fd_ref.so missing: run `make -C oracle ref` where /root/reference exists")
        self.R = R
        if walls is None:
            self.h = R.ref_fd_create_dirichlet0(C.c_int(width), C.c_int(height), C.c_double(npm), C.c_double(dt), C.c_double(D))
        else:
            a = lambda t: (C.c_double * 4)(*[float(v) for v in t])
            self.h = R.ref_fd_create(C.c_int(width), C.c_int(height), C.c_double(npm), C.c_double(dt), C.c_double(D),
                                     a(walls.Dc), a(walls.Nc), a(walls.BV))
        self.N = int(R.ref_fd_size(self.h))

    def step(self, u0, rtol=1e-5):
        """-> (u1, rhs ApplyBoundaryConditions built, Krylov iterations of the stand-in solve)"""
        u = np.array(u0, dtype=np.float64, copy=True)
        rhs = np.empty(self.N)
        its = self.R.ref_fd_step(self.h, _dp(u), C.c_double(rtol), _dp(rhs))
        return u, rhs, int(its)

    def matmult(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.empty(self.N)
        rc = self.R.ref_fd_matmult(self.h, _dp(x), _dp(y))
        assert rc == 0, "call step() once first: the shell matrix is created on first use"
        return y

    def close(self):
        if self.h:
            self.R.ref_fd_destroy(self.h)
            self.h = None


# --------------------------------------------------------------------------
# The reference's OWN agent-based-model classes (oracle/_ref/libeq_cell_ref.so: eQabm, Ecoli, cpmEcoli, Strain
# compiled in place on the Chipmunk interface shim oracle/shim_cpm/): pins the cell <-> mesh restatements above
# and generates tests/golden/cells_ref.json.  None when the library is absent.
# --------------------------------------------------------------------------
_CELLREF = None


def cell_ref_lib():
    global _CELLREF
    if _CELLREF is None:
        path = os.path.join(_HERE, "_ref", "libeq_cell_ref.so")
        if not os.path.exists(path):
            return None
        R = C.CDLL(path)
        R.ref_abm_create.restype = C.c_void_p
        R.ref_abm_count.restype = C.c_long
        for name in ("ref_abm_add_cell", "ref_abm_count", "ref_abm_move_cell", "ref_abm_record", "ref_abm_point_in_cell",
                     "ref_abm_update_cells", "ref_abm_destroy"):
            getattr(R, name)
        _CELLREF = R
    return _CELLREF


class ABMReference:
    """eQabm itself on one identity-lookup HSL layer.  Cells are addressed in the reference's list order
    (newest first: forward_list push_front)."""

    def __init__(self, width, height, npm, Dx=1.0, Dy=1.0):
        R = cell_ref_lib()
        if R is None:
            raise RuntimeError("oracle/_ref/libeq_cell_ref.so missing: run `make -C oracle ref` where /root/reference exists")
        self.R = R
        self.h = C.c_void_p(R.ref_abm_create(C.c_int(width), C.c_int(height), C.c_double(npm), C.c_double(Dx), C.c_double(Dy)))
        self.npm = npm
        self.nW, self.nH = int(width * npm) + 1, int(height * npm) + 1

    def add_cell(self, x, y, angle, length, a0=100.0, a1=0.0):
        self.R.ref_abm_add_cell(self.h, C.c_double(x), C.c_double(y), C.c_double(angle), C.c_double(length),
                                C.c_double(a0), C.c_double(a1))

    def count(self):
        return int(self.R.ref_abm_count(self.h))

    def move_cell(self, k, a, b, calls=1):
        """a, b = (x, y, angle) of body halves A and B; then `calls` post-step updates (ratchet, poles)."""
        self.R.ref_abm_move_cell(self.h, C.c_long(k), C.c_double(a[0]), C.c_double(a[1]), C.c_double(a[2]),
                                 C.c_double(b[0]), C.c_double(b[1]), C.c_double(b[2]), C.c_int(calls))

    def records(self):
        n = self.count()
        rec = np.zeros((n, CELL_STRIDE))
        for k in range(n):
            self.R.ref_abm_record(self.h, C.c_long(k), rec[k].ctypes.data_as(c_dp))
        return rec

    def point_in_cell(self, k, x, y):
        return bool(self.R.ref_abm_point_in_cell(self.h, C.c_long(k), C.c_double(x), C.c_double(y)))

    def update_cells(self, u):
        u = np.array(u, dtype=np.float64, copy=True)
        N = self.nW * self.nH
        g = np.zeros(self.count())
        d11, d22, d12 = np.empty(N), np.empty(N), np.empty(N)
        self.R.ref_abm_update_cells(self.h, _dp(u), _dp(g), _dp(d11), _dp(d22), _dp(d12))
        return u, g, (d11, d22, d12)

    def close(self):
        if self.h:
            self.R.ref_abm_destroy(self.h)
            self.h = None


# --------------------------------------------------------------------------
# The reference's OWN P1 solver class (oracle/_ref/libeq_fenics_ref.so = /root/reference/src/fHSL.cpp compiled in
# place on the one-process DOLFIN interface shim of oracle/shim_dolfin/, see oracle/fenics_ref.cpp): fenicsInterface
# itself runs -- boundary decoding, Robin rates, form wiring, trap solve, wall flux, channel sub-steps, flux
# functional -- on top of the reference's own generated element kernels.  Pins `step` / `problem_from_parameters`
# and generates tests/golden/fenics_ref.json.  None when the library is absent.
# --------------------------------------------------------------------------
_FENICSREF = None


def fenics_ref_lib():
    global _FENICSREF
    if _FENICSREF is None:
        path = os.path.join(_HERE, "_ref", "libeq_fenics_ref.so")
        if not os.path.exists(path):
            return None
        R = C.CDLL(path)
        R.ref_fenics_create.restype = C.c_void_p
        R.ref_fenics_last_error.restype = C.c_char_p
        R.ref_fenics_total_boundary_flux.restype = C.c_double
        _FENICSREF = R
    return _FENICSREF


def bc_entry(kind, value=0.0):
    """One wall of eQ::data::parameters["boundaries"] as eQ::boundaryCondition writes it (src/eQ.h:397-418)."""
    if kind == "Dirichlet":
        return ["Dirichlet", [0.0, 1.0, float(value)]]
    if kind == "Neumann":
        return ["Neumann", [1.0, 0.0, float(value)]]
    if kind == "Robin":
        return ["Robin", [1.0 if value == 0.0 else float(value), 1.0, 0.0]]
    raise ValueError(kind)


def default_parameters(width, height, npm=2.0, **over):
    """The eQ::data::parameters keys the path reads, with the shipped defaults of src/main.cpp."""
    P = {"nodesPerMicronSignaling": float(npm), "lengthScaling": 5.0, "boundaryType": "DIRICHLET_0",
         "trapType": "NOWALLED", "simulationTrapWidthMicrons": width, "simulationTrapHeightMicrons": height,
         "simulationFlowRate": 120.0, "simulationChannelLengthLeft": 100.0, "simulationChannelLengthRight": 100.0,
         "channelSolverNumberIterations": 4}
    P.update(over)
    return P


def problem_from_parameters(P, dt, D, width, height, npm) -> Problem:
    """fenicsClassInit + createHSL's boundary decoding + setRobinBoundaryConditions restated
    (src/fHSL.cpp:37-53,242-243,281-283,331-364,436-574)."""
    import math
    nH = int(math.ceil(height * npm)) + 1
    nW = int(math.ceil(width * npm)) + 1
    v = float(P["simulationFlowRate"])
    left_rate, right_rate = robin_rates(v, D, float(P["simulationChannelLengthLeft"]), float(P["simulationChannelLengthRight"]))
    h = 1.0 / float(P["nodesPerMicronSignaling"])
    well = 10.0 * (25.0 / float(P["lengthScaling"])) * h
    bt = [NEUMANN] * 4
    bv = [0.0] * 4
    channels = False
    btype, ttype = P["boundaryType"], P["trapType"]
    if btype == "MICROFLUIDIC_TRAP":
        if ttype == "H_TRAP":
            left_rate = right_rate = v
        rates = (left_rate, right_rate)
        for w, name in ((LEFT, "left"), (RIGHT, "right"), (TOP, "top"), (BOTTOM, "bottom")):
            d = P["boundaries"][name][1]
            if d[0] == 0.0:
                if w in (TOP, BOTTOM) and d[2] == -1.0:
                    bt[w] = DIRICHLET_CHANNEL
                else:
                    bt[w], bv[w] = DIRICHLET, d[2]
            elif d[1] == 0.0:
                bt[w] = NEUMANN
            elif w in (LEFT, RIGHT):
                bt[w], bv[w] = ROBIN, rates[w]
        channels = ttype != "H_TRAP"
    elif btype == "DIRICHLET_UPDATE":
        bt = [DIRICHLET] * 4
    elif btype == "DIRICHLET_0":
        on = {"NOWALLED": (1, 1, 1, 1), "THREEWALLED": (0, 0, 0, 1), "TWOWALLED": (0, 0, 1, 1), "ONEWALLED": (1, 1, 0, 1)}.get(ttype, (0, 0, 0, 0))
        bt = [DIRICHLET if o else NEUMANN for o in on]
    elif btype == "NEUMANN_3WALLED_TEST":
        bt = [DIRICHLET, DIRICHLET, NEUMANN, DIRICHLET]
    else:
        bt = [DIRICHLET] * 4
    return Problem(nW=nW, nH=nH, h=width / (nW - 1), hy=height / (nH - 1), dt=dt, D=D, bc_type=tuple(bt), bc_value=tuple(bv),
                   channels=channels, channel_v=v, channel_r=(left_rate, right_rate),
                   channel_iters=int(P["channelSolverNumberIterations"]), well_scaling=well)


class FenicsReference:
    """fenicsInterface itself (one layer, one process)."""

    def __init__(self, P, dt, D, width, height, npm, channel_velocity=0.0):
        import json
        R = fenics_ref_lib()
        if R is None:
            raise RuntimeError("oracle/_ref/libeq_fenics_ref.so missing: run `make -C oracle ref` where /root/reference exists")
        self.R = R
        h = R.ref_fenics_create(json.dumps(P).encode(), C.c_double(dt), C.c_double(D), C.c_double(width), C.c_double(height),
                                C.c_double(npm), C.c_double(channel_velocity))
        if not h:
            raise RuntimeError("fenicsInterface: " + R.ref_fenics_last_error().decode())
        self.h = C.c_void_p(h)
        a, b, c = C.c_long(), C.c_long(), C.c_long()
        R.ref_fenics_sizes(self.h, C.byref(a), C.byref(b), C.byref(c))
        self.nW, self.nH, self.nC = a.value, b.value, c.value
        self.N = self.nW * self.nH

    def mesh(self):
        xy = np.zeros(2 * self.N)
        dof = np.zeros(self.N, dtype=np.int32)
        self.R.ref_fenics_mesh(self.h, _dp(xy), dof.ctypes.data_as(C.POINTER(C.c_int)))
        return xy.reshape(-1, 2), dof

    def lookup(self):
        t = np.zeros(self.N, dtype=np.int64)
        self.R.ref_fenics_lookup(self.h, t.ctypes.data_as(c_lp))
        return t.reshape(self.nH, self.nW)

    def set_field(self, u):
        self.R.ref_fenics_set_field(self.h, _dp(np.ascontiguousarray(u, dtype=np.float64)))

    def field(self):
        u = np.zeros(self.N)
        self.R.ref_fenics_get_field(self.h, _dp(u))
        return u

    def set_tensor(self, d11, d22, d12):
        self.R.ref_fenics_set_tensor(self.h, _dp(np.ascontiguousarray(d11)), _dp(np.ascontiguousarray(d22)), _dp(np.ascontiguousarray(d12)))

    def set_channels(self, top, bottom):
        self.R.ref_fenics_set_channels(self.h, _dp(np.ascontiguousarray(top)), _dp(np.ascontiguousarray(bottom)))

    def channels(self):
        t, b, ft, fb = (np.zeros(self.nC) for _ in range(4))
        self.R.ref_fenics_get_channels(self.h, _dp(t), _dp(b), _dp(ft), _dp(fb))
        return t, b, ft, fb

    def set_boundary_value(self, v):
        self.R.ref_fenics_set_boundary_value(self.h, C.c_double(v))

    def step(self):
        if self.R.ref_fenics_step(self.h) != 0:
            raise RuntimeError("fenicsInterface::stepDiffusion: " + self.R.ref_fenics_last_error().decode())

    def total_boundary_flux(self):
        return float(self.R.ref_fenics_total_boundary_flux(self.h))

    def robin(self):
        v = [C.c_double() for _ in range(5)]
        self.R.ref_fenics_robin(self.h, *[C.byref(x) for x in v])
        return dict(trap_left=v[0].value, trap_right=v[1].value, chan_left=v[2].value, chan_right=v[3].value, well=v[4].value)

    def close(self):
        if self.h:
            self.R.ref_fenics_destroy(self.h)
            self.h = None


def fenics_form_assemble(which, nW, nH, W, H, D, dt, f=0.0, rA=0.0, sA=0.0, rB=0.0, sB=0.0, tensor=None, u0=None):
    """a and L of one of the reference's three trap forms -- "hsl", "hslRobin", "hslD" (fenics/*.ufl) -- assembled by the
    reference's own generated wrappers and kernels on the DOLFIN shim, with arbitrary constants (source f, Robin rates
    rA/rB on the left/right wall and external concentrations sA/sB).  Returns (scipy CSR matrix, load vector)."""
    import scipy.sparse as sp
    R = fenics_ref_lib()
    if R is None:
        raise RuntimeError("oracle/_ref/libeq_fenics_ref.so missing")
    R.ref_form_assemble.restype = C.c_long
    N = nW * nH
    rows, cols = np.zeros(7 * N, dtype=np.int64), np.zeros(7 * N, dtype=np.int64)
    vals, b = np.zeros(7 * N), np.zeros(N)
    u0 = np.zeros(N) if u0 is None else np.ascontiguousarray(u0, dtype=np.float64)
    t = [np.ascontiguousarray(x, dtype=np.float64) for x in tensor] if tensor is not None else [None] * 3
    nnz = R.ref_form_assemble(C.c_int({"hsl": 0, "hslRobin": 1, "hslD": 2}[which]), C.c_int(nW - 1), C.c_int(nH - 1),
                              C.c_double(W), C.c_double(H), C.c_double(D), C.c_double(dt), C.c_double(f),
                              C.c_double(rA), C.c_double(sA), C.c_double(rB), C.c_double(sB), _dp(t[0]), _dp(t[1]), _dp(t[2]),
                              _dp(u0), rows.ctypes.data_as(c_lp), cols.ctypes.data_as(c_lp), _dp(vals), _dp(b))
    if nnz < 0:
        raise RuntimeError(R.ref_fenics_last_error().decode())
    return sp.csr_matrix((vals[:nnz], (rows[:nnz], cols[:nnz])), shape=(N, N)), b
