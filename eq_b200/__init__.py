"""eq_b200 -- B200-native drop-in for eQ's HSL diffusion hot path.

Python is plumbing only: this module loads ``eq_b200/csrc/libeqgpu.so`` (the
C-ABI declared in ``include/eqgpu.h``) with ctypes and mirrors the public
surface of the reference's ``fenicsInterface`` (src/fHSL.h:333-449) --
``initDiffusion / stepDiffusion / solution_vector / totalBoundaryFlux /
setBoundaryValues`` -- so the parity tests read like calls on the reference
class.  There is no CPU path: if the CUDA library is missing or no GPU is
present the constructor raises.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libeqgpu.so")

NEUMANN, DIRICHLET, ROBIN, DIRICHLET_CHANNEL = 0, 1, 2, 3
DISC_P1, DISC_FD = 0, 1   # fenicsInterface's P1 finite elements / diffusionPETSc's 5-point finite differences
LEFT, RIGHT, TOP, BOTTOM = 0, 1, 2, 3
CELL_STRIDE = 16

API_SYMBOLS = [
    "eqgpu_default_params", "eqgpu_create", "eqgpu_destroy", "eqgpu_last_error",
    "eqgpu_set_field", "eqgpu_get_field", "eqgpu_set_tensor", "eqgpu_set_boundary_value",
    "eqgpu_step", "eqgpu_step_host", "eqgpu_get_stats", "eqgpu_get_channels",
    "eqgpu_set_channels", "eqgpu_get_channel_flux", "eqgpu_cells_upload", "eqgpu_cells_raster",
    "eqgpu_cells_gather", "eqgpu_cells_scatter", "eqgpu_apply_operator", "eqgpu_build_rhs",
    "eqgpu_field_device_ptr", "eqgpu_sync", "eqgpu_cells_set_amounts",
    "eqgpu_cells_gather_resident", "eqgpu_cells_scatter_resident", "eqgpu_cells_get_gathered",
    "eqgpu_bench_kernel", "eqgpu_create_slab", "eqgpu_nccl_unique_id", "eqgpu_slab_rows",
    "eqgpu_slab_plan", "eqgpu_set_scatter_mode", "eqgpu_solver_path", "eqgpu_set_warm_start",
    "eqgpu_last_guess", "eqgpu_cells_tensor", "eqgpu_get_tensor",
    "eqgpu_ls_solve3", "eqgpu_ring_solve", "eqgpu_cells_upload_device", "eqgpu_get_warm_start", "eqgpu_apply_preconditioner", "eqgpu_comm_stats",
    "eqgpu_set_nonconvergence_policy", "eqgpu_unconverged_steps", "eqgpu_comm_peer_stats",
]


class Params(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("nW", C.c_int32), ("nH", C.c_int32),
        ("hx", C.c_double), ("hy", C.c_double), ("dt", C.c_double), ("D", C.c_double),
        ("bc_type", C.c_int32 * 4), ("bc_value", C.c_double * 4), ("robin_s", C.c_double * 2),
        ("channels", C.c_int32), ("channel_iters", C.c_int32), ("channel_v", C.c_double),
        ("channel_r", C.c_double * 2), ("well_scaling", C.c_double), ("rtol", C.c_double),
        ("max_iters", C.c_int32), ("device", C.c_int32), ("stream", C.c_void_p),
        ("smooth_sweeps", C.c_int32), ("max_levels", C.c_int32), ("discretisation", C.c_int32),
        ("reserved", C.c_int32 * 7),
    ]


class Stats(C.Structure):
    _fields_ = [
        ("iterations", C.c_int32), ("levels", C.c_int32), ("relres", C.c_double),
        ("total_boundary_flux", C.c_double), ("kernel_launches", C.c_int64), ("steps", C.c_int64),
    ]


class EqGpuError(RuntimeError):
    pass


_lib = None


def build(verbose: bool = False) -> str:
    """Compile libeqgpu.so in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    subprocess.run(["make", "-C", os.path.join(_HERE, "csrc"), "-j4", "all"], check=True,
                   stdout=None if verbose else subprocess.DEVNULL)
    return LIB_PATH


def lib():
    """The loaded C-ABI.  Fails loudly when the CUDA library has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise EqGpuError(f"{LIB_PATH} is missing: build it with eq_b200.build() "
                             "(__graft_entry__.build()); there is no CPU fallback")
        L = C.CDLL(LIB_PATH)
        L.eqgpu_last_error.restype = C.c_char_p
        L.eqgpu_last_error.argtypes = [C.c_void_p]
        L.eqgpu_create.argtypes = [C.POINTER(Params), C.POINTER(C.c_void_p)]
        L.eqgpu_create_slab.argtypes = [C.POINTER(Params), C.c_int, C.c_int, C.c_char_p, C.POINTER(C.c_void_p)]
        L.eqgpu_destroy.argtypes = [C.c_void_p]
        L.eqgpu_destroy.restype = None
        L.eqgpu_default_params.restype = None
        for name in API_SYMBOLS:
            getattr(L, name)
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def nccl_unique_id() -> bytes:
    """128-byte NCCL id for eqgpu_create_slab; make it on rank 0 and hand it to the other ranks."""
    buf = C.create_string_buffer(128)
    rc = lib().eqgpu_nccl_unique_id(buf)
    if rc != 0:
        raise EqGpuError("eqgpu_nccl_unique_id failed (is libnccl.so.2 loadable?)")
    return buf.raw


def slab_plan(nH: int, world: int, rank: int, max_levels: int = 16):
    """Owned rows per multigrid level for one rank (host arithmetic only): [(g0, g1, rows), ...]."""
    a = (C.c_int32 * max_levels)()
    b = (C.c_int32 * max_levels)()
    n = (C.c_int32 * max_levels)()
    k = lib().eqgpu_slab_plan(C.c_int32(nH), C.c_int32(world), C.c_int32(rank), C.c_int32(max_levels), a, b, n)
    if k < 0:
        raise EqGpuError("bad slab plan arguments")
    return [(a[i], b[i], n[i]) for i in range(k)]


def ls_solve3(G, f, bb):
    """Host-only hook: the device's 3x3 least-squares solve of warm-start mode 4 -> (c[3], predicted ||r||^2)."""
    G, f = _f64(G), _f64(f)
    c = np.zeros(3)
    pred = C.c_double()
    rc = lib().eqgpu_ls_solve3(_dp(G), _dp(f), C.c_double(bb), _dp(c), C.byref(pred))
    if rc != 0:
        raise EqGpuError("eqgpu_ls_solve3 failed")
    return c, pred.value


def ring_solve(G, f):
    """Host-only hook: the device's K x K normal-equation solve of warm-start mode 7.  G = full symmetric K x K Gram
    matrix (packed here), f = right-hand side -> correction coefficients c[K]."""
    G, f = np.asarray(G, dtype=np.float64), _f64(f)
    K = f.size
    packed = _f64(np.concatenate([G[i, i:] for i in range(K)]))
    c = np.zeros(K)
    if lib().eqgpu_ring_solve(C.c_int(K), _dp(packed), _dp(f), _dp(c)) != 0:
        raise EqGpuError("eqgpu_ring_solve failed")
    return c


def default_params() -> Params:
    p = Params()
    lib().eqgpu_default_params(C.byref(p))
    return p


class GpuHSL:
    """One HSL layer on one GPU: the Python face of ``gpuHSL`` (eq_b200/host/gpuHSL.h),
    itself the drop-in for ``fenicsInterface`` (src/fHSL.h:333-449)."""

    def __init__(self, nW, nH, h=0.5, dt=0.1, D=1200.0,
                 bc_type=(DIRICHLET,) * 4, bc_value=(0.0,) * 4, robin_s=(0.0, 0.0),
                 channels=False, channel_v=120.0, channel_r=(0.0, 0.0), channel_iters=48,
                 well_scaling=25.0, rtol=1e-12, max_iters=200, device=0, stream=None,
                 smooth_sweeps=0, max_levels=0, hy=None, slab=None, discretisation=DISC_P1):
        """slab = (rank, world, nccl_id_bytes) selects the row-slab decomposition (eqgpu_create_slab)."""
        L = lib()
        p = default_params()
        p.nW, p.nH, p.hx, p.hy, p.dt, p.D = nW, nH, h, (hy if hy else h), dt, D
        for w in range(4):
            p.bc_type[w] = int(bc_type[w])
            p.bc_value[w] = float(bc_value[w])
        p.robin_s[0], p.robin_s[1] = robin_s
        p.channels = 1 if channels else 0
        p.channel_v = channel_v
        p.channel_r[0], p.channel_r[1] = channel_r
        p.channel_iters = channel_iters
        p.well_scaling = well_scaling
        p.rtol, p.max_iters, p.device = rtol, max_iters, device
        p.stream = stream
        p.smooth_sweeps, p.max_levels = smooth_sweeps, max_levels
        p.discretisation = int(discretisation)
        self.params = p
        self.nW, self.nH, self.N = nW, nH, nW * nH
        self._h = C.c_void_p()
        if slab is None:
            rc = L.eqgpu_create(C.byref(p), C.byref(self._h))
        else:
            rank, world, uid = slab
            rc = L.eqgpu_create_slab(C.byref(p), C.c_int(rank), C.c_int(world),
                                     C.c_char_p(uid) if uid is not None else None, C.byref(self._h))
        if rc != 0:
            raise EqGpuError(f"eqgpu_create failed ({rc}): {L.eqgpu_last_error(None).decode()}")
        # the members simulation.cpp touches
        self.solution_vector = np.zeros(self.N)
        self.totalBoundaryFlux = 0.0
        self.ncells = 0

    # -- plumbing ----------------------------------------------------------
    def _ck(self, rc):
        if rc != 0:
            raise EqGpuError(f"eqgpu error {rc}: {lib().eqgpu_last_error(self._h).decode()}")

    def close(self):
        if self._h:
            lib().eqgpu_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- fenicsInterface surface --------------------------------------------
    def stepDiffusion(self):
        """src/fHSL.cpp:98-161 with the host-vector contract (solution_vector in/out)."""
        self.solution_vector = _f64(self.solution_vector)
        self._ck(lib().eqgpu_step_host(self._h, _dp(self.solution_vector)))
        self.totalBoundaryFlux = self.stats().total_boundary_flux
        return self.solution_vector

    def setBoundaryValues(self, v: float):
        self._ck(lib().eqgpu_set_boundary_value(self._h, C.c_double(v)))

    def getBoundaryFlux(self):
        return {"totalFlux": self.totalBoundaryFlux}

    # -- device-resident path -----------------------------------------------
    def set_field(self, u):
        u = _f64(u)
        assert u.size == self.N
        self._ck(lib().eqgpu_set_field(self._h, _dp(u)))

    def path(self) -> dict:
        b = lib().eqgpu_solver_path(self._h)
        return {"fused": bool(b & 1), "slab": bool(b & 2), "slab_fused": bool(b & 4), "cluster_tail": bool(b & 8),
                "tiled_coarsest": bool(b & 16), "tensor": bool(b & 32), "register_tile_levels": (b >> 8) & 15}

    def slab_rows(self):
        a, b = C.c_int32(), C.c_int32()
        self._ck(lib().eqgpu_slab_rows(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def get_field(self, out=None):
        """Whole-field host array; in slab mode only this rank's owned rows are (over)written."""
        if out is not None:
            if not (isinstance(out, np.ndarray) and out.dtype == np.float64 and out.flags.c_contiguous
                    and out.size == self.N):
                raise ValueError("get_field(out=...) needs a C-contiguous float64 array of nW*nH elements "
                                 "(anything else would be copied and the caller's buffer never written)")
            self._ck(lib().eqgpu_get_field(self._h, _dp(out)))
            return out
        u = np.zeros(self.N)
        self._ck(lib().eqgpu_get_field(self._h, _dp(u)))
        return u

    def step(self):
        self._ck(lib().eqgpu_step(self._h))
        self.totalBoundaryFlux = self.stats().total_boundary_flux

    def step_host_ptr(self, ptr: int):
        """eqgpu_step_host on a raw host pointer (e.g. a pinned buffer)."""
        self._ck(lib().eqgpu_step_host(self._h, C.cast(ptr, C.POINTER(C.c_double))))

    def set_tensor(self, d11=None, d22=None, d12=None):
        if d11 is None:
            self._ck(lib().eqgpu_set_tensor(self._h, None, None, None))
        else:
            a, b, c = _f64(d11), _f64(d22), _f64(d12)
            self._ck(lib().eqgpu_set_tensor(self._h, _dp(a), _dp(b), _dp(c)))

    def stats(self) -> Stats:
        st = Stats()
        self._ck(lib().eqgpu_get_stats(self._h, C.byref(st)))
        return st

    def channels(self):
        t, b = np.empty(self.nW), np.empty(self.nW)
        self._ck(lib().eqgpu_get_channels(self._h, _dp(t), _dp(b)))
        return t, b

    def set_channels(self, top, bottom):
        t, b = _f64(top), _f64(bottom)
        self._ck(lib().eqgpu_set_channels(self._h, _dp(t), _dp(b)))

    def channel_flux(self):
        t, b = np.empty(self.nW), np.empty(self.nW)
        self._ck(lib().eqgpu_get_channel_flux(self._h, _dp(t), _dp(b)))
        return t, b

    def apply_operator(self, x, constrained=False):
        x = _f64(x)
        y = np.empty(self.N)
        self._ck(lib().eqgpu_apply_operator(self._h, _dp(x), _dp(y), C.c_int(1 if constrained else 0)))
        return y

    def apply_preconditioner(self, r):
        """z = B r: one V-cycle of the PCG preconditioner (verification hook; r zero on Dirichlet rows)."""
        r = _f64(r)
        z = np.empty(self.N)
        self._ck(lib().eqgpu_apply_preconditioner(self._h, _dp(r), _dp(z)))
        return z

    def build_rhs(self, u0):
        u0 = _f64(u0)
        b = np.empty(self.N)
        self._ck(lib().eqgpu_build_rhs(self._h, _dp(u0), _dp(b)))
        return b

    def field_device_ptr(self) -> int:
        p = C.c_void_p()
        self._ck(lib().eqgpu_field_device_ptr(self._h, C.byref(p)))
        return p.value

    def sync(self):
        self._ck(lib().eqgpu_sync(self._h))

    # -- cells (eQabm::updateCells lambdas) ----------------------------------
    def upload_cells(self, records, nodes_per_micron):
        rec = _f64(records).reshape(-1, CELL_STRIDE)
        self.ncells = rec.shape[0]
        self._ck(lib().eqgpu_cells_upload(self._h, _dp(rec), C.c_int64(self.ncells),
                                          C.c_double(nodes_per_micron)))

    def upload_cells_device(self, dev_ptr: int, ncells: int, nodes_per_micron):
        """eqgpu_cells_upload_device: records already in HBM (stream-ordered copy, no host sync)."""
        self.ncells = int(ncells)
        self._ck(lib().eqgpu_cells_upload_device(self._h, C.cast(dev_ptr, C.POINTER(C.c_double)), C.c_int64(self.ncells),
                                                 C.c_double(nodes_per_micron)))

    def raster(self, cap=512):
        counts = np.zeros(self.ncells, dtype=np.int32)
        nodes = np.full((self.ncells, cap), -1, dtype=np.int64)
        self._ck(lib().eqgpu_cells_raster(self._h, counts.ctypes.data_as(C.POINTER(C.c_int32)),
                                          nodes.ctypes.data_as(C.POINTER(C.c_int64)), C.c_int32(cap)))
        return counts, nodes

    def gather(self):
        out = np.empty(self.ncells)
        self._ck(lib().eqgpu_cells_gather(self._h, _dp(out)))
        return out

    def set_nonconvergence_policy(self, policy: int):
        """0: a step that stops at max_iters above rtol raises (EQGPU_ENOCONV); 1: report and continue."""
        self._ck(lib().eqgpu_set_nonconvergence_policy(self._h, int(policy)))

    def unconverged_steps(self) -> int:
        n = C.c_int64()
        self._ck(lib().eqgpu_unconverged_steps(self._h, C.byref(n)))
        return int(n.value)

    def set_warm_start(self, mode: int):
        """Starting guess of the PCG solve: 0 = field as given or zero, 1 = also the previous solution,
        2 = also the linear, 3 = also the quadratic extrapolation of the previous solutions, 4 = also the residual-minimising
        (least-squares) combination of the last three solutions, 5 = mode 3 plus the cubic extrapolation of the
        last four, 6 = mode 5 plus the quartic extrapolation of the last five, 7 = image ring (opt-in: the last seven
        solutions and their images, fixed extrapolation plus a least-squares correction in the backward-difference
        basis).  Default: 4 up to 512^2 nodes, 6 above."""
        self._ck(lib().eqgpu_set_warm_start(self._h, C.c_int(mode)))

    def comm_stats(self) -> dict:
        """Row-slab mode: cumulative communication counters of this rank (eqgpu_comm_stats)."""
        out = (C.c_int64 * 4)()
        self._ck(lib().eqgpu_comm_stats(self._h, out))
        peer = (C.c_int64 * 2)()
        self._ck(lib().eqgpu_comm_peer_stats(self._h, peer))
        return {"allreduce_calls": int(out[0]), "allreduce_doubles": int(out[1]), "halo_exchanges": int(out[2]),
                "halo_bytes_sent": int(out[3]), "peer_exchange_kernels": int(peer[0]), "peer_allreduces": int(peer[1])}

    def warm_mode(self) -> int:
        return int(lib().eqgpu_get_warm_start(self._h))

    def last_guess(self) -> int:
        """0 field as given, 1 zero, 2 previous solution, 3 linear, 4 quadratic extrapolation, 5 least-squares
        combination, 6 cubic, 7 quartic extrapolation, 8 image-ring guess (last step's start)."""
        return int(lib().eqgpu_last_guess(self._h))

    def set_scatter_mode(self, mode: int):
        """0 = direct global atomics, 1 = shared-memory-binned (dense colonies)."""
        self._ck(lib().eqgpu_set_scatter_mode(self._h, C.c_int(mode)))

    def scatter(self, amount_nM):
        a = _f64(amount_nM)
        assert a.size == self.ncells
        self._ck(lib().eqgpu_cells_scatter(self._h, _dp(a)))

    def cells_tensor(self, Dx: float, Dy: float):
        """setDiffusionTensor for every uploaded rod (src/abm/eQabm.cpp:306-325); becomes the solver's tensor."""
        self._ck(lib().eqgpu_cells_tensor(self._h, C.c_double(Dx), C.c_double(Dy)))

    def get_tensor(self):
        a, b, c = np.empty(self.N), np.empty(self.N), np.empty(self.N)
        self._ck(lib().eqgpu_get_tensor(self._h, _dp(a), _dp(b), _dp(c)))
        return a, b, c

    # -- device-resident cell ops (no host copies inside) ---------------------
    def set_amounts(self, amount_nM):
        a = _f64(amount_nM)
        assert a.size == self.ncells
        self._ck(lib().eqgpu_cells_set_amounts(self._h, _dp(a)))

    def gather_resident(self):
        self._ck(lib().eqgpu_cells_gather_resident(self._h))

    def scatter_resident(self):
        self._ck(lib().eqgpu_cells_scatter_resident(self._h))

    def get_gathered(self):
        out = np.empty(self.ncells)
        self._ck(lib().eqgpu_cells_get_gathered(self._h, _dp(out)))
        return out

    def bench_kernel(self, name: str, reps: int = 20):
        ms, nbytes = C.c_double(), C.c_double()
        self._ck(lib().eqgpu_bench_kernel(self._h, name.encode(), C.c_int(reps), C.byref(ms), C.byref(nbytes)))
        return ms.value, nbytes.value
