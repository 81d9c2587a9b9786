#!/usr/bin/env python
"""bench.py -- HSL diffusion steps/sec at the 2048^2 mesh (BASELINE.json metric).

A "step" = one pass of the hot path over one layer: gather (readHSL) ->
scatter-add (writeHSL) -> backward-Euler solve to rel. residual 1e-12 ->
boundary-flux functional, i.e. eQabm::updateCells' lambdas + fenicsInterface::
stepDiffusion (src/abm/eQabm.cpp:268-359, src/fHSL.cpp:98-161) for BASELINE
configs[2]: synthetic 2048^2-node trap, 20k rods, single GPU.

N > 1 (torchrun, one rank per GPU): one independent HSL layer per GPU, the
reference's own MPI model (src/simulation.cpp:628-645) -> weak scaling, no
data-path collective.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NW = NH = 2048
NCELLS = 20000
H, DT, D = 0.5, 0.1, 1200.0
NPM = 2.0
METRIC = "hsl_diffusion_steps_per_sec_2048x2048"
UNIT = "steps/s"
WORKLOAD = ("configs[2]: synthetic 2048x2048-node trap mesh (h=0.5, dt=0.1, D=1200, Dirichlet-0 walls), "
            "20k rod cells secreting/sampling, one HSL layer per GPU")


def ncu_traffic(kernel):
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of a level-0 kernel from the committed
    `ncu --set full` capture (profiles/ncu_traffic.json, written by profiles/ncu_summary.py); None if absent."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            return float(json.load(f)["kernels"][kernel]["dram_bytes"])
    except Exception:
        return None


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """SM clock + throttle reasons DURING the timed region.  NVML in-process (nvidia_ml_py), polled every 5 ms: the
    default timed region is about 0.1 s, shorter than nvidia-smi's start-up, which is why the first round's lines said
    "unavailable"; `nvidia-smi -lms` stays as the fallback when NVML cannot be loaded."""
    NAMES = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")

    def __init__(self, index, uuid=None):
        super().__init__(daemon=True)
        self.index = index
        self.uuid = uuid
        self.rows = []          # (sm_mhz, sm_max_mhz, set of reason names)
        self.proc = None
        self.halt = threading.Event()
        self.ready = threading.Event()
        self.source = None

    def _nvml_handle(self, nv):
        if self.uuid:
            for u in (f"GPU-{self.uuid}", str(self.uuid)):
                try:
                    return nv.nvmlDeviceGetHandleByUUID(u.encode() if isinstance(u, str) else u)
                except Exception:
                    pass
        idx = self.index
        vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        parts = [v for v in vis.split(",") if v.strip()]
        if parts and self.index < len(parts) and parts[self.index].strip().isdigit():
            idx = int(parts[self.index])
        return nv.nvmlDeviceGetHandleByIndex(idx)

    def _run_nvml(self):
        import pynvml as nv
        nv.nvmlInit()
        h = self._nvml_handle(nv)
        mx = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        bits = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        self.source = "nvml"
        while not self.halt.is_set():
            sm = float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
            r = int(get_reasons(h))
            self.rows.append((sm, mx, {n for n, b in bits.items() if r & b}))
            self.ready.set()
            time.sleep(0.005)
        nv.nvmlShutdown()

    def _run_smi(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        self.source = "nvidia-smi"
        self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}",
                                      "--format=csv,noheader,nounits", "-lms", "20"],
                                     stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        for line in self.proc.stdout:
            c = [v.strip() for v in line.split(",")]
            try:
                self.rows.append((float(c[0]), float(c[1]),
                                  {n for n, v in zip(self.NAMES, c[2:6]) if v.lower().startswith("active")}))
                self.ready.set()
            except Exception:
                continue
            if self.halt.is_set():
                break

    def run(self):
        try:
            self._run_nvml()
        except Exception:
            try:
                self._run_smi()
            except Exception:
                pass
        self.ready.set()

    def wait_ready(self, timeout=5.0):
        """Block until the first sample has arrived, so that a short timed region is covered."""
        self.ready.wait(timeout)
        self.mark = len(self.rows)

    def stop(self):
        self.halt.set()
        if self.proc:
            self.proc.terminate()
        self.join(timeout=2.0)
        rows = self.rows[getattr(self, "mark", 0):] or self.rows
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        reasons = set()
        for r in rows:
            reasons |= r[2]
        return {"sm_mhz": float(np.median([r[0] for r in rows])), "sm_max_mhz": float(max(r[1] for r in rows)),
                "reasons": sorted(reasons), "samples": len(rows), "source": self.source}


def cpu_baseline(sample_n=512, threads=1):
    """Restated reference CPU path ("Fenics-equivalent": re-assemble, DirichletBC, sparse direct LU
    every step, src/fHSL.cpp:104-108) on a bounded sample, scaled to the 2048^2 metric by DOF count."""
    from oracle import oracle as O
    p = O.Problem(nW=sample_n, nH=sample_n, h=H, dt=DT, D=D)
    cells = O.synthetic_colony(int(NCELLS * (sample_n / NW) ** 2), p.W, p.H)
    u = np.zeros(p.N)
    t0 = time.perf_counter()
    amount = 100.0 + 0.0 * O.gather(cells, NPM, p.nH, p.nW, u)
    u = O.scatter(cells, NPM, p.nH, p.nW, amount, u)
    s = O.new_state(p)
    s.u = u
    O.step(p, s, solver="lu")
    dt = time.perf_counter() - t0
    value = (1.0 / dt) * (p.N / float(NW * NH))
    return {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": (f"one full step (gather+scatter+assemble+DirichletBC+SuperLU factor/solve+flux) on a "
                       f"{sample_n}x{sample_n} mesh with {len(cells)} rods in {dt:.2f} s, scaled by DOF ratio "
                       f"{p.N}/{NW * NH} to the 2048^2 metric (LU cost is superlinear, so this flatters the CPU)"),
            "seconds": dt}


def cpu_baseline_fd():
    """diffusionPETSc's own CPU path restated (oracle/eq_oracle.c: ApplyBoundaryConditions + matrix-free
    MyMatMult + unpreconditioned BiCGStab to PETSc's default rtol 1e-5, diffuclass.cpp:386-413) on the FULL
    2048^2 workload, OpenMP over all host cores (upstream: one DMDA decomposition over MPI ranks)."""
    from oracle import oracle as O
    p = O.Problem(nW=NW, nH=NH, h=H, dt=DT, D=D)
    cells = O.synthetic_colony(NCELLS, p.W, p.H)
    u = np.zeros(p.N)
    steps, iters = 3, []
    t0 = time.perf_counter()
    for _ in range(steps):
        amount = 100.0 + 0.0 * O.gather(cells, NPM, p.nH, p.nW, u)
        u = O.scatter(cells, NPM, p.nH, p.nW, amount, u)
        u, it, _ = O.fd_step_krylov(p, u)
        iters.append(int(it))
    dt = (time.perf_counter() - t0) / steps
    cores = int(O.lib().eqo_num_threads())
    return {"value": 1.0 / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": (f"{steps} full steps at 2048x2048 with {len(cells)} rods (gather + scatter + ApplyBoundaryConditions "
                       f"+ matrix-free BiCGStab to rtol 1e-5, {iters} iterations; the GPU solves to 1e-12), "
                       f"{dt:.2f} s per step on {cores} OpenMP threads"),
            "seconds": dt * steps}


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU path (oracle port; the Fenics/PETSc original cannot be
    built here) timed on the host cores."""
    if rank != 0:
        return
    from oracle import oracle as O
    n = 320
    p = O.Problem(nW=n, nH=n, h=H, dt=DT, D=D)
    cells = O.synthetic_colony(int(NCELLS * (n / NW) ** 2), p.W, p.H)
    s = O.new_state(p)

    def one():
        amount = 100.0 + 0.0 * O.gather(cells, NPM, p.nH, p.nW, s.u)
        s.u = O.scatter(cells, NPM, p.nH, p.nW, amount, s.u)
        O.step(p, s, solver="lu")

    for _ in range(args.warmup):
        one()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        one()
    dt = (time.perf_counter() - t0) / args.steps
    value = (1.0 / dt) * (p.N / float(NW * NH))
    sample = (f"each step = one full step on a {n}x{n} mesh with {len(cells)} rods (SuperLU, 1 thread: one MPI "
              f"rank per layer upstream), scaled by DOF ratio to 2048^2")
    # Where the reference's own class was compiled (oracle/_ref/libeq_fenics_ref.so: src/fHSL.cpp on the one-process
    # DOLFIN interface shim), run it beside the port on a smaller sample: same answer, and the port is the FASTER of
    # the two (SuperLU against the shim's banded LU), so the headline ratio is taken against the stronger CPU number.
    ref_class = None
    if O.fenics_ref_lib() is not None:
        m, npm_ = 160, NPM
        Wm = (m - 1) / npm_
        P = O.default_parameters(int(round(Wm)), int(round(Wm)), npm_)
        try:
            F = O.FenicsReference(P, DT, D, float(int(round(Wm))), float(int(round(Wm))), npm_)
            q = O.problem_from_parameters(P, DT, D, float(int(round(Wm))), float(int(round(Wm))), npm_)
            rng = np.random.default_rng(3)
            u = rng.uniform(0, 50, q.N)
            F.set_field(u)
            t1 = time.perf_counter()
            F.step()
            t_ref = time.perf_counter() - t1
            sq = O.new_state(q)
            sq.u = u.copy()
            t1 = time.perf_counter()
            O.step(q, sq, solver="lu")
            t_port = time.perf_counter() - t1
            uf = F.field()
            ref_class = {"mesh": f"{q.nW}x{q.nH}", "seconds_reference_class_on_shim": t_ref, "seconds_port": t_port,
                         "rel_l2_port_vs_reference_class": float(np.linalg.norm(sq.u - uf) / np.linalg.norm(uf))}
            F.close()
        except Exception as e:  # the checker's checker must not take the arm down
            ref_class = {"error": str(e)}
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3 * (NW * NH) / p.N,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": {"workload": WORKLOAD},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    if ref_class is not None:
        line["reference_class_check"] = ref_class
    print(json.dumps(line), flush=True)


def run_slab(args, rank, local_rank, world):
    """--mode slab: BASELINE configs[4] style -- one mesh split into row slabs over the GPUs of the box,
    one-row halo exchange + CG all-reduce over NCCL.  16384 columns, 2048 rows and 25k rods per GPU
    (= the 16384^2 / 200k-rod configuration at 8 GPUs)."""
    import torch
    import torch.distributed as dist
    import eq_b200 as E
    from oracle import oracle as O
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    nW, nH, ncells = args.slab_cols, 2048 * world, 25000 * world
    ids = [E.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    stream = torch.cuda.current_stream()
    g = E.GpuHSL(nW, nH, h=H, dt=DT, D=D, device=local_rank, stream=stream.cuda_stream, slab=(rank, world, ids[0]))
    cells = O.synthetic_colony(ncells, (nW - 1) * H, (nH - 1) * H, seed=777)
    g.upload_cells(cells, NPM)
    g.set_amounts(np.full(len(cells), 100.0))

    def step():
        g.gather_resident()
        g.scatter_resident()
        g.step()

    for _ in range(args.warmup):
        step()
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = g.stats().kernel_launches
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    dist.barrier(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    if rank == 0:
        N = nW * nH
        print(json.dumps({
            "metric": f"hsl_diffusion_steps_per_sec_{nW}x{nH}_row_slab", "value": args.steps / (ms / 1e3), "unit": UNIT,
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"configs[4] style: {nW}x{nH} mesh in {world} row slabs of 2048 rows, {len(cells)} rods, "
                                   "fused tile kernels with 6-row NCCL halo exchange + CG all-reduce",
                       "pcg_iterations": int(g.stats().iterations), "relres": g.stats().relres,
                       "mg_levels": int(g.stats().levels), "dof_updates_per_sec": N * args.steps / (ms / 1e3)},
            "gpu_launches": int(g.stats().kernel_launches - l0)}), flush=True)
    g.close()
    dist.barrier()
    dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)    # 100 steps after 10 warm-up
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="eq_b200")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--mode", default="layers", choices=["layers", "slab"])
    ap.add_argument("--config", type=int, default=3, choices=[2, 3, 4, 6, 7],
                    help="BASELINE.json configs index+1: 3 = the headline 2048^2 workload (default); 2 = dual layers "
                         "with Robin walls (C4/C14 on alternating GPUs); 4 = channel-flow trap at 4096^2.  Widened "
                         "rows (SURVEY 8f): 6 = the headline workload on diffusionPETSc's finite-difference "
                         "discretisation; 7 = with the anisotropic tensor rasterised from the rods every step")
    ap.add_argument("--slab-cols", type=int, default=16384)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.mode == "slab":
        if world < 2:
            raise SystemExit("--mode slab needs torchrun with >= 2 ranks")
        run_slab(args, rank, local_rank, world)
        return

    import torch
    import torch.distributed as dist
    import eq_b200 as E
    from oracle import oracle as O  # synthetic colony generator + cpu_baseline only

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    stream = torch.cuda.current_stream()
    global NW, NH, WORKLOAD, METRIC
    kw = {}
    Dl = D
    if args.config == 2:    # H_TRAP: Robin left/right at the flow rate, Neumann top/bottom (src/fHSL.cpp:448-453)
        Dl = D if rank % 2 == 0 else 640.0     # C4HSL / C14HSL (src/eQinit.h:12-13)
        kw = dict(bc_type=(2, 2, 0, 0), bc_value=(120.0, 120.0, 0.0, 0.0))
        WORKLOAD = "configs[1]: dual QS layers (C4 D=1200 / C14 D=640 on alternating GPUs), Robin left/right r=120, 2048x2048, 20k rods"
        METRIC = "hsl_diffusion_steps_per_sec_2048x2048_robin"
    elif args.config == 4:  # MICROFLUIDIC_TRAP: Robin left/right, channel Dirichlet top/bottom, 48 CN sub-steps
        NW = NH = 4096
        rl, rr = O.robin_rates(120.0, D, 20.0, 20.0)
        kw = dict(bc_type=(2, 2, 3, 3), bc_value=(rl, rr, 0.0, 0.0), channels=True, channel_v=120.0,
                  channel_r=(rl, rr), channel_iters=48, well_scaling=10.0 * (25.0 / 5.0) * 0.5)
        WORKLOAD = "configs[3]: channel-flow trap (1-D advection-diffusion channels, 48 CN sub-steps) at 4096x4096, 20k rods"
        METRIC = "hsl_diffusion_steps_per_sec_4096x4096_channels"
    elif args.config == 6:  # diffusionPETSc (diffuclass.cpp): 5-point FD, DIRICHLET_0 as its initDiffusion wires it
        kw = dict(discretisation=E.DISC_FD)
        WORKLOAD = "SURVEY 8(f)4: configs[2] workload on diffusionPETSc's 5-point finite-difference discretisation (DIRICHLET_0), 2048x2048, 20k rods"
        METRIC = "hsl_diffusion_steps_per_sec_2048x2048_fd"
    elif args.config == 7:  # setDiffusionTensor from the rods each step (src/abm/eQabm.cpp:306-325), Dx=1.5, Dy=0.6
        WORKLOAD = "SURVEY 8(f)3: configs[2] workload with D11/D22/D12 rasterised from the 20k rods every step (axial 1.5, transverse 0.6), variable-tensor operator, 2048x2048"
        METRIC = "hsl_diffusion_steps_per_sec_2048x2048_tensor"
    tensor_feed = args.config == 7
    W = (NW - 1) * H
    cells = O.synthetic_colony(NCELLS, W, W, seed=12345 + rank)
    ncells = len(cells)
    amount = np.full(ncells, 100.0)  # nM per step

    def make_solver(warm=None):
        s = E.GpuHSL(NW, NH, h=H, dt=DT, D=Dl, device=local_rank, stream=stream.cuda_stream,
                     smooth_sweeps=int(os.environ.get("EQ_NU", "0")), **kw)
        if warm is not None:
            s.set_warm_start(warm)
        s.upload_cells(cells, NPM)
        s.set_amounts(amount)
        return s

    # Starting-guess policy.  The library default above 512^2 nodes is mode 6 (fixed extrapolations).  Mode 7, the image
    # ring (profiles/r01_guess_study.md), first ran on a B200 in the last seconds of round 1 (gpurun_out/ring_quick.json:
    # 1.46 iterations per step against 2.66, same field to 1e-13) and is opt-in in the library until the whole GPU suite
    # has seen it; the bench opts in, but only after checking it against mode 6 on this very workload first: 30 steps
    # with both, fields must agree to 1e-9, else the run falls back to the default and says so.
    warm_mode = int(os.environ["EQGPU_WARM"]) if "EQGPU_WARM" in os.environ else (4 if NW * NH <= 512 * 512 else 6)
    ring_check = None
    if "EQGPU_WARM" not in os.environ and NW * NH > 512 * 512 and args.config == 3:   # the headline workload, the one it has run on
        try:
            fields = {}
            for mode in (6, 7):
                c = make_solver(mode)
                for _ in range(30):
                    c.gather_resident()
                    c.scatter_resident()
                    c.step()
                fields[mode] = (c.get_field(), int(c.last_guess()))
                c.close()
            diff = float(np.linalg.norm(fields[7][0] - fields[6][0]) / np.linalg.norm(fields[6][0]))
            ring_check = {"steps": 30, "rel_l2_mode7_vs_mode6": diff, "mode7_last_guess": fields[7][1],
                          "ok": bool(diff < 1e-9 and fields[7][1] == 8)}
        except Exception as e:
            ring_check = {"error": str(e), "ok": False}
        if ring_check["ok"]:
            warm_mode = 7
    g = make_solver(None if "EQGPU_WARM" in os.environ else warm_mode)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    it_hist = []

    def step_resident():
        g.gather_resident()
        g.scatter_resident()
        if tensor_feed:
            g.cells_tensor(1.5, 0.6)
        g.step()
        it_hist.append(g.stats().iterations)   # host-side read of the last step's count (the step has synchronised)

    # pinned host buffers for the end-to-end legs
    rec_pin = torch.from_numpy(cells.copy()).pin_memory()
    amt_pin = torch.from_numpy(amount.copy()).pin_memory()
    out_pin = torch.zeros(ncells, dtype=torch.float64).pin_memory()
    fld_pin = torch.zeros(NW * NH, dtype=torch.float64).pin_memory()
    import ctypes as C
    dpp = lambda t: C.cast(t.data_ptr(), C.POINTER(C.c_double))
    L = E.lib()

    def step_e2e():
        # fused drop-in: only per-cell data crosses PCIe (cell records in, sampled HSL out, deposits in)
        g._ck(L.eqgpu_cells_upload(g._h, dpp(rec_pin), C.c_int64(ncells), C.c_double(NPM)))
        g._ck(L.eqgpu_cells_gather(g._h, dpp(out_pin)))
        g._ck(L.eqgpu_cells_scatter(g._h, dpp(amt_pin)))
        if tensor_feed:
            g.cells_tensor(1.5, 0.6)
        g.step()
        return g.stats().total_boundary_flux

    def step_compat():
        # strict drop-in: fenicsInterface's host solution_vector in and out every step
        g.step_host_ptr(fld_pin.data_ptr())

    def timed(fn, k):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(k):
            fn()
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    for _ in range(args.warmup):
        step_resident()
    try:
        uuid = str(torch.cuda.get_device_properties(local_rank).uuid)
    except Exception:
        uuid = None
    sampler = ClockSampler(local_rank, uuid)
    sampler.start()
    sampler.wait_ready()
    l0 = g.stats().kernel_launches
    ms = timed(step_resident, args.steps)
    launches = g.stats().kernel_launches - l0
    iters = g.stats().iterations
    iters_mean = float(np.mean(it_hist[-args.steps:]))
    relres = g.stats().relres
    clocks = sampler.stop()
    ms_step = ms / args.steps
    value = world * args.steps / (ms / 1e3)

    for _ in range(3):
        step_e2e()
    ms_e2e = timed(step_e2e, args.steps) / args.steps
    # compat leg: the run goes on, but the field now lives on the host as in the reference: each step the host
    # adds the deposits to solution_vector (writeHSL on the CPU, outside the timed region) and hands it in;
    # the solution comes back in the same buffer.  (Restarting every step from one fixed field would let the
    # warm start return the previous, identical answer in zero iterations.)
    fld_pin.copy_(torch.from_numpy(g.get_field()))
    dep_t = torch.from_numpy(O.scatter(cells, NPM, NH, NW, amount, np.zeros(NW * NH)))
    kc = max(3, args.steps // 5)
    ms_compat = 0.0
    for it in range(3 + kc):
        fld_pin.add_(dep_t)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        step_compat()
        e1.record(stream)
        torch.cuda.synchronize()
        if it >= 3:
            ms_compat += e0.elapsed_time(e1) / kc
    if world > 1:
        t = torch.tensor([ms_compat], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_compat = float(t.item())

    if rank == 0:
        peak, peak_src = peaks()
        # dominant kernel, timed alone with CUDA events on the launching stream
        roof = {}
        for name in ("presmooth", "postsmooth", "apply_p", "update_r", "update_xr"):
            try:
                kms, kbytes = g.bench_kernel(name, 50)
            except E.EqGpuError:
                continue
            roof[name] = {"ms": kms, "bytes": kbytes, "gbs": kbytes / (kms * 1e-3) / 1e9}
        dom = max(roof, key=lambda k: roof[k]["ms"])
        N = NW * NH
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "rods": ncells, "pcg_iterations": int(iters),
                       "pcg_iterations_mean": iters_mean,
                       "relres": relres, "rtol": 1e-12, "mg_levels": int(g.stats().levels),
                       "parallelism": f"layer-per-gpu x{world}",
                       "l2": "working set (6 fine fp64 vectors = 201 MB + MG hierarchy) exceeds the 126 MB L2; no explicit flush",
                       "initial_guess": ("image ring (warm mode 7, opt-in): fixed extrapolation through the last <= 7 solutions plus a "
                                         "least-squares correction in the backward-difference basis; stop test relative to the "
                                         "right-hand side (rtol 1e-12) whatever the guess") if warm_mode == 7 else
                                        "best of {zero, previous solution, linear / quadratic / cubic / quartic extrapolation "
                                        "of the previous solutions} (warm mode 6; mode 4, the default up to 512^2 nodes, has "
                                        "the least-squares combination of the last three instead of cubic and quartic), "
                                        "picked on the device by residual norm; stop test relative to the right-hand side "
                                        "(rtol 1e-12) whatever the guess",
                       "warm_mode": warm_mode, "ring_check": ring_check,
                       "last_guess": int(g.last_guess()),
                       "dof_updates_per_sec": value * N},
            "clocks": clocks,
            "e2e": {"value": world * 1e3 / ms_e2e, "unit": UNIT,
                    "h2d_bytes_per_step": int(rec_pin.numel() * 8 + amt_pin.numel() * 8),
                    "d2h_bytes_per_step": int(out_pin.numel() * 8 + 8),
                    "path": "eqgpu_cells_upload+gather+scatter+step with pinned host buffers (field stays in HBM)"},
            "e2e_compat": {"value": world * 1e3 / ms_compat, "unit": UNIT,
                           "h2d_bytes_per_step": N * 8, "d2h_bytes_per_step": N * 8,
                           "path": "eqgpu_step_host: fenicsInterface::stepDiffusion contract, full solution_vector in/out"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": f"k_{dom} (level 0, 2048^2)", "achieved": roof[dom]["gbs"],
                         "peak": peak, "unit": "GB/s", "frac": roof[dom]["gbs"] / peak,
                         "traffic": ncu_traffic(dom) if (NW, NH) == (2048, 2048) else None,
                         "traffic_note": "bytes/launch, dram__bytes_read.sum + dram__bytes_write.sum from the ncu "
                                         "--set full capture in profiles/ (below the algorithmic bytes where "
                                         "freshly written vectors are still in the 126 MB L2)",
                         "peak_source": peak_src, "kernels": roof,
                         "step_ideal_frac": (16.0 * N / (ms_step * 1e-3) / 1e9) / peak,
                         # whole step, algorithmic bytes of DESIGN.md section 5: per PCG iteration 124 B/DOF on
                         # level 0 (presmooth 18, postsmooth 26, apply_p 32, update_r 24, update_x 24) + 44/3 on the
                         # coarser levels; per step 224 B/DOF for the warm start (k_init_tile 96, k_impose 96,
                         # k_finish_x 32; isotropic one-GPU path with the five-solution history)
                         # (warm mode 7, image ring at depth 7: k_init_tile without history 24, copy of the right-hand
                         # side 16, k_ring_gram 64, k_ring_impose 136, k_finish_x 32, k_ring_image 24 = 296)
                         "step_algorithmic": (lambda fixed: (lambda bts: {
                             "bytes": bts, "gbs": bts / (ms_step * 1e-3) / 1e9, "frac": bts / (ms_step * 1e-3) / 1e9 / peak,
                             "formula": f"N*({fixed:.0f} + (124 + 44/3)*mean_iterations)"})(
                             N * (fixed + (124.0 + 44.0 / 3.0) * iters_mean)))(
                             296.0 if warm_mode == 7 else 224.0)},
        }
        if not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_fd() if args.config == 6 else cpu_baseline()
        print(json.dumps(line), flush=True)
    g.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
