#!/usr/bin/env python
"""bench.py -- HSL diffusion steps/sec at the 2048^2 mesh (BASELINE.json metric).

A "step" = one pass of the hot path over one layer: cell records for this step (the rods have moved) ->
gather (readHSL) -> scatter-add (writeHSL) -> backward-Euler solve to rel. residual 1e-12 -> boundary-flux
functional, i.e. eQabm::updateCells' lambdas + fenicsInterface::stepDiffusion (src/abm/eQabm.cpp:254-425,
src/fHSL.cpp:98-161) for BASELINE configs[2]: synthetic 2048^2-node trap, 20k rods, single GPU.

The colony CHANGES every step by default (--colony moving: rods advance 0.05-0.2 node per step and turn; `growing`
adds exponential growth with the ratchet and division) because eQ recomputes every rod's node set each step; the
static colony of round 1 (whose history-based starting guesses are unrepresentatively good) and a cold start are
timed beside it and reported as value_static / value_cold.

N > 1 (torchrun, one rank per GPU): one independent HSL layer per GPU, the reference's own MPI model
(src/simulation.cpp:628-645) -> weak scaling, no data-path collective; plus a second leg, `slab`: ONE mesh of
16384 x (2048 N) nodes cut into N row slabs with NCCL halo exchange and CG all-reduce (BASELINE configs[4]),
preceded by an in-run parity check of the slab solver.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--colony static|moving|growing]
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
    """STATIC cross-reference, not measured in this run: DRAM bytes per launch (dram__bytes_read.sum +
    dram__bytes_write.sum) of a level-0 kernel from the committed `ncu --set full` capture
    (profiles/ncu_traffic.json, written by profiles/ncu_summary.py); None if absent."""
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
    """SM clock + throttle reasons DURING the timed region.  NVML in-process (nvidia_ml_py), polled every 5 ms;
    `nvidia-smi -lms` is the fallback when NVML cannot be loaded."""
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


# ---------------------------------------------------------------------------------------------------------------
# workload: the record sets of a colony that changes (eq_b200/colony.py; host side, numpy)
# ---------------------------------------------------------------------------------------------------------------
def colony_record_sets(mode, n, W, Hh, nsets, seed):
    """[nsets, ncells, 16] cell records of successive steps (one set for a static colony)."""
    from eq_b200.colony import Colony
    col = Colony(n, W, Hh, npm=NPM, mode=mode, seed=seed, dt=DT)
    sets = [col.records()]
    for _ in range(1 if mode == "static" else nsets - 1):
        col.advance()
        sets.append(col.records())
    if mode == "static":
        sets = sets[:1]
    return np.ascontiguousarray(np.stack(sets))


# Algorithmic bytes per DOF of the once-per-step passes, by starting-guess mode: k_init_tile
# (reads u0 and the history tiles it walks, writes r1, b~ and the difference images), k_impose / the ring pair (reads
# what the picked guess combines, writes u and r), k_finish_x (32), the ring's image walk (16).
FIXED_BYTES = {0: 88, 1: 112, 2: 144, 3: 176, 4: 208, 5: 200, 6: 224, 7: 272}


def pingpong(k, R):
    """0 1 .. R-1 R-2 .. 1 0 1 ..: every step of the replay is one small move of the rods, never a jump back."""
    if R == 1:
        return 0
    k %= 2 * (R - 1)
    return k if k < R else 2 * (R - 1) - k


# ---------------------------------------------------------------------------------------------------------------
# reference arm: the reference's CPU path on the box's host cores, at the FULL configuration, nothing extrapolated
# ---------------------------------------------------------------------------------------------------------------
def run_reference(args, rank, world):
    """--impl reference.  The Fenics/PETSc executable cannot be built here, so this times the
    restated reference CPU path (oracle/, `kind: port`) on configs[2] exactly: 2048^2 nodes, 20k rods that move,
    every step = gather + scatter + one backward-Euler solve.
      value            diffusionPETSc's own algorithm (diffuclass.cpp:191-275,386-413,786-862: ApplyBoundaryConditions +
                       matrix-free MyMatMult + unpreconditioned BiCGStab to PETSc's default rtol 1e-5), OpenMP on all
                       host cores, for exactly --warmup + --steps steps.  It is the FASTER of the reference's two
                       solvers and runs at the reference's own (looser) tolerance, so value/this is conservative.
      matched_tolerance  the P1 operator of fenics/hslD.ufl (what fenicsInterface assembles) solved to the GPU arm's
                       rtol 1e-12 by Jacobi-CG on all cores, timed on a bounded number of full-size steps.
      lu               fenicsInterface's own solver is a sparse LU every step (src/fHSL.cpp:104-108): SuperLU of this
                       matrix is NOT run inside the driver's run (one factorisation takes minutes); the measured
                       figure from profiles/r02_cpu_lu.json is quoted, labelled as such.
    Under torchrun rank 0 alone runs; the N layers of the GPU arm would share the same host cores, so the CPU
    aggregate does not grow with N."""
    if rank != 0:
        return
    from oracle import oracle as O
    # all the host threads it can use: torchrun exports OMP_NUM_THREADS=1 to its ranks, which would time one core
    try:
        ncpu = len(os.sched_getaffinity(0))
    except AttributeError:
        ncpu = os.cpu_count() or 1
    O.lib().eqo_set_num_threads(int(ncpu))
    p = O.Problem(nW=NW, nH=NH, h=H, dt=DT, D=D)
    total = args.warmup + args.steps
    nsets = min(total + 1, 64)
    recs = colony_record_sets(args.colony, NCELLS, p.W, p.H, nsets, 12345)
    ncells = recs.shape[1]
    amount = np.full(ncells, 100.0)
    cores = int(O.lib().eqo_num_threads())
    u = np.zeros(p.N)
    iters = []

    def fd_step(k):
        nonlocal u
        c = recs[pingpong(k, len(recs))]
        O.gather(c, NPM, p.nH, p.nW, u)
        u0 = O.scatter(c, NPM, p.nH, p.nW, amount, u)
        u, it, _ = O.fd_step_krylov(p, u0)
        iters.append(int(it))

    t_start = time.perf_counter()
    for k in range(args.warmup):
        fd_step(k)
    t0 = time.perf_counter()
    for k in range(args.steps):
        fd_step(args.warmup + k)
    dt = (time.perf_counter() - t0) / args.steps
    value = 1.0 / dt
    # matched tolerance, bounded: 1-2 further full-size steps with the P1 operator to 1e-12
    matched = None
    if not args.no_matched:
        k_m = 1 if args.bounded else 2
        um = u.copy()
        t1 = time.perf_counter()
        its = []
        for k in range(k_m):
            c = recs[pingpong(total + k, len(recs))]
            O.gather(c, NPM, p.nH, p.nW, um)
            u0 = O.scatter(c, NPM, p.nH, p.nW, amount, um)
            um, it, rel = O.solve_cg(p, u0, rtol=1e-12, x0=um)
            its.append(int(it))
        dm = (time.perf_counter() - t1) / k_m
        matched = {"value": 1.0 / dt if dm <= 0 else 1.0 / dm, "unit": UNIT, "seconds_per_step": dm, "steps_timed": k_m,
                   "solver": "P1 operator (fenics/hslD.ufl) assembled to 7 bands + Jacobi-preconditioned CG, rtol 1e-12, "
                             "started from the previous solution", "iterations": its, "cores": cores}
    lu = None
    try:
        with open(os.path.join(ROOT, "profiles", "r02_cpu_lu.json")) as f:
            lu = json.load(f)
            lu["note"] = "quoted from profiles/r02_cpu_lu.json (measured once, outside this run)"
    except Exception:
        pass
    sample = (f"{args.warmup}+{args.steps} FULL steps at {NW}x{NH} with {ncells} rods ({args.colony} colony): gather + "
              f"scatter + ApplyBoundaryConditions + matrix-free BiCGStab to rtol 1e-5 (diffusionPETSc restated; "
              f"{int(np.mean(iters[-args.steps:]))} iterations per step), {dt:.2f} s per step on {cores} OpenMP threads; "
              f"nothing scaled or extrapolated")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": {"workload": WORKLOAD, "colony": args.colony, "rods": ncells,
                                            "solver": "diffusionPETSc path: 5-point FD, unpreconditioned BiCGStab, rtol 1e-5",
                                            "same_config": True, "extrapolated": False},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "matched_tolerance": matched, "lu": lu,
            "scales_with_n": False,
            "scales_with_n_note": "the N layers of an N-GPU run share this box's host cores: the CPU's aggregate "
                                  "layer-steps/s is this figure at every N",
            "wall_seconds": time.perf_counter() - t_start}
    print(json.dumps(line), flush=True)


def cpu_baseline_subprocess(colony):
    """cpu_baseline of the GPU arm: the reference arm on a bounded sample (3 + 1 full-size steps, about 20-40 s of CPU
    work), run in a SUBPROCESS so that the GPU arm's own process never maps anything under oracle/."""
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "3",
           "--bounded", "--colony", colony]
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "OMP_NUM_THREADS")}
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
        line = json.loads(r.stdout.strip().splitlines()[-1])
        cb = line["cpu_baseline"]
        cb["matched_tolerance"] = line.get("matched_tolerance")
        cb["lu"] = line.get("lu")
        return cb
    except Exception as e:
        return {"value": None, "unit": UNIT, "cores": None, "kind": "port", "sample": f"failed: {e}"}


def oracle_lu_subprocess(nW, nH, seed, path):
    """Checker for the slab leg: the oracle's direct solve of one seeded step, in a subprocess (see above)."""
    code = ("import sys, numpy as np; sys.path.insert(0, %r); from oracle import oracle as O;"
            "p = O.Problem(nW=%d, nH=%d, h=%r, dt=%r, D=%r);"
            "u0 = np.random.default_rng(%d).uniform(0.0, 50.0, p.N); np.save(%r, O.solve_lu(p, u0))"
            % (ROOT, nW, nH, H, DT, D, seed, path))
    subprocess.run([sys.executable, "-c", code], check=True, timeout=600)


# ---------------------------------------------------------------------------------------------------------------
# slab leg (N > 1): BASELINE configs[4] -- one mesh in N row slabs, halo exchange + CG all-reduce over NCCL
# ---------------------------------------------------------------------------------------------------------------
def run_slab_leg(args, rank, local_rank, world, stream, ids_fn):
    import torch
    import torch.distributed as dist
    import eq_b200 as E
    out = {}
    # (1) parity first: a 384 x 256 problem on the same N ranks against the oracle's LU, <= 1e-8 or no slab number
    pW, pH, seed = 384, max(256, 32 * world), 4242
    path = f"/tmp/eq_b200_slab_check_{os.getpid()}.npy"
    ok = True
    if rank == 0:
        try:
            oracle_lu_subprocess(pW, pH, seed, path)
        except Exception as e:
            ok = False
            out["parity"] = {"error": f"checker failed: {e}"}
    g = E.GpuHSL(pW, pH, h=H, dt=DT, D=D, device=local_rank, stream=stream.cuda_stream, slab=(rank, world, ids_fn()))
    u0 = np.random.default_rng(seed).uniform(0.0, 50.0, pW * pH)
    g.set_field(u0)
    g.step()
    mine = np.zeros(pW * pH)
    g.get_field(out=mine)
    r0, r1 = g.slab_rows()
    part = torch.zeros(pW * pH, dtype=torch.float64, device="cuda")
    part[r0 * pW:r1 * pW] = torch.from_numpy(mine[r0 * pW:r1 * pW]).cuda()
    dist.all_reduce(part)
    it_small = int(g.stats().iterations)
    g.close()
    flag = torch.tensor([1.0 if ok else 0.0], device="cuda", dtype=torch.float64)
    if rank == 0 and ok:
        ref = np.load(path)
        os.remove(path)
        err = float(np.linalg.norm(part.cpu().numpy() - ref) / np.linalg.norm(ref))
        out["parity"] = {"mesh": f"{pW}x{pH}", "ranks": world, "rel_l2_vs_oracle_lu": err, "tolerance": 1e-8,
                         "pcg_iterations": it_small, "ok": bool(err <= 1e-8)}
        flag[0] = 1.0 if err <= 1e-8 else 0.0
    dist.broadcast(flag, src=0)
    if flag.item() == 0.0:
        out["skipped"] = "slab parity self-check failed: no slab number is reported"
        return out
    # (2) the timed leg: 16384 columns, 2048 rows and 25k rods per GPU (= 16384^2 / 200k rods at 8 GPUs)
    nW, nH, ncells = args.slab_cols, 2048 * world, 25000 * world
    g = E.GpuHSL(nW, nH, h=H, dt=DT, D=D, device=local_rank, stream=stream.cuda_stream, slab=(rank, world, ids_fn()))
    k_steps, k_warm = max(3, min(args.steps, 20)), 3
    recs = colony_record_sets(args.colony, ncells, (nW - 1) * H, (nH - 1) * H, k_steps + k_warm + 1, 777)
    rec_dev = torch.from_numpy(recs).cuda()
    nrec = recs.shape[1]
    stride = nrec * 16 * 8
    g.upload_cells(recs[0], NPM)
    g.set_amounts(np.full(nrec, 100.0))
    its = []

    def step(k):
        g.upload_cells_device(rec_dev.data_ptr() + pingpong(k, len(recs)) * stride, nrec, NPM)
        g.gather_resident()
        g.scatter_resident()
        g.step()
        its.append(int(g.stats().iterations))

    for k in range(k_warm):
        step(k)
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = g.stats().kernel_launches
    c0 = g.comm_stats()
    e0.record(stream)
    for k in range(k_steps):
        step(k_warm + k)
    e1.record(stream)
    dist.barrier(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item()) / k_steps
    st = g.stats()
    c1 = g.comm_stats()
    it_sum = max(1, sum(its[-k_steps:]))
    comm = {"allreduce_calls_per_step": (c1["allreduce_calls"] - c0["allreduce_calls"]) / k_steps,
            "allreduce_calls_per_iteration": (c1["allreduce_calls"] - c0["allreduce_calls"] - 2 * k_steps) / it_sum,
            "allreduce_doubles_per_step": (c1["allreduce_doubles"] - c0["allreduce_doubles"]) / k_steps,
            "halo_exchanges_per_step": (c1["halo_exchanges"] - c0["halo_exchanges"]) / k_steps,
            "halo_bytes_sent_per_step_this_rank": (c1["halo_bytes_sent"] - c0["halo_bytes_sent"]) / k_steps,
            "peer_exchange_kernels_per_step": (c1["peer_exchange_kernels"] - c0["peer_exchange_kernels"]) / k_steps,
            "peer_allreduces_per_step": (c1["peer_allreduces"] - c0["peer_allreduces"]) / k_steps,
            "transport": ("peer memory over NVLink/NVSwitch (CUDA IPC): one kernel per halo exchange (boundary rows stored into the neighbours' staging buffers, theirs copied out of mine), scalar "
                          "all-reduces by one warp summing every rank's partials in rank order; NCCL only for the set-up and "
                          "the per-cell sample reduction" if c1["peer_exchange_kernels"] >= 0 else
                          "NCCL send/recv + all-reduce over NVLink/NVSwitch, stream-ordered")}
    out.update({"metric": f"hsl_diffusion_steps_per_sec_{nW}x{nH}_row_slab", "value": 1e3 / ms, "unit": UNIT,
                "ms_per_step": ms, "steps": k_steps, "warmup": k_warm, "scaling": "weak",
                "workload": f"configs[4] style: {nW}x{nH} nodes in {world} row slabs of 2048 rows, {nrec} rods "
                            f"({args.colony} colony)",
                "pcg_iterations_mean": float(np.mean(its[-k_steps:])), "relres": st.relres,
                "mg_levels": int(st.levels), "dof_updates_per_sec": nW * nH * 1e3 / ms,
                "gpu_launches": int(st.kernel_launches - l0), "comm": comm})
    g.close()
    return out


def run_slab_baseline(args, local_rank, stream):
    """N = 1: the weak-scaling baseline of the slab leg -- ONE GPU solving what one rank of the slab leg owns (16384 x 2048
    nodes, 25k moving rods), so that the driver's 1/2/4/8 runs give the slab curve from its first point.  Two figures:
    `value` goes through eqgpu_create_slab with one rank, which resolves to the one-GPU path (a single rank has no slab:
    register-tile smoothers, iteration graph) -- the denominator of weak-scaling efficiency is therefore the BEST one-GPU
    time for a rank's share, and everything a rank of the N > 1 leg loses (tile smoothers instead of register tiles, no
    iteration graph, exchanges, reductions, rank skew) counts against the efficiency; `single_gpu_path` is the same solve
    through eqgpu_create, as a cross-check.  Same step, same timing as run_slab_leg."""
    import torch
    import eq_b200 as E
    nW, nH, ncells = args.slab_cols, 2048, 25000
    k_steps, k_warm = max(3, min(args.steps, 20)), 3
    recs = colony_record_sets(args.colony, ncells, (nW - 1) * H, (nH - 1) * H, k_steps + k_warm + 1, 777)
    rec_dev = torch.from_numpy(recs).cuda()
    nrec = recs.shape[1]
    stride = nrec * 16 * 8

    def leg(g):
        g.upload_cells(recs[0], NPM)
        g.set_amounts(np.full(nrec, 100.0))
        its = []

        def step(k):
            g.upload_cells_device(rec_dev.data_ptr() + pingpong(k, len(recs)) * stride, nrec, NPM)
            g.gather_resident()
            g.scatter_resident()
            g.step()
            its.append(int(g.stats().iterations))

        for k in range(k_warm):
            step(k)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for k in range(k_steps):
            step(k_warm + k)
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / k_steps
        st = g.stats()
        return ms, float(np.mean(its[-k_steps:])), st

    g = E.GpuHSL(nW, nH, h=H, dt=DT, D=D, device=local_rank, stream=stream.cuda_stream)
    ms1, it1, st1 = leg(g)
    g.close()
    out = {"metric": f"hsl_diffusion_steps_per_sec_{nW}x{nH}_row_slab", "unit": UNIT, "steps": k_steps, "warmup": k_warm,
           "scaling": "weak", "ranks": 1,
           "workload": f"weak-scaling baseline of the slab leg: {nW}x{nH} nodes on one GPU (what one rank of the N-GPU slab "
                       f"leg owns), {nrec} rods ({args.colony} colony)",
           "single_gpu_path": {"value": 1e3 / ms1, "ms_per_step": ms1, "pcg_iterations_mean": it1, "relres": st1.relres,
                               "note": "the tuned one-GPU path on the same mesh, no decomposition"}}
    try:
        g = E.GpuHSL(nW, nH, h=H, dt=DT, D=D, device=local_rank, stream=stream.cuda_stream, slab=(0, 1, E.nccl_unique_id()))
        ms, it, st = leg(g)
        out.update({"value": 1e3 / ms, "ms_per_step": ms, "pcg_iterations_mean": it, "relres": st.relres,
                    "mg_levels": int(st.levels), "dof_updates_per_sec": nW * nH * 1e3 / ms,
                    "note": "eqgpu_create_slab with one rank = the one-GPU path (no decomposition): the best one-GPU time for a rank's share"})
        g.close()
    except Exception as e:   # no NCCL library: the tuned path is all there is to report
        out.update({"value": 1e3 / ms1, "ms_per_step": ms1, "pcg_iterations_mean": it1, "relres": st1.relres,
                    "note": f"slab path with one rank unavailable ({e}); this is the single-GPU path"})
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)    # 100 steps after 10 warm-up
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="eq_b200")
    ap.add_argument("--colony", default="moving", choices=["static", "moving", "growing"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-slab", action="store_true", help="N > 1: skip the row-slab leg")
    ap.add_argument("--no-side-legs", action="store_true", help="skip value_static / value_cold / e2e_compat")
    ap.add_argument("--bounded", action="store_true", help="reference arm: the bounded cpu_baseline sample")
    ap.add_argument("--no-matched", action="store_true", help="reference arm: skip the matched-tolerance steps")
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

    import ctypes as C
    import torch
    import torch.distributed as dist
    import eq_b200 as E

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    stream = torch.cuda.current_stream()

    def nccl_ids():
        ids = [E.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        return ids[0]

    if args.mode == "slab":   # the slab leg alone (development aid; the default N > 1 run carries it as `slab`)
        if world < 2:
            raise SystemExit("--mode slab needs torchrun with >= 2 ranks")
        slab = run_slab_leg(args, rank, local_rank, world, stream, nccl_ids)
        if rank == 0:
            print(json.dumps({"n_gpus": world, "higher_is_better": True, "dtype": "f64", "data": "synthetic", **slab}), flush=True)
        dist.barrier()
        dist.destroy_process_group()
        return

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
        # Robin rates of src/fHSL.cpp:331-364 for v = 120, D = 1200, channel lengths 20 / 20
        pe = 120.0 * 20.0 / D
        rl, rr = 120.0 / (1.0 - np.exp(-pe)), 120.0 / (np.exp(pe) - 1.0)
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
    N = NW * NH

    # record sets: enough distinct steps for the longest leg, replayed forwards then backwards
    nsets = min(args.warmup + args.steps + 1, 128)
    recs = colony_record_sets(args.colony, NCELLS, W, W, nsets, 12345 + rank)
    ncells = recs.shape[1]
    rec_dev = torch.from_numpy(recs).cuda()
    rec_pin = torch.from_numpy(recs).pin_memory()
    stride = ncells * 16 * 8
    amount = np.full(ncells, 100.0)  # nM per step
    warm_env = "EQGPU_WARM" in os.environ

    def make_solver(warm=None):
        s = E.GpuHSL(NW, NH, h=H, dt=DT, D=Dl, device=local_rank, stream=stream.cuda_stream,
                     smooth_sweeps=int(os.environ.get("EQ_NU", "0")), **kw)
        if warm is not None and not warm_env:
            s.set_warm_start(warm)
        s.upload_cells(recs[0], NPM)
        s.set_amounts(amount)
        return s

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, k0, k):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for i in range(k):
            fn(k0 + i)
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    def resident_leg(s, sets, warmup, steps, sample=False):
        """W untimed + K timed device-resident steps on solver s; the records of step k are record set pingpong(k)."""
        R = len(sets)
        its = []

        def step(k):
            if R > 1:
                s.upload_cells_device(rec_dev.data_ptr() + pingpong(k, R) * stride, ncells, NPM)
            s.gather_resident()
            s.scatter_resident()
            if tensor_feed:
                s.cells_tensor(1.5, 0.6)
            s.step()
            its.append(int(s.stats().iterations))   # host-side read of the last step's count (the step has synchronised)

        for k in range(warmup):
            step(k)
        sampler = None
        if sample:
            try:
                uuid = str(torch.cuda.get_device_properties(local_rank).uuid)
            except Exception:
                uuid = None
            sampler = ClockSampler(local_rank, uuid)
            sampler.start()
            sampler.wait_ready()
        l0 = s.stats().kernel_launches
        ms = timed(step, warmup, steps)
        res = {"ms_per_step": ms / steps, "value": world * steps / (ms / 1e3), "launches": int(s.stats().kernel_launches - l0),
               "iterations_mean": float(np.mean(its[-steps:])), "iterations_last": int(its[-1]),
               "relres": float(s.stats().relres), "next_k": warmup + steps}
        if sampler:
            res["clocks"] = sampler.stop()
        return res

    # ---- the headline leg: the colony named by --colony (moving by default), library-default starting guess -------------
    g = make_solver()
    main_leg = resident_leg(g, recs, args.warmup, args.steps, sample=True)
    ms_step, value = main_leg["ms_per_step"], main_leg["value"]
    kpos = main_leg["next_k"]
    warm_mode = int(g.warm_mode())

    # ---- true residual of one more step, through the verification hooks (outside every timed region) ---------------------
    true_relres = None
    if hasattr(g, "build_rhs") and not tensor_feed and NW * NH <= 2048 * 2048:
        try:
            if len(recs) > 1:
                g.upload_cells_device(rec_dev.data_ptr() + pingpong(kpos, len(recs)) * stride, ncells, NPM)
            g.gather_resident()
            g.scatter_resident()
            u0 = g.get_field()
            g.step()
            kpos += 1
            u1 = g.get_field()
            b = g.build_rhs(u0)
            Au = g.apply_operator(u1, constrained=False)
            free = np.ones((NH, NW), dtype=bool)
            bt = kw.get("bc_type", (1, 1, 1, 1))
            if bt[0] in (1, 3): free[:, 0] = False
            if bt[1] in (1, 3): free[:, -1] = False
            if bt[2] in (1, 3): free[-1, :] = False
            if bt[3] in (1, 3): free[0, :] = False
            f = free.ravel()
            true_relres = float(np.linalg.norm((b - Au)[f]) / np.linalg.norm(b[f])) if not kw.get("channels") else None
        except Exception as e:   # the check must not take the measurement down
            true_relres = f"unavailable: {e}"

    # ---- end to end: per-step H2D of this step's records and deposits from pinned host memory, D2H of the samples --------
    amt_pin = torch.from_numpy(amount.copy()).pin_memory()
    out_pin = torch.zeros(ncells, dtype=torch.float64).pin_memory()
    dpp = lambda t, off=0: C.cast(t.data_ptr() + off, C.POINTER(C.c_double))
    L = E.lib()
    R = len(recs)

    def step_e2e(k):
        # fused drop-in: only per-cell data crosses PCIe (cell records in, sampled HSL out, deposits in, flux out)
        g._ck(L.eqgpu_cells_upload(g._h, dpp(rec_pin, pingpong(k, R) * stride), C.c_int64(ncells), C.c_double(NPM)))
        g._ck(L.eqgpu_cells_gather(g._h, dpp(out_pin)))
        g._ck(L.eqgpu_cells_scatter(g._h, dpp(amt_pin)))
        if tensor_feed:
            g.cells_tensor(1.5, 0.6)
        g.step()
        return g.stats().total_boundary_flux

    for i in range(3):
        step_e2e(kpos + i)
    kpos += 3
    ms_e2e = timed(step_e2e, kpos, args.steps) / args.steps
    kpos += args.steps

    side = {}
    ms_compat = None
    if not args.no_side_legs:
        # compat leg: the strict fenicsInterface contract -- the field lives on the host, each step the host adds this
        # step's deposits to solution_vector (writeHSL on the CPU, outside the timed region) and hands the whole vector
        # in; the solution comes back in the same buffer.  The deposits of four successive record sets are rasterised
        # once, by the device's own scatter kernel, before the leg.
        fld_pin = torch.zeros(N, dtype=torch.float64).pin_memory()
        cur = g.get_field()
        deps = []
        for q in range(4 if R > 1 else 1):
            if R > 1:
                g.upload_cells(recs[pingpong(kpos + q, R)], NPM)
            g.set_field(np.zeros(N))
            g.scatter(amount)
            deps.append(torch.from_numpy(g.get_field()))
        g.set_field(cur)
        fld_pin.copy_(torch.from_numpy(cur))
        kc = max(3, args.steps // 5)
        ms_compat = 0.0
        for it in range(3 + kc):
            fld_pin.add_(deps[pingpong(it, len(deps))])
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            g.step_host_ptr(fld_pin.data_ptr())
            e1.record(stream)
            torch.cuda.synchronize()
            if it >= 3:
                ms_compat += e0.elapsed_time(e1) / kc
        if world > 1:
            t = torch.tensor([ms_compat], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_compat = float(t.item())
        # the same workload with a colony that never changes (round 1's headline) and from a cold start
        if args.colony != "static":
            s2 = make_solver()
            side["static"] = resident_leg(s2, recs[:1], args.warmup, args.steps)
            s2.close()
        s3 = make_solver(0)
        side["cold"] = resident_leg(s3, recs, args.warmup, min(args.steps, 30))
        s3.close()

    slab = None
    if world > 1 and not args.no_slab and args.config == 3:
        try:
            slab = run_slab_leg(args, rank, local_rank, world, stream, nccl_ids)
        except Exception as e:
            slab = {"error": str(e)}
    elif world == 1 and not args.no_slab and not args.no_side_legs and args.config == 3:
        try:
            slab = run_slab_baseline(args, local_rank, stream)
        except Exception as e:
            slab = {"error": str(e)}

    if rank == 0:
        peak, peak_src = peaks()
        # level-0 kernels timed alone with CUDA events on the launching stream, L2 flushed before every launch
        roof = {}
        for name in ("presmooth", "postsmooth", "apply_p", "update_r", "update_xr"):
            try:
                kms, kbytes = g.bench_kernel(name, 30)
            except E.EqGpuError:
                continue
            roof[name] = {"ms": kms, "bytes": kbytes, "gbs": kbytes / (kms * 1e-3) / 1e9}
        # the dominant kernel of the step as it runs: k_update_xr (x and r in one pass) is timed for reference only -- the
        # single-GPU path splits it into k_update_r + the deferred k_update_x
        rt_on = hasattr(g, "path") and g.path().get("register_tile_levels", 0) >= 1

        def kernel_name(k):   # the level-0 smoothers are the register-tile kernels when the solver selected them
            return {"presmooth": "k_pre_rt", "postsmooth": "k_post_rt"}.get(k, "k_" + k) if rt_on else "k_" + k

        on_path = [k for k in roof if k != "update_xr"] or list(roof)
        dom = max(on_path, key=lambda k: roof[k]["ms"])
        iters_mean = main_leg["iterations_mean"]
        # once-per-step bytes of the passes that actually ran: with the adaptive history depth a moving colony walks the
        # previous solution only (last_guess 2 = previous solution), whatever the configured mode
        eff_mode = 1 if (warm_mode in (2, 3, 4, 5, 6) and int(g.last_guess()) == 2) else warm_mode
        fixed = float(FIXED_BYTES.get(eff_mode, 224))
        per_it = 124.0 + 44.0 / 3.0
        bts = N * (fixed + per_it * iters_mean)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "colony": args.colony, "rods": ncells,
                       "colony_note": "rods advance 0.05-0.2 node per step along their axis and turn (+ exponential growth, "
                                      "ratchet and division with `growing`); every step uses that step's records"
                                      if args.colony != "static" else "records never change",
                       "pcg_iterations": main_leg["iterations_last"], "pcg_iterations_mean": iters_mean,
                       "relres": main_leg["relres"], "true_relres_next_step": true_relres, "rtol": 1e-12,
                       "mg_levels": int(g.stats().levels), "parallelism": f"layer-per-gpu x{world}",
                       "l2": "working set (6 fine fp64 vectors = 201 MB + MG hierarchy) exceeds the 126 MB L2; no explicit flush "
                             "in the step legs; the isolated-kernel roofline timings flush L2 before every launch",
                       "warm_mode": warm_mode, "last_guess": int(g.last_guess()),
                       "dof_updates_per_sec": value * N},
            "clocks": main_leg.get("clocks"),
            "e2e": {"value": world * 1e3 / ms_e2e, "unit": UNIT,
                    "h2d_bytes_per_step": int(ncells * 16 * 8 + amt_pin.numel() * 8),
                    "d2h_bytes_per_step": int(out_pin.numel() * 8 + 8),
                    "path": "eqgpu_cells_upload (this step's records) + gather + scatter + step with pinned host buffers "
                            "(field stays in HBM)"},
            "gpu_launches": main_leg["launches"],
            "roofline": {"bound": "hbm", "kernel": f"{kernel_name(dom)} (level 0, {NW}x{NH})", "achieved": roof[dom]["gbs"],
                         "peak": peak, "unit": "GB/s", "frac": roof[dom]["gbs"] / peak,
                         "traffic": ncu_traffic(dom) if (NW, NH) == (2048, 2048) else None,
                         "traffic_note": "STATIC: bytes/launch, dram__bytes_read.sum + dram__bytes_write.sum from the committed "
                                         "ncu --set full capture in profiles/, not measured in this run",
                         "timing": "each launch bracketed by CUDA events on the launching stream, L2 flushed "
                                   "(256 MB memset) before every launch",
                         "peak_source": peak_src, "kernels": roof,
                         "step_ideal_frac": (16.0 * N / (ms_step * 1e-3) / 1e9) / peak,
                         # whole step, algorithmic bytes of DESIGN.md section 5: per PCG iteration 124 B/DOF on level 0
                         # (presmooth 18, postsmooth 26, apply_p 32, update_r 24, update_x 24) + 44/3 on the coarser
                         # levels; per step the starting-guess and step-tail passes (`fixed`, by warm mode)
                         "step_algorithmic": {"bytes": bts, "gbs": bts / (ms_step * 1e-3) / 1e9,
                                              "frac": bts / (ms_step * 1e-3) / 1e9 / peak,
                                              "formula": f"N*({fixed:.0f} + (124 + 44/3)*mean_iterations)"}},
        }
        if ms_compat is not None:
            line["e2e_compat"] = {"value": world * 1e3 / ms_compat, "unit": UNIT,
                                  "h2d_bytes_per_step": N * 8, "d2h_bytes_per_step": N * 8,
                                  "path": "eqgpu_step_host: fenicsInterface::stepDiffusion contract, full solution_vector in/out"}
        if "static" in side:
            line["value_static"] = side["static"]["value"]
            line["config"]["static"] = {k: side["static"][k] for k in ("ms_per_step", "iterations_mean", "launches")}
        if "cold" in side:
            line["value_cold"] = side["cold"]["value"]
            line["config"]["cold"] = {k: side["cold"][k] for k in ("ms_per_step", "iterations_mean", "launches")}
        line["value_" + args.colony] = value
        if slab is not None:
            line["slab"] = slab
        if not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_subprocess(args.colony)
        print(json.dumps(line), flush=True)
    g.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
