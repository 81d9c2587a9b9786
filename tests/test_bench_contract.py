"""bench.py's own arm end to end WITHOUT a GPU: the solver, torch.cuda's stream/event calls and pinned memory are
replaced by inert stand-ins so that every line of the measurement script runs on the CPU -- a NameError or a key
missing from the JSON line would otherwise first show up in the driver's round-end run.  Checks the contract keys
(metric, value, unit, n_gpus, steps, warmup, ms_per_step, higher_is_better, scaling, vs_baseline, dtype, data, config,
clocks, e2e with its byte counts, gpu_launches, roofline, cpu_baseline) and that the headline leg feeds every step its
own record set (the colony moves).  Nothing here measures anything."""
import importlib.util
import io
import json
import os
import sys
import types
from contextlib import redirect_stdout

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class _Stats:
    def __init__(self, g):
        self.iterations, self.kernel_launches, self.relres, self.levels = 2, g.launches, 3e-13, 6
        self.total_boundary_flux, self.steps = -1.0, g.steps


class _FakeSolver:
    """the slice of eq_b200.GpuHSL bench.py touches"""
    created = []

    def __init__(self, nW, nH, **kw):
        self.N, self.launches, self.steps, self.mode, self._h = nW * nH, 0, 0, 6, 0
        self.kw = kw
        self.device_uploads = 0
        _FakeSolver.created.append(self)

    def set_warm_start(self, mode): self.mode = mode
    def warm_mode(self): return self.mode
    def upload_cells(self, cells, npm): self.ncells = len(cells)
    def upload_cells_device(self, ptr, n, npm): self.device_uploads += 1
    def set_amounts(self, a): pass
    def gather_resident(self): self.launches += 1
    def scatter_resident(self): self.launches += 1
    def scatter(self, a): self.launches += 1
    def cells_tensor(self, dx, dy): self.launches += 1
    def step(self): self.launches += 20; self.steps += 1
    def step_host_ptr(self, ptr): self.step()
    def stats(self): return _Stats(self)
    def last_guess(self): return 7
    def get_field(self, out=None): return np.full(self.N, 1.0 + self.steps)
    def set_field(self, u): pass
    def build_rhs(self, u0): return np.ones(self.N)
    def apply_operator(self, x, constrained=False): return np.ones(self.N) * (1.0 - 1e-13)
    def bench_kernel(self, name, reps): return 0.05, 1.0e8
    def _ck(self, rc): assert rc == 0
    def close(self): pass


class _Event:
    def __init__(self, enable_timing=False): pass
    def record(self, stream=None): pass
    def elapsed_time(self, other): return 3.0


def run_bench(monkeypatch, argv, env=None):
    import torch
    fake = types.ModuleType("eq_b200")
    fake.__path__ = [os.path.join(ROOT, "eq_b200")]   # so that eq_b200.colony (host-side numpy) still imports
    fake.GpuHSL, fake.EqGpuError, fake.DISC_FD = _FakeSolver, RuntimeError, 1
    fake.lib = lambda: types.SimpleNamespace(eqgpu_cells_upload=lambda *a: 0, eqgpu_cells_gather=lambda *a: 0,
                                             eqgpu_cells_scatter=lambda *a: 0)
    monkeypatch.setitem(sys.modules, "eq_b200", fake)
    monkeypatch.delitem(sys.modules, "eq_b200.colony", raising=False)
    monkeypatch.setattr(torch.cuda, "set_device", lambda d: None)
    monkeypatch.setattr(torch.cuda, "current_stream", lambda: types.SimpleNamespace(cuda_stream=0))
    monkeypatch.setattr(torch.cuda, "synchronize", lambda: None)
    monkeypatch.setattr(torch.cuda, "Event", _Event)
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self: self)
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self: self)
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "EQGPU_WARM"):
        monkeypatch.delenv(k, raising=False)
    for k, v in (env or {}).items():
        monkeypatch.setenv(k, v)
    spec = importlib.util.spec_from_file_location("bench_under_test", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    bench.NW = bench.NH = 600
    bench.NCELLS = 40
    monkeypatch.setattr(bench, "cpu_baseline_subprocess",
                        lambda *a, **k: {"value": 0.01, "unit": "steps/s", "cores": 1, "kind": "port", "sample": "stub"})
    _FakeSolver.created = []
    monkeypatch.setattr(sys, "argv", ["bench.py"] + argv)
    out = io.StringIO()
    with redirect_stdout(out):
        bench.main()
    return json.loads(out.getvalue().strip().splitlines()[-1])


def test_bench_line_carries_the_contract_keys(monkeypatch):
    line = run_bench(monkeypatch, ["--steps", "4", "--warmup", "3"])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline",
                "value_static", "value_cold", "value_moving", "e2e_compat"):
        assert key in line, key
    assert line["n_gpus"] == 1 and line["steps"] == 4 and line["warmup"] == 3 and line["higher_is_better"] is True
    assert line["unit"] == "steps/s" and line["dtype"] == "f64" and line["scaling"] == "weak" and line["vs_baseline"] is None
    assert "workload" in line["config"] and "model" not in line["config"] and line["config"]["colony"] == "moving"
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(line["e2e"]) and line["e2e"]["h2d_bytes_per_step"] > 0
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(line["roofline"]) and line["roofline"]["bound"] == "hbm"
    assert {"value", "unit", "cores", "kind", "sample"} <= set(line["cpu_baseline"])
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(line["clocks"])
    assert line["gpu_launches"] == 4 * 22                        # counted from the solver's own launch counter
    assert line["value"] == line["value_moving"]                  # the headline is the colony that changes
    main = _FakeSolver.created[0]
    assert main.device_uploads >= 7                               # every resident step took that step's records
    assert line["config"]["warm_mode"] == 6 and "224" in line["roofline"]["step_algorithmic"]["formula"]
    assert abs(line["config"]["true_relres_next_step"] - 1e-13) < 1e-15


def test_bench_static_colony_and_the_widened_configs(monkeypatch):
    line = run_bench(monkeypatch, ["--steps", "3", "--warmup", "3", "--no-cpu-baseline", "--colony", "static"])
    assert line["config"]["colony"] == "static" and "value_static" in line and "cpu_baseline" not in line
    assert _FakeSolver.created[0].device_uploads == 0             # one record set, uploaded once
    for cfg in ("2", "6", "7"):
        line = run_bench(monkeypatch, ["--steps", "3", "--warmup", "3", "--no-cpu-baseline", "--no-side-legs", "--config", cfg])
        assert line["value"] > 0 and "e2e_compat" not in line


def test_pingpong_replay_never_jumps():
    spec = importlib.util.spec_from_file_location("bench_pp", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    seq = [bench.pingpong(k, 5) for k in range(20)]
    assert seq[:9] == [0, 1, 2, 3, 4, 3, 2, 1, 0] and all(abs(a - b) == 1 for a, b in zip(seq, seq[1:]))
    assert [bench.pingpong(k, 1) for k in range(3)] == [0, 0, 0]
