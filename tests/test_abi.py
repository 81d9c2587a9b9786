"""No-GPU checks of the C-ABI: the library loads, exports every symbol include/eqgpu.h declares, agrees with
the ctypes mirror on struct layout, and refuses to run without a CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "eqgpu.h")).read()
    return sorted(set(re.findall(r"EQGPU_API[^;(]*?\b(eqgpu_\w+)\s*\(", txt)))


def test_header_declares_the_expected_surface():
    syms = header_symbols()
    for must in ("eqgpu_create", "eqgpu_destroy", "eqgpu_step", "eqgpu_step_host", "eqgpu_cells_gather",
                 "eqgpu_cells_scatter", "eqgpu_cells_raster", "eqgpu_set_tensor", "eqgpu_get_channels"):
        assert must in syms
    assert len(syms) >= 25


def test_library_exports_every_declared_symbol():
    import eq_b200 as E
    if not os.path.exists(E.LIB_PATH):
        E.build()
    out = subprocess.run(["nm", "-D", "--defined-only", E.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (eqgpu_\w+)", out))
    missing = [s for s in header_symbols() if s not in exported]
    assert not missing, missing
    # nothing but the ABI leaks out of the library
    leaked = [l for l in out.splitlines() if " T " in l and "eqgpu_" not in l]
    assert not leaked, leaked[:5]
    L = E.lib()
    for s in header_symbols():
        assert hasattr(L, s)
    assert sorted(E.API_SYMBOLS) == header_symbols()


def test_struct_layout_matches_header(tmp_path):
    import eq_b200 as E
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "eqgpu.h"\nint main(){printf("%zu %zu %zu %zu %zu\\n",'
                   'sizeof(eqgpu_params),sizeof(eqgpu_stats),offsetof(eqgpu_params,stream),'
                   'offsetof(eqgpu_params,rtol),offsetof(eqgpu_stats,kernel_launches));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.run(["/usr/bin/gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    a, b, c, d, e = map(int, subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split())
    assert a == C.sizeof(E.Params) and b == C.sizeof(E.Stats)
    assert c == E.Params.stream.offset and d == E.Params.rtol.offset and e == E.Stats.kernel_launches.offset


def test_default_params_are_the_shipped_problem():
    import eq_b200 as E
    p = E.default_params()
    assert (p.nW, p.nH, p.hx, p.dt, p.D) == (201, 41, 0.5, 0.1, 1200.0)
    assert list(p.bc_type) == [1, 1, 1, 1] and p.channel_iters == 48 and p.rtol == 1e-12


def test_no_gpu_means_loud_failure_not_cpu_fallback():
    import eq_b200 as E
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("GPU present")
    with pytest.raises(E.EqGpuError) as ei:
        E.GpuHSL(33, 17)
    assert "no CPU fallback" in str(ei.value)


def test_invalid_configurations_are_rejected_before_touching_the_device():
    import eq_b200 as E
    L = E.lib()
    h = C.c_void_p()
    p = E.default_params()
    p.nW = 2
    assert L.eqgpu_create(C.byref(p), C.byref(h)) == -1
    p = E.default_params()
    p.bc_type[E.TOP] = E.ROBIN          # Robin exists on left/right only (fenics/hslD.ufl:39-42)
    assert L.eqgpu_create(C.byref(p), C.byref(h)) == -1
    p = E.default_params()
    p.abi_version = 99
    assert L.eqgpu_create(C.byref(p), C.byref(h)) == -1
    assert b"ABI" in L.eqgpu_last_error(None)


def test_product_package_does_not_import_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "eq_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "eq_oracle" not in txt and "from oracle" not in txt and "import oracle" not in txt, f
