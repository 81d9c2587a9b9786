#!/bin/bash
mkdir -p gpurun_out/c20
cd /root/repo
run() { name=$1; shift; env "$@" timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-side-legs > gpurun_out/c20/bench_$name.json 2> gpurun_out/c20/bench_$name.err; }
run base
run nu0_4 EQGPU_NU0=4
run nu1_3 EQGPU_NU1=3
run nu0_4_nu1_3 EQGPU_NU0=4 EQGPU_NU1=3
run nu1_2 EQGPU_NU1=2
run cheb_lo3 EQGPU_CHEB_LO=0.3
run cheb_lo2 EQGPU_CHEB_LO=0.2
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/c20/bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        k=d["roofline"]["kernels"]
        print(f.split("/")[-1], "value %.1f ms %.3f it %.2f true %s" % (d["value"], d["ms_per_step"], d["config"]["pcg_iterations_mean"], d["config"].get("true_relres_next_step")))
    except Exception as e:
        print(f, "unreadable", e)
PY
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "tensor" 2>&1 | tail -2
