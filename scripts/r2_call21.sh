#!/bin/bash
mkdir -p gpurun_out/c21
cd /root/repo
timeout 600 python -m pytest tests/test_rt_gpu.py -x -q > gpurun_out/c21/pytest_rt.log 2>&1; echo "rt rc=$?"
tail -12 gpurun_out/c21/pytest_rt.log
run() { name=$1; shift; env "$@" timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-side-legs > gpurun_out/c21/bench_$name.json 2> gpurun_out/c21/bench_$name.err; }
run base
run l01 EQGPU_RT_LEVELS=2
run l012 EQGPU_RT_LEVELS=3 EQGPU_RT_MIN_TILES=50
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/c21/bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        k=d["roofline"]["kernels"]
        print(f.split("/")[-1], "value %.1f ms %.3f it %.2f true %s | pre %.1f us post %.1f us" % (d["value"], d["ms_per_step"], d["config"]["pcg_iterations_mean"], d["config"].get("true_relres_next_step"), k["presmooth"]["ms"]*1e3, k["postsmooth"]["ms"]*1e3))
    except Exception as e:
        print(f, "unreadable", e)
PY
EQGPU_RT_LEVELS=3 EQGPU_RT_MIN_TILES=50 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 300 --csv --log-file gpurun_out/c21/launches_warm.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-side-legs > gpurun_out/c21/ncu_bench.log 2>&1; echo "ncu list rc=$?"
