#!/bin/bash
mkdir -p gpurun_out/c24
cd /root/repo
timeout 600 python -m pytest tests/test_rt_gpu.py -x -q > gpurun_out/c24/pytest_rt.log 2>&1; echo "rt rc=$?"
tail -5 gpurun_out/c24/pytest_rt.log
run() { name=$1; shift; env "$@" timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-side-legs > gpurun_out/c24/bench_$name.json 2> gpurun_out/c24/bench_$name.err; }
run lean1
run lean0 EQGPU_RT_LEAN=0
run lean1_ctas592 EQGPU_RT_CTAS=400
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/c24/bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        k=d["roofline"]["kernels"]
        print(f.split("/")[-1], "value %.1f ms %.3f it %.2f true %s | pre %.1f us post %.1f us" % (d["value"], d["ms_per_step"], d["config"]["pcg_iterations_mean"], d["config"].get("true_relres_next_step"), k["presmooth"]["ms"]*1e3, k["postsmooth"]["ms"]*1e3))
    except Exception as e:
        print(f, "unreadable", e)
PY
tail -2 gpurun_out/c24/bench_lean1.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 150 --csv --log-file gpurun_out/c24/launches_warm.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-side-legs > gpurun_out/c24/ncu_bench.log 2>&1; echo "ncu list rc=$?"
grep -E "rt3|k_pre_rt|k_post_rt" gpurun_out/c24/launches_warm.csv | tail -6 | cut -d, -f5,9,15
