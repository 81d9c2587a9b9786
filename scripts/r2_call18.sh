#!/bin/bash
mkdir -p gpurun_out/c18
cd /root/repo
timeout 600 python -m pytest tests/test_rt_gpu.py -x -q > gpurun_out/c18/pytest_rt.log 2>&1; echo "rt rc=$?"
tail -3 gpurun_out/c18/pytest_rt.log
run() { name=$1; shift; env "$@" timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-side-legs > gpurun_out/c18/bench_$name.json 2> gpurun_out/c18/bench_$name.err; }
run pair
run pair2
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/c18/bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        k=d["roofline"]["kernels"]
        print(f.split("/")[-1], "value %.1f ms %.3f it %.2f true %s | pre %.1f us post %.1f us apply_p %.1f" % (d["value"], d["ms_per_step"], d["config"]["pcg_iterations_mean"], d["config"].get("true_relres_next_step"), k["presmooth"]["ms"]*1e3, k["postsmooth"]["ms"]*1e3, k["apply_p"]["ms"]*1e3))
    except Exception as e:
        print(f, "unreadable", e)
PY
