#!/bin/bash
# the other bench configurations on the round's final code (one GPU)
mkdir -p gpurun_out/c51
cd /root/repo
for c in 2 4 6 7; do
  timeout 120 python bench.py --config $c --steps 20 --warmup 5 --no-cpu-baseline --no-side-legs > gpurun_out/c51/bench_config$c.json 2> gpurun_out/c51/bench_config$c.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/c51/bench_config$c.json").read().strip().splitlines()[-1])
    print("config $c", d["metric"], "value %.1f ms %.3f it %.2f" % (d["value"], d["ms_per_step"], d["config"]["pcg_iterations_mean"]), d["config"]["workload"][:90])
except Exception as e:
    print("config $c unreadable", e)
PY
done
