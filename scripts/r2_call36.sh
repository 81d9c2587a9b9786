#!/bin/bash
# full GPU suite + smoke + default N=1 bench on the final code of the session
mkdir -p gpurun_out/c36
cd /root/repo
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/c36/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/c36/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/c36/bench_n1_20.json 2> gpurun_out/c36/bench_n1_20.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/c36/bench_n1_20.json").read().strip().splitlines()[-1])
r=d["roofline"]
print("value", d["value"], "ms", d["ms_per_step"], "it", d["config"]["pcg_iterations_mean"], "static", d.get("value_static"), "cold", d.get("value_cold"), "e2e", d["e2e"]["value"], "compat", d["e2e_compat"]["value"])
print("roofline", r["kernel"], r["frac"], "step", r["step_algorithmic"]["frac"], "cpu", d["cpu_baseline"]["value"], d.get("slab",{}).get("value"))
PY
