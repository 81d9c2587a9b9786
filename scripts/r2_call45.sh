#!/bin/bash
# final evidence on one GPU: ncu launch list of the headline step, then the driver's bench command
mkdir -p gpurun_out/c45
cd /root/repo
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 300 --csv --log-file gpurun_out/c45/launches_final.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-side-legs > gpurun_out/c45/ncu_bench.log 2>&1; echo "ncu rc=$?"
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/c45/bench_n1_20.json 2> gpurun_out/c45/bench_n1_20.err; echo "bench rc=$?"
timeout 900 python bench.py > gpurun_out/c45/bench_n1_default.json 2> gpurun_out/c45/bench_n1_default.err; echo "bench default rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/c45/bench_reference.json 2> gpurun_out/c45/bench_reference.err; echo "reference rc=$?"
python - <<'PY'
import json
for f in ["bench_n1_20","bench_n1_default","bench_reference"]:
    try:
        d=json.loads(open(f"gpurun_out/c45/{f}.json").read().strip().splitlines()[-1])
        r=d.get("roofline",{})
        print(f, "value", d["value"], "ms", d["ms_per_step"], "it", d.get("config",{}).get("pcg_iterations_mean"), "static", d.get("value_static"), "e2e", d.get("e2e",{}).get("value"), "compat", d.get("e2e_compat",{}).get("value"), r.get("kernel"), r.get("frac"), r.get("step_algorithmic",{}).get("frac"))
    except Exception as e:
        print(f, "unreadable", e)
PY
