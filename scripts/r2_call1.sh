#!/bin/bash
# round 2, GPU call 1: state of the suite after the bench/ABI changes, the moving-colony baseline for every starting
# guess, in-situ kernel times, launch list
mkdir -p gpurun_out/c1
cd /root/repo
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/c1/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c1/pytest.log
tail -3 gpurun_out/c1/pytest.log
timeout 300 python bench.py --steps 60 --warmup 10 > gpurun_out/c1/bench_default.json 2> gpurun_out/c1/bench_default.err; echo "bench rc=$?"
for m in 0 1 3 6 7; do
  EQGPU_WARM=$m timeout 120 python bench.py --steps 60 --warmup 10 --no-cpu-baseline --no-side-legs > gpurun_out/c1/bench_warm$m.json 2> gpurun_out/c1/bench_warm$m.err
done
EQGPU_WARM=6 timeout 120 python bench.py --steps 60 --warmup 10 --no-cpu-baseline --no-side-legs --colony growing > gpurun_out/c1/bench_growing.json 2>&1
EQGPU_TRACE=1 EQGPU_WARM=1 timeout 120 python bench.py --steps 10 --warmup 4 --no-cpu-baseline --no-side-legs > gpurun_out/c1/trace.json 2> gpurun_out/c1/trace.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 300 --csv --log-file gpurun_out/c1/launches.csv python bench.py --steps 6 --warmup 4 --no-cpu-baseline --no-side-legs > gpurun_out/c1/ncu_bench.log 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/c1/smi.txt
nproc >> gpurun_out/c1/smi.txt; free -g >> gpurun_out/c1/smi.txt
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/c1/bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "value %.1f ms %.3f it %.2f e2e %.1f static %s cold %s true_relres %s" % (d["value"], d["ms_per_step"], d["config"]["pcg_iterations_mean"], d["e2e"]["value"], d.get("value_static"), d.get("value_cold"), d["config"].get("true_relres_next_step")))
    except Exception as e:
        print(f, "unreadable", e)
PY
