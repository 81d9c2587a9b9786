#!/bin/bash
# First GPU call of round 2 (gpurun --timeout 900 -- 'bash scripts/round2_first_call.sh'): everything round 1 wrote
# after its GPU budget ran out gets its first full run, then the mode-7 step is profiled.
#  1. the whole GPU suite (the reference-class golden tests and the mode-7 tests close it)
#  2. the default bench line (opts into mode 7 after its 30-step check), and the same with mode 6 / depth 5 forced
#  3. ncu launch list of a mode-7 bench run: where does the fixed part of the step go?
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out; O=gpurun_out
date +%s > $O/t0
python -m pytest tests -q -m gpu -x > $O/r2_pytest_gpu.log 2>&1; tail -3 $O/r2_pytest_gpu.log
python bench.py > $O/r2_bench_default.json 2> $O/r2_bench_default.err
EQGPU_WARM=6 python bench.py --no-cpu-baseline > $O/r2_bench_mode6.json 2> $O/r2_bench_mode6.err
EQGPU_WARM=7 EQGPU_RING_DEPTH=5 python bench.py --no-cpu-baseline > $O/r2_bench_mode7_depth5.json 2> $O/r2_bench_mode7_depth5.err
EQGPU_WARM=7 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 2000 -c 400 --csv \
    --log-file $O/r2_launches_mode7.csv python bench.py --steps 60 --warmup 40 --no-cpu-baseline > $O/r2_ncu_mode7.log 2>&1
for f in r2_bench_default r2_bench_mode6 r2_bench_mode7_depth5; do
    python - "$O/$f.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    c = d["config"]
    print(sys.argv[1], "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "iters_mean", c.get("pcg_iterations_mean"),
          "warm", c.get("warm_mode"), "ring_check", c.get("ring_check"), "clocks", d.get("clocks"))
except Exception as e:
    print(sys.argv[1], "unreadable:", e)
PY
done
python profiles/summarize_launches.py $O/r2_launches_mode7.csv 2>/dev/null | head -40
echo "total elapsed $(( $(date +%s) - $(cat $O/t0) )) s"
