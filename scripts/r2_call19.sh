#!/bin/bash
mkdir -p gpurun_out/c19
cd /root/repo
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "tensor" > gpurun_out/c19/pytest_tensor.log 2>&1; echo "tensor pytest rc=$?"
tail -3 gpurun_out/c19/pytest_tensor.log
run() { name=$1; shift; env "$@" timeout 300 python bench.py --config 7 --steps 20 --warmup 5 --no-cpu-baseline --no-side-legs > gpurun_out/c19/bench_$name.json 2> gpurun_out/c19/bench_$name.err; }
run tensor_asm
run tensor_tprec EQGPU_TENSOR_PRECOND=tensor
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/c19/bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "value %.1f ms %.3f it %.2f relres %s" % (d["value"], d["ms_per_step"], d["config"]["pcg_iterations_mean"], d["config"].get("relres")))
    except Exception as e:
        print(f, "unreadable", e)
PY
tail -3 gpurun_out/c19/bench_tensor_asm.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 200 --csv --log-file gpurun_out/c19/launches_tensor.csv python bench.py --config 7 --steps 2 --warmup 2 --no-cpu-baseline --no-side-legs > gpurun_out/c19/ncu_bench.log 2>&1; echo "ncu list rc=$?"
