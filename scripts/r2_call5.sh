#!/bin/bash
mkdir -p gpurun_out/c5
cd /root/repo
timeout 300 python -m pytest tests/test_stream_gpu.py -x -q > gpurun_out/c5/pytest_stream.log 2>&1; echo "stream rc=$?"
tail -5 gpurun_out/c5/pytest_stream.log
run() { name=$1; shift; env "$@" EQGPU_WARM=1 timeout 120 python bench.py --steps 40 --warmup 6 --no-cpu-baseline --no-side-legs > gpurun_out/c5/bench_$name.json 2> gpurun_out/c5/bench_$name.err; }
run tile EQGPU_STREAM=0
run stream_all EQGPU_STREAM=1
run stream_l0 EQGPU_STREAM=1 EQGPU_STREAM_MIN=2000000
run stream_l01 EQGPU_STREAM=1 EQGPU_STREAM_MIN=1000000
run stream_l0_nouni EQGPU_STREAM=1 EQGPU_STREAM_MIN=2000000 EQGPU_STREAM_UNI=0
run stream_l0_noapply EQGPU_STREAM=1 EQGPU_STREAM_MIN=2000000 EQGPU_STREAM_APPLY=0
EQGPU_TRACE=1 EQGPU_WARM=1 timeout 120 python bench.py --steps 10 --warmup 4 --no-cpu-baseline --no-side-legs > gpurun_out/c5/trace.json 2> gpurun_out/c5/trace.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/c5/bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        k=d["roofline"]["kernels"]
        print(f.split("/")[-1], "value %.1f ms %.3f it %.2f mode %s true_relres %s | pre %.1f us post %.1f us apply_p %.1f" % (d["value"], d["ms_per_step"], d["config"]["pcg_iterations_mean"], d["config"]["warm_mode"], d["config"].get("true_relres_next_step"), k["presmooth"]["ms"]*1e3, k["postsmooth"]["ms"]*1e3, k["apply_p"]["ms"]*1e3))
    except Exception as e:
        print(f, "unreadable", e)
PY
head -14 gpurun_out/c5/trace.err
EQGPU_WARM=1 timeout 600 ncu --set full --import-source on --clock-control none -k regex:ks_ -s 22 -c 12 -o gpurun_out/c5/ks python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-side-legs > gpurun_out/c5/ncu_full.log 2>&1; echo "ncu rc=$?"
