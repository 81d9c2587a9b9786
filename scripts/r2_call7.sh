#!/bin/bash
mkdir -p gpurun_out/c7
cd /root/repo
timeout 300 python -m pytest tests/test_stream_gpu.py -x -q > gpurun_out/c7/pytest_stream.log 2>&1; echo "stream rc=$?"
tail -15 gpurun_out/c7/pytest_stream.log
run() { name=$1; shift; env "$@" EQGPU_WARM=1 timeout 120 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-side-legs > gpurun_out/c7/bench_$name.json 2> gpurun_out/c7/bench_$name.err; }
run pipe_l0 EQGPU_STREAM=1 EQGPU_STREAM_MIN=2000000 EQGPU_STREAM_APPLY=0
run pipe_l0_apply EQGPU_STREAM=1 EQGPU_STREAM_MIN=2000000
run tile EQGPU_STREAM=0
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/c7/bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        k=d["roofline"]["kernels"]
        print(f.split("/")[-1], "value %.1f ms %.3f it %.2f true %s | pre %.1f us post %.1f us apply_p %.1f" % (d["value"], d["ms_per_step"], d["config"]["pcg_iterations_mean"], d["config"].get("true_relres_next_step"), k["presmooth"]["ms"]*1e3, k["postsmooth"]["ms"]*1e3, k["apply_p"]["ms"]*1e3))
    except Exception as e:
        print(f, "unreadable", e)
PY
tail -3 gpurun_out/c7/bench_pipe_l0.err
EQGPU_WARM=1 EQGPU_STREAM=1 EQGPU_STREAM_MIN=2000000 EQGPU_STREAM_APPLY=0 timeout 600 ncu --set full --import-source on --clock-control none -k regex:kp_ -s 6 -c 4 -o gpurun_out/c7/kp python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-side-legs > gpurun_out/c7/ncu_full.log 2>&1; echo "ncu rc=$?"
