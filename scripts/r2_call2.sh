#!/bin/bash
# round 2, GPU call 2: streaming smoothers -- parity against the tile kernels, suite, bench with both, traces
mkdir -p gpurun_out/c2
cd /root/repo
timeout 300 python -m pytest tests/test_stream_gpu.py -x -q > gpurun_out/c2/pytest_stream.log 2>&1; echo "stream rc=$?"
tail -15 gpurun_out/c2/pytest_stream.log
for st in 1 0; do
  EQGPU_STREAM=$st EQGPU_WARM=1 timeout 120 python bench.py --steps 40 --warmup 6 --no-cpu-baseline --no-side-legs > gpurun_out/c2/bench_stream$st.json 2> gpurun_out/c2/bench_stream$st.err
done
EQGPU_STREAM=1 EQGPU_NO_TMA=1 EQGPU_WARM=1 timeout 120 python bench.py --steps 40 --warmup 6 --no-cpu-baseline --no-side-legs > gpurun_out/c2/bench_stream_notma.json 2>&1
EQGPU_STREAM=1 EQGPU_STREAM_MIN=2000000 EQGPU_WARM=1 timeout 120 python bench.py --steps 40 --warmup 6 --no-cpu-baseline --no-side-legs > gpurun_out/c2/bench_stream_l0only.json 2>&1
EQGPU_TRACE=1 EQGPU_WARM=1 timeout 120 python bench.py --steps 10 --warmup 4 --no-cpu-baseline --no-side-legs > gpurun_out/c2/trace.json 2> gpurun_out/c2/trace.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/c2/bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        k=d["roofline"]["kernels"]
        print(f.split("/")[-1], "value %.1f ms %.3f it %.2f true_relres %s | pre %.1f us post %.1f us apply_p %.1f" % (d["value"], d["ms_per_step"], d["config"]["pcg_iterations_mean"], d["config"].get("true_relres_next_step"), k["presmooth"]["ms"]*1e3, k["postsmooth"]["ms"]*1e3, k["apply_p"]["ms"]*1e3))
    except Exception as e:
        print(f, "unreadable", e)
PY
head -30 gpurun_out/c2/trace.err
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c2/pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/c2/pytest.log
