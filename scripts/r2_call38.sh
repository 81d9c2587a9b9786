#!/bin/bash
# fused p-update + tensor apply (k_apply_p_asm): tensor parity tests, config 7 with and without it
mkdir -p gpurun_out/c38
cd /root/repo
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_parity_large_gpu.py -x -q -k "tensor" > gpurun_out/c38/pytest_tensor.log 2>&1; echo "tensor pytest rc=$?"; tail -2 gpurun_out/c38/pytest_tensor.log
run() { name=$1; shift; env "$@" timeout 300 python bench.py --config 7 --steps 20 --warmup 5 --no-cpu-baseline --no-side-legs > gpurun_out/c38/bench_$name.json 2> gpurun_out/c38/bench_$name.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/c38/bench_$name.json").read().strip().splitlines()[-1])
    print("$name", "value %.1f ms %.3f it %.2f relres %s" % (d["value"], d["ms_per_step"], d["config"]["pcg_iterations_mean"], d["config"].get("relres")))
except Exception as e:
    print("$name unreadable", e)
PY
}
run fused_p
run unfused_p EQGPU_TENSOR_FUSE_P=0
