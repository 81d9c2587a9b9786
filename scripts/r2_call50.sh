#!/bin/bash
# last knob sweep on one GPU (moving colony, 30 steps after 5)
mkdir -p gpurun_out/c50
cd /root/repo
run() { name=$1; shift; env "$@" timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-side-legs > gpurun_out/c50/bench_$name.json 2> gpurun_out/c50/bench_$name.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/c50/bench_$name.json").read().strip().splitlines()[-1])
    print("$name", "value %.1f ms %.4f it %.2f relres %.1e" % (d["value"], d["ms_per_step"], d["config"]["pcg_iterations_mean"], d["config"]["relres"]))
except Exception as e:
    print("$name unreadable", e)
PY
}
run base
run nuc3 EQGPU_NUC=3
run lo02 EQGPU_CHEB_LO=0.2
run lo025_nuc3 EQGPU_CHEB_LO=0.25 EQGPU_NUC=3
run ctarget30 EQGPU_COARSE_TARGET=30
run base2
