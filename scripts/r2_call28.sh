#!/bin/bash
mkdir -p gpurun_out/c28
cd /root/repo
run() { name=$1; shift; env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --mode slab --steps 12 --warmup 3 > gpurun_out/c28/slab_$name.json 2> gpurun_out/c28/slab_$name.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/c28/slab_$name.json").read().strip().splitlines()[-1])
    print("$name", "value %.2f ms %.3f it %.1f" % (d["value"], d["ms_per_step"], d["pcg_iterations_mean"]), d.get("skipped"))
except Exception as e:
    print("$name unreadable", e)
PY
}

run noxch EQGPU_DEBUG_SKIP_EXCHANGE=1
tail -3 gpurun_out/c28/slab_noxch.err
