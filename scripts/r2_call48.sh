#!/bin/bash
# deferred x update on slabs: how many CTAs (it runs beside the coarse levels and must not crowd them out)
mkdir -p gpurun_out/c48
cd /root/repo
export EQGPU_PEER_TIMEOUT_MS=5000
run() { name=$1; shift; env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --mode slab --steps 16 --warmup 4 > gpurun_out/c48/slab_$name.json 2> gpurun_out/c48/slab_$name.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/c48/slab_$name.json").read().strip().splitlines()[-1])
    print("$name", "value %.2f ms %.3f it %.2f" % (d["value"], d["ms_per_step"], d["pcg_iterations_mean"]), d.get("skipped"))
except Exception as e:
    print("$name unreadable", e)
PY
}
run x74
run x37 EQGPU_XUPD_BLOCKS=37
run x148 EQGPU_XUPD_BLOCKS=148
run x296 EQGPU_XUPD_BLOCKS=296
