#!/bin/bash
# adaptive history depth on slabs (2 GPUs): on / off, 20 steps
mkdir -p gpurun_out/c44
cd /root/repo
export EQGPU_PEER_TIMEOUT_MS=5000
run() { name=$1; shift; env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --mode slab --steps 20 --warmup 5 > gpurun_out/c44/slab_$name.json 2> gpurun_out/c44/slab_$name.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/c44/slab_$name.json").read().strip().splitlines()[-1])
    print("$name", "value %.2f ms %.3f it %.2f relres %.2e" % (d["value"], d["ms_per_step"], d["pcg_iterations_mean"], d["relres"]), d.get("skipped"), d["parity"]["rel_l2_vs_oracle_lu"])
except Exception as e:
    print("$name unreadable", e)
PY
}
run adaptive
run full EQGPU_SLAB_ADAPTIVE=0
