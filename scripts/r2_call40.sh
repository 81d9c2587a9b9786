#!/bin/bash
# exchange-kernel grid size sweep on 2 GPUs + 4-rank parity test (skips on 2 GPUs)
mkdir -p gpurun_out/c40
cd /root/repo
export EQGPU_PEER_TIMEOUT_MS=5000
run() { name=$1; shift; env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --mode slab --steps 12 --warmup 3 > gpurun_out/c40/slab_$name.json 2> gpurun_out/c40/slab_$name.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/c40/slab_$name.json").read().strip().splitlines()[-1])
    print("$name", "value %.2f ms %.3f it %.1f" % (d["value"], d["ms_per_step"], d["pcg_iterations_mean"]), d.get("skipped"))
except Exception as e:
    print("$name unreadable", e)
PY
}
run b148_p1024 EQGPU_PEER_BLOCKS=148 EQGPU_PEER_PER_BLOCK=1024
run b296_p512 EQGPU_PEER_BLOCKS=296 EQGPU_PEER_PER_BLOCK=512
run b296_p256 EQGPU_PEER_BLOCKS=296 EQGPU_PEER_PER_BLOCK=256
run b592_p256 EQGPU_PEER_BLOCKS=592 EQGPU_PEER_PER_BLOCK=256
run b148_p256 EQGPU_PEER_BLOCKS=148 EQGPU_PEER_PER_BLOCK=256
