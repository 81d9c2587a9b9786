#!/bin/bash
# exchange-kernel grid size sweep on 2 GPUs + 4-rank parity test (skips on 2 GPUs)
mkdir -p gpurun_out/c39
cd /root/repo
export EQGPU_PEER_TIMEOUT_MS=5000
run() { name=$1; shift; env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --mode slab --steps 12 --warmup 3 > gpurun_out/c39/slab_$name.json 2> gpurun_out/c39/slab_$name.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/c39/slab_$name.json").read().strip().splitlines()[-1])
    print("$name", "value %.2f ms %.3f it %.1f" % (d["value"], d["ms_per_step"], d["pcg_iterations_mean"]), d.get("skipped"))
except Exception as e:
    print("$name unreadable", e)
PY
}
run b64_p4096
run b128_p2048 EQGPU_PEER_BLOCKS=128 EQGPU_PEER_PER_BLOCK=2048
run b32_p8192 EQGPU_PEER_BLOCKS=32 EQGPU_PEER_PER_BLOCK=8192
run b148_p1024 EQGPU_PEER_BLOCKS=148 EQGPU_PEER_PER_BLOCK=1024
run b16_p16384 EQGPU_PEER_BLOCKS=16 EQGPU_PEER_PER_BLOCK=16384
