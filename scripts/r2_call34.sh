#!/bin/bash
# peer transport with interior ranks (two neighbours): 4 GPUs, slab leg with its parity self-check, peer and NCCL
mkdir -p gpurun_out/c34
cd /root/repo
export EQGPU_PEER_TIMEOUT_MS=5000
run() { name=$1; shift; env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --mode slab --steps 12 --warmup 3 > gpurun_out/c34/slab_$name.json 2> gpurun_out/c34/slab_$name.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/c34/slab_$name.json").read().strip().splitlines()[-1])
    print("$name", "value %.2f ms %.3f it %.1f relres %.2e" % (d["value"], d["ms_per_step"], d["pcg_iterations_mean"], d["relres"]), d.get("skipped"), d.get("parity"), {k: v for k, v in d["comm"].items() if k != "transport"})
except Exception as e:
    print("$name unreadable", e)
PY
tail -3 gpurun_out/c34/slab_$name.err | grep -v "^\*\*\*\|OMP_NUM"
}
run peer EQGPU_SLAB_PEER=1
run nccl EQGPU_SLAB_PEER=0
