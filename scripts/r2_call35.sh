#!/bin/bash
# LL staging exchange + PDL on 2 GPUs: parity (both transports), timeline, then the slab leg
mkdir -p gpurun_out/c35
cd /root/repo
export EQGPU_PEER_TIMEOUT_MS=5000
timeout 600 python -m pytest tests/test_slab_gpu.py -x -q 2>&1 | tail -15 | tee gpurun_out/c35/pytest_slab.log
run() { name=$1; shift; env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --mode slab --steps 12 --warmup 3 > gpurun_out/c35/slab_$name.json 2> gpurun_out/c35/slab_$name.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/c35/slab_$name.json").read().strip().splitlines()[-1])
    print("$name", "value %.2f ms %.3f it %.1f" % (d["value"], d["ms_per_step"], d["pcg_iterations_mean"]), d.get("skipped"), d.get("parity"), {k: v for k, v in d["comm"].items() if k != "transport"})
except Exception as e:
    print("$name unreadable", e)
PY
tail -3 gpurun_out/c35/slab_$name.err | grep -v "^\*\*\*\|OMP_NUM"
}
run peer EQGPU_SLAB_PEER=1
run nodefer EQGPU_SLAB_PEER=1 EQGPU_SLAB_DEFER_X=0
run trace EQGPU_SLAB_PEER=1 EQGPU_TRACE=1
grep "^trace rank 0" gpurun_out/c35/slab_trace.err | sed -n 1,31p
