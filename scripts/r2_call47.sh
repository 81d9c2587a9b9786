#!/bin/bash
# N = 8 (or 4): the driver's scaling run -- layers leg + slab leg (16384 x 2048 N)
N=${1:-8}
mkdir -p gpurun_out/c47
cd /root/repo
EQGPU_PEER_TIMEOUT_MS=8000 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/c47/bench_n$N.json 2> gpurun_out/c47/bench_n$N.err; echo "bench n$N rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/c47/bench_n$N.json").read().strip().splitlines()[-1])
print("n$N value", d["value"], "ms", d["ms_per_step"], "slab", json.dumps(d.get("slab"))[:1500])
PY
grep -v "^\*\|OMP_NUM" gpurun_out/c47/bench_n$N.err | tail -3
