#!/bin/bash
mkdir -p gpurun_out/c26
cd /root/repo
# N = 1: default bench with the slab baseline (one GPU of the two)
CUDA_VISIBLE_DEVICES=0 timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/c26/bench_n1.json 2> gpurun_out/c26/bench_n1.err; echo "bench n1 rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/c26/bench_n1.json").read().strip().splitlines()[-1])
print("n1 value", d["value"], "slab", json.dumps(d.get("slab"))[:900])
PY
timeout 600 python -m pytest tests/test_slab_gpu.py -x -q 2>&1 | tail -2
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/c26/bench_n2.json 2> gpurun_out/c26/bench_n2.err; echo "bench n2 rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/c26/bench_n2.json").read().strip().splitlines()[-1])
print("n2 value", d["value"], "slab", json.dumps(d.get("slab"))[:1200])
PY
tail -3 gpurun_out/c26/bench_n2.err
