#!/bin/bash
mkdir -p gpurun_out/c41
cd /root/repo
export EQGPU_PEER_TIMEOUT_MS=5000 EQGPU_TRACE=1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --mode slab --steps 6 --warmup 3 > gpurun_out/c41/slab_trace.json 2> gpurun_out/c41/slab_trace.err
grep "^trace rank 0" gpurun_out/c41/slab_trace.err | sed -n 29,56p
