#!/bin/bash
# in-situ timeline of two slab iterations (events between all launches), peer and NCCL transports, 2 GPUs
mkdir -p gpurun_out/c31
cd /root/repo
export EQGPU_PEER_TIMEOUT_MS=5000 EQGPU_TRACE=1
run() { name=$1; shift; env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --mode slab --steps 6 --warmup 3 > gpurun_out/c31/slab_$name.json 2> gpurun_out/c31/slab_$name.err; grep "^trace rank 0" gpurun_out/c31/slab_$name.err | head -80; }
run peer EQGPU_SLAB_PEER=1
run nccl EQGPU_SLAB_PEER=0
