#!/bin/bash
mkdir -p gpurun_out/c16
cd /root/repo
timeout 600 python -m pytest tests/test_slab_gpu.py tests/test_layers_gloo.py -x -q > gpurun_out/c16/pytest_slab.log 2>&1; echo "slab pytest rc=$?"
tail -5 gpurun_out/c16/pytest_slab.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/c16/bench_n2.json 2> gpurun_out/c16/bench_n2.err; echo "bench n2 rc=$?"
tail -c 3000 gpurun_out/c16/bench_n2.json
tail -5 gpurun_out/c16/bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/c16/bench_ref_n2.json 2> gpurun_out/c16/bench_ref_n2.err; echo "ref n2 rc=$?"
tail -c 1500 gpurun_out/c16/bench_ref_n2.json
