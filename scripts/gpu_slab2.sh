#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out; O=gpurun_out
timeout 300 python -m pytest tests/test_slab_gpu.py tests/test_layers_gloo.py -q --timeout 200 -p no:cacheprovider > $O/pytest_slab2.log 2>&1
echo "rc=$?" >> $O/pytest_slab2.log; tail -5 $O/pytest_slab2.log
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --mode slab --slab-cols 4096 --steps 20 --warmup 5 > $O/bench_slab2.json 2> $O/bench_slab2.err
cut -c1-400 $O/bench_slab2.json; tail -3 $O/bench_slab2.err
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 50 --warmup 10 --no-cpu-baseline > $O/bench_layers2.json 2> $O/bench_layers2.err
cut -c1-300 $O/bench_layers2.json
