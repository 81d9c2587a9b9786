#!/bin/bash
mkdir -p gpurun_out/c23
cd /root/repo
SECONDS=0; timeout 1200 python bench.py > gpurun_out/c23/bench_default.json 2> gpurun_out/c23/bench_default.err; echo "bench rc=$?"
echo "wall ${SECONDS}s"
tail -c 2500 gpurun_out/c23/bench_default.json
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "nonconvergence" 2>&1 | tail -2
