#!/bin/bash
mkdir -p gpurun_out/c42
cd /root/repo
export EQGPU_PEER_TIMEOUT_MS=5000
timeout 600 python -m pytest tests/test_slab_gpu.py -x -q 2>&1 | tail -15 | tee gpurun_out/c42/pytest_slab4.log
