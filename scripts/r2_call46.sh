#!/bin/bash
# final validation on one GPU: the driver's GPU tier (pytest -m gpu, smoke)
mkdir -p gpurun_out/c46
cd /root/repo
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/c46/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/c46/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
