#!/bin/bash
mkdir -p gpurun_out/c17
cd /root/repo
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/c17/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/c17/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c17/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/c17/smoke.log
