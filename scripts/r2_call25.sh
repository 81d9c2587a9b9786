#!/bin/bash
mkdir -p gpurun_out/c25
cd /root/repo
timeout 600 python -m pytest tests/test_rt_gpu.py -x -q > gpurun_out/c25/pytest_rt.log 2>&1; echo "rt rc=$?"
tail -3 gpurun_out/c25/pytest_rt.log
EQGPU_RT_LEAN=1 timeout 600 ncu --set full --clock-control none -k regex:'k_pre_rt3|k_post_rt3' -s 8 -c 2 -o gpurun_out/c25/rt3 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-side-legs > gpurun_out/c25/ncu_full.log 2>&1; echo "ncu rc=$?"
