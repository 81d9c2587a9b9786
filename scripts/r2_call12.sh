#!/bin/bash
mkdir -p gpurun_out/c12
cd /root/repo
timeout 600 ncu --set full --import-source on --clock-control none --cache-control none -k regex:'^k_presmooth|^k_postsmooth|^k_coarsest' -s 40 -c 9 -o gpurun_out/c12/small python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-side-legs > gpurun_out/c12/ncu_full.log 2>&1; echo "ncu rc=$?"
