#!/bin/bash
mkdir -p gpurun_out/c22
cd /root/repo
timeout 200 python bench.py --steps 10 --warmup 2 --no-cpu-baseline --no-side-legs > gpurun_out/c22/bench_kt.json 2> gpurun_out/c22/bench_kt.err
grep ktrace gpurun_out/c22/bench_kt.err
