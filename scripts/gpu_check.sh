#!/bin/bash
# One gpurun call: GPU parity suite, bench lines (warm modes, widened rows), smoke, ncu launch list.
# Everything lands in gpurun_out/ (merged back by gpurun).
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt 2>&1
date +%s > $O/t0
timeout 600 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider > $O/pytest_gpu.log 2>&1
echo "pytest rc=$? elapsed $(( $(date +%s) - $(cat $O/t0) )) s" >> $O/pytest_gpu.log
tail -30 $O/pytest_gpu.log
python bench.py > $O/bench_w4.json 2> $O/bench_w4.err
EQGPU_WARM=3 python bench.py --no-cpu-baseline > $O/bench_w3.json 2> $O/bench_w3.err
cat $O/bench_w4.json $O/bench_w3.json | cut -c1-700
python bench.py --config 6 --steps 50 > $O/bench_fd.json 2> $O/bench_fd.err
python bench.py --config 7 --steps 20 --warmup 3 --no-cpu-baseline > $O/bench_tensor.json 2> $O/bench_tensor.err
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1
tail -2 $O/smoke.log
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 400 --csv --log-file $O/launches_r1c.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu_bench.log 2>&1
echo "total elapsed $(( $(date +%s) - $(cat $O/t0) )) s"
