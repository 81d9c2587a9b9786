#!/bin/bash
# One gpurun call: GPU parity suite (default and with the least-squares guess forced on), bench lines
# (headline, widened rows), smoke, ncu launch list, small-mesh warm-start comparison.  Output: gpurun_out/.
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt 2>&1
date +%s > $O/t0
timeout 600 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider > $O/pytest_gpu.log 2>&1
echo "pytest rc=$? elapsed $(( $(date +%s) - $(cat $O/t0) )) s" >> $O/pytest_gpu.log
tail -5 $O/pytest_gpu.log
EQGPU_WARM=4 timeout 600 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider > $O/pytest_gpu_w4.log 2>&1
echo "pytest(EQGPU_WARM=4) rc=$?" >> $O/pytest_gpu_w4.log
tail -5 $O/pytest_gpu_w4.log
python bench.py > $O/bench_default.json 2> $O/bench_default.err
cut -c1-400 $O/bench_default.json
python bench.py --config 6 > $O/bench_fd.json 2> $O/bench_fd.err
python bench.py --config 7 --steps 20 --warmup 3 --no-cpu-baseline > $O/bench_tensor.json 2> $O/bench_tensor.err
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1
tail -2 $O/smoke.log
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 400 --csv --log-file $O/launches_r1c.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu_bench.log 2>&1
for w in 3 4; do
  EQGPU_LS_DEBUG=1 EQGPU_WARM=$w python scripts/ls_debug.py 257 40 2> $O/dbg_257_w$w.log
  EQGPU_LS_DEBUG=1 EQGPU_WARM=$w python scripts/ls_debug.py 512 40 2> $O/dbg_512_w$w.log
done
grep -h "^iterations" $O/dbg_257_w3.log $O/dbg_257_w4.log $O/dbg_512_w3.log $O/dbg_512_w4.log
echo "total elapsed $(( $(date +%s) - $(cat $O/t0) )) s"
