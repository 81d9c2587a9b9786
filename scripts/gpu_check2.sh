#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out; O=gpurun_out
date +%s > $O/t0
timeout 600 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider > $O/pytest_gpu.log 2>&1
echo "pytest rc=$? elapsed $(( $(date +%s) - $(cat $O/t0) )) s" >> $O/pytest_gpu.log
tail -6 $O/pytest_gpu.log
python bench.py --config 7 --steps 30 --warmup 5 --no-cpu-baseline > $O/bench_tensor.json 2> $O/bench_tensor.err
EQGPU_TENSOR_COLD=1 python bench.py --config 7 --steps 20 --warmup 3 --no-cpu-baseline > $O/bench_tensor_cold.json 2> $O/bench_tensor_cold.err
python bench.py --no-cpu-baseline > $O/bench_default2.json 2> $O/bench_default2.err
EQGPU_LS_DEBUG=1 python scripts/ls_debug.py 257 40 2> $O/dbg_257_default.log
grep -h "^iterations" $O/dbg_257_default.log
cut -c1-330 $O/bench_tensor.json; echo; cut -c1-330 $O/bench_tensor_cold.json; echo; cut -c1-330 $O/bench_default2.json; echo
echo "total elapsed $(( $(date +%s) - $(cat $O/t0) )) s"
