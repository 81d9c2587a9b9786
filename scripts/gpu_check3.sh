#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out; O=gpurun_out
date +%s > $O/t0
timeout 600 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider > $O/pytest_gpu.log 2>&1
echo "pytest rc=$? elapsed $(( $(date +%s) - $(cat $O/t0) )) s" >> $O/pytest_gpu.log
tail -6 $O/pytest_gpu.log
python bench.py --no-cpu-baseline > $O/bench_w6.json 2> $O/bench_w6.err
EQGPU_WARM=5 python bench.py --no-cpu-baseline > $O/bench_w5b.json 2> $O/bench_w5b.err
EQGPU_LS_DEBUG=1 python scripts/ls_debug.py 2048 60 2> $O/dbg_2048_w6.log
grep -h "^iterations" $O/dbg_2048_w6.log
cut -c1-330 $O/bench_w6.json; echo; cut -c1-330 $O/bench_w5b.json; echo
echo "total elapsed $(( $(date +%s) - $(cat $O/t0) )) s"
