#!/bin/bash
# Final artefacts of the round: bench lines of the default configuration and the widened rows, reference arm,
# ncu launch list of the same command.
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out; O=gpurun_out
date +%s > $O/t0
python bench.py > $O/bench_final.json 2> $O/bench_final.err
python bench.py --config 6 > $O/bench_fd_final.json 2> $O/bench_fd_final.err
python bench.py --config 2 --no-cpu-baseline > $O/bench_robin_final.json 2> $O/bench_robin_final.err
python bench.py --impl reference --steps 3 --warmup 3 > $O/bench_reference_arm.json 2> $O/bench_reference_arm.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 400 --csv --log-file $O/launches_final3.csv \
    python bench.py --steps 3 --warmup 4 --no-cpu-baseline > $O/ncu_bench3.log 2>&1
cut -c1-300 $O/bench_final.json; echo; cut -c1-300 $O/bench_fd_final.json; echo; cut -c1-300 $O/bench_robin_final.json; echo; cut -c1-200 $O/bench_reference_arm.json; echo
echo "total elapsed $(( $(date +%s) - $(cat $O/t0) )) s"
