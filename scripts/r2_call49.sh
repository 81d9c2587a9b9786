#!/bin/bash
mkdir -p gpurun_out/c49
cd /root/repo
timeout 400 python -m pytest tests/test_host_class.py -x -q -k "row_slabs or matches_oracle" 2>&1 | tail -15 | tee gpurun_out/c49/pytest_host_slab.log
