#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out; O=gpurun_out
export EQGPU_LS_DEBUG=1
EQGPU_LS_FORM=1 python scripts/ls_debug.py 2048 45 2> $O/dbg_2048_form1.log
EQGPU_LS_FORM=0 python scripts/ls_debug.py 2048 30 2> $O/dbg_2048_form0.log
EQGPU_WARM=3 python scripts/ls_debug.py 2048 45 2> $O/dbg_2048_w3.log
EQGPU_LS_FORM=1 python scripts/ls_debug.py 257 20 2> $O/dbg_257_form1.log
tail -3 $O/dbg_2048_form1.log $O/dbg_2048_form0.log $O/dbg_2048_w3.log $O/dbg_257_form1.log
