#!/bin/bash
# One `ncu --set full` capture of a named kernel of the hot path (tools/ncu_pass.py: 400 chunks, third pass).
#   tools/ncu_kernel.sh <kernel regex> <output stem under gpurun_out/>
set -e
K=$(python tools/ncu_pass.py --count)
ncu --set full --clock-control none --import-source on -k regex:$1 --launch-skip 2 --launch-count 1 -f -o gpurun_out/$2 python tools/ncu_pass.py > gpurun_out/$2.log 2>&1
