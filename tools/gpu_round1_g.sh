#!/bin/bash
mkdir -p gpurun_out
CMD="python tools/profile_run.py --fwd 2 --bwd 4,2,3 --iters 2"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:msda_ -s 4 -c 4 -f -o gpurun_out/prof_r1c $CMD > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
