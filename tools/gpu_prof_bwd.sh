#!/bin/bash
mkdir -p gpurun_out
CMD="python tools/profile_run.py --fwd 2 --bwd 2,5,1 --iters 2"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none -k regex:msda_bwd -s 3 -c 3 -f -o gpurun_out/prof_bwd $CMD > gpurun_out/ncu_bwd.log 2>&1
tail -2 gpurun_out/ncu_bwd.log
