#!/bin/bash
# usage: tools/gpu_r2_prof.sh <bwd variants> [mode]: ncu --set full of the backward kernels (one launch each, second iteration)
mkdir -p gpurun_out
MODE=${2:-model}
CMD="python tools/profile_run.py --fwd 0 --bwd $1 --mode $MODE --iters 2"
$CMD > gpurun_out/prof_plain.log 2>&1 || { tail -5 gpurun_out/prof_plain.log; exit 1; }
NV=$(echo "$1" | tr ',' '\n' | wc -l)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:msda_bwd -s $NV -c $NV -f -o gpurun_out/prof_bwd_r2 $CMD > gpurun_out/ncu_bwd_r2.log 2>&1
tail -3 gpurun_out/ncu_bwd_r2.log
