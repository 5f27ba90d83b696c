#!/bin/bash
# Profiles of the bench command (run only after the plain command exits 0):
#   1. every launch with its device time  -> gpurun_out/launches.csv
#   2. ncu --set full of the library's kernels in the timed region -> gpurun_out/prof_bench.ncu-rep
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:msda_ -s 6 -c 4 -f -o gpurun_out/prof_bench $CMD > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_list.log; tail -2 gpurun_out/ncu_full.log
