#!/bin/bash
# launch list (device time per kernel) of the encoder training step (BASELINE configs[4]) with the round-1 kernels
mkdir -p gpurun_out
CMD="python tools/bench_configs.py --configs 5 --steps 1 --linear tf32x3 --fused --out gpurun_out/prof_train.jsonl"
timeout 600 $CMD > gpurun_out/prof_train_plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_train.csv $CMD > gpurun_out/prof_train_ncu.log 2>&1
tail -2 gpurun_out/prof_train_ncu.log | cut -c1-200; wc -l gpurun_out/launches_train.csv
