#!/bin/bash
# launch list (device time per kernel) of the pixel-decoder forward at BASELINE configs[2] (1024x2048, batch 8)
mkdir -p gpurun_out
CMD="python tools/bench_configs.py --configs 3 --steps 1 ${LINEAR:+--linear $LINEAR} ${FUSED:+--fused} --out gpurun_out/prof_decoder.jsonl"
timeout 600 $CMD > gpurun_out/prof_decoder_plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_decoder.csv $CMD > gpurun_out/prof_decoder_ncu.log 2>&1
tail -3 gpurun_out/prof_decoder_ncu.log; wc -l gpurun_out/launches_decoder.csv
