#!/bin/bash
# ncu --set full of one launch of the weight-gradient kernel (172 032 x 256 x 256), after the plain run
mkdir -p gpurun_out
CMD="python tools/bench_weight_grad.py"
timeout 300 $CMD > gpurun_out/wgrad_plain.log 2>&1 && \
timeout 300 ncu --set full --clock-control none --import-source on -k regex:linear_wgrad -s 12 -c 1 -f -o gpurun_out/prof_wgrad $CMD > gpurun_out/ncu_wgrad.log 2>&1
tail -2 gpurun_out/ncu_wgrad.log
