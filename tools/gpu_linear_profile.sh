#!/bin/bash
# numbers for profiles/r1_linear.md: GEMM timings + what-if breakdown, decoder configs, one ncu capture of the GEMM
mkdir -p gpurun_out
timeout 200 python tools/bench_linear.py > gpurun_out/linear_try.jsonl 2>&1
timeout 200 python tools/whatif_linear.py 2>&1 | grep '^{' > gpurun_out/whatif_linear.jsonl
rm -f gpurun_out/configs_linear.jsonl
timeout 600 python tools/bench_configs.py --configs 3,4 --out gpurun_out/configs_linear.jsonl > /dev/null 2>&1
timeout 600 python tools/bench_configs.py --configs 3,4 --linear tf32x3 --out gpurun_out/configs_linear.jsonl > /dev/null 2>&1
timeout 600 python tools/bench_configs.py --configs 3,4 --linear tf32x3 --fused --out gpurun_out/configs_linear.jsonl > /dev/null 2>&1
rm -f gpurun_out/configs_train.jsonl
for a in "" "--linear tf32x3" "--linear tf32x3 --fused"; do timeout 400 python tools/bench_configs.py --configs 5 $a --out gpurun_out/configs_train.jsonl > /dev/null 2>&1; done
cut -c1-200 gpurun_out/configs_linear.jsonl gpurun_out/configs_train.jsonl
CMD="python tools/bench_linear.py"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:linear_tf32x3 -s 20 -c 1 -f -o gpurun_out/prof_linear $CMD > gpurun_out/ncu_linear.log 2>&1
tail -2 gpurun_out/ncu_linear.log
