#!/bin/bash
# pixel-decoder forward (configs[2], [3]) and training step (configs[4]) with the round-2 kernels + launch list
mkdir -p gpurun_out
rm -f gpurun_out/configs_r2.jsonl
timeout 900 python tools/bench_configs.py --configs 3,4 --linear tf32x3 --fused --out gpurun_out/configs_r2.jsonl > gpurun_out/configs_r2.log 2>&1; echo "configs exit $?"
timeout 900 python tools/bench_configs.py --configs 3,4,5 --out gpurun_out/configs_r2.jsonl >> gpurun_out/configs_r2.log 2>&1; echo "configs (torch) exit $?"
timeout 900 python tools/bench_configs.py --configs 5 --linear tf32x3 --fused --out gpurun_out/configs_r2.jsonl >> gpurun_out/configs_r2.log 2>&1; echo "configs 5 exit $?"
cut -c1-420 gpurun_out/configs_r2.jsonl
LINEAR=tf32x3 FUSED=1 bash tools/gpu_prof_decoder.sh
