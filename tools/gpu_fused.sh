#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/configs_fused.jsonl
timeout 900 python tools/bench_configs.py --configs f,3,5 --out gpurun_out/configs_fused.jsonl > gpurun_out/configs_fused.log 2>&1
timeout 900 python tools/bench_configs.py --configs 3,5 --fused --out gpurun_out/configs_fused.jsonl >> gpurun_out/configs_fused.log 2>&1
grep '^{' gpurun_out/configs_fused.log | cut -c1-420
