#!/bin/bash
# tensor-core fp32 Linear: tests, then the pixel decoder / encoder with and without it
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_linear_gpu.py tests/test_pixel_decoder.py -x -q -m gpu 2>&1 | tail -6
rm -f gpurun_out/configs_linear.jsonl
timeout 600 python tools/bench_configs.py --configs 3,4 --out gpurun_out/configs_linear.jsonl 2>&1 | grep '^{' | cut -c1-330
timeout 600 python tools/bench_configs.py --configs 3,4 --linear tf32x3 --out gpurun_out/configs_linear.jsonl 2>&1 | grep '^{' | cut -c1-330
timeout 600 python tools/bench_configs.py --configs 3,4 --linear tf32x3 --fused --out gpurun_out/configs_linear.jsonl 2>&1 | grep '^{' | cut -c1-330
