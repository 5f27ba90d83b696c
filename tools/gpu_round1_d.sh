#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
rm -f gpurun_out/configs.jsonl
timeout 900 python tools/bench_configs.py --configs 1,2,3,4,5,q > gpurun_out/configs.log 2>&1; echo "configs exit $?" >> gpurun_out/configs.log
tail -4 gpurun_out/pytest_gpu.log; tail -12 gpurun_out/configs.log
