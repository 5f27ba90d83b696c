#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
rm -f gpurun_out/sweep.jsonl
timeout 900 python tools/sweep.py --out gpurun_out/sweep.jsonl --variants 1,2,3,4,5,6,7,8 --caps 0 --orders 0 --workloads cityscapes_512x1024_b8 > gpurun_out/sweep.log 2>&1; echo "sweep exit $?" >> gpurun_out/sweep.log
tail -3 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/sweep.log
