#!/bin/bash
# usage: tools/gpu_sweep.sh "<variants>" ["<pytest -k expr>"]
mkdir -p gpurun_out
if [ -n "$2" ]; then
  timeout 1500 python -m pytest tests -m gpu -q -x -k "$2" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
  tail -3 gpurun_out/pytest_gpu.log
fi
rm -f gpurun_out/sweep.jsonl
timeout 900 python tools/sweep.py --out gpurun_out/sweep.jsonl --variants "$1" --caps 0 --orders 0 > gpurun_out/sweep.log 2>&1; echo "sweep exit $?"
