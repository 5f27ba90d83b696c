#!/bin/bash
mkdir -p gpurun_out
timeout 300 ./tools/microbench > gpurun_out/microbench2.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -k "variants or golden" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
rm -f gpurun_out/sweep.jsonl
timeout 600 python tools/sweep.py --out gpurun_out/sweep.jsonl --variants 2,9,10,11,12 --caps 0 --orders 0 > gpurun_out/sweep.log 2>&1; echo "sweep exit $?" >> gpurun_out/sweep.log
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/microbench2.txt
