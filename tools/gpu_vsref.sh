#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/vs_reference_cuda.jsonl
timeout 600 python -m pytest tests/test_parity_gpu.py -q -k reference_cuda > gpurun_out/pytest_refcuda.log 2>&1; tail -3 gpurun_out/pytest_refcuda.log
timeout 900 python tools/bench_vs_reference_cuda.py > gpurun_out/vs_reference_cuda.log 2>&1; cut -c1-700 gpurun_out/vs_reference_cuda.log | tail -8
