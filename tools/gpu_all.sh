#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
rm -f gpurun_out/vs_reference_cuda.jsonl
timeout 900 python tools/bench_vs_reference_cuda.py > gpurun_out/vs_reference_cuda.log 2>&1
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/bench.err
tail -1 gpurun_out/smoke.log; tail -4 gpurun_out/pytest_gpu.log; cut -c1-330 gpurun_out/vs_reference_cuda.log | tail -6; tail -1 gpurun_out/bench.err; head -c 300 gpurun_out/bench.json
