#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/sweep.jsonl
timeout 600 python tools/sweep.py --out gpurun_out/sweep.jsonl --variants 2 --caps 0 --orders 0 > gpurun_out/sweep.log 2>&1; echo "sweep exit $?" >> gpurun_out/sweep.log
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_parity_gpu.py -q -x -k "golden or variants or lattice" > gpurun_out/sanitizer_memcheck.log 2>&1; echo "memcheck exit $?" >> gpurun_out/sanitizer_memcheck.log
tail -2 gpurun_out/sweep.log; tail -12 gpurun_out/sanitizer_memcheck.log
