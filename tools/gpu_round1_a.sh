#!/bin/bash
# First GPU pass: smoke, parity tests, bench, variant sweep, then (only if parity is green) ncu.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_all.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_all.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/bench.err
timeout 600 python tools/sweep.py --out gpurun_out/sweep.jsonl > gpurun_out/sweep.log 2>&1; echo "sweep exit $?" >> gpurun_out/sweep.log
if grep -q "pytest exit 0" gpurun_out/pytest_gpu.log; then
  CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline"
  $CMD > gpurun_out/plain.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
  $CMD > gpurun_out/plain2.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:msda_ -s 6 -c 4 -f -o gpurun_out/prof_r1a $CMD > gpurun_out/ncu_full.log 2>&1
fi
tail -5 gpurun_out/pytest_gpu.log; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json | head -c 3000
