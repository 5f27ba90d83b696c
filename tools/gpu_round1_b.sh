#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
rm -f gpurun_out/sweep.jsonl
timeout 600 python tools/sweep.py --out gpurun_out/sweep.jsonl --variants 2,5,6,7,8 --caps 0 > gpurun_out/sweep.log 2>&1; echo "sweep exit $?" >> gpurun_out/sweep.log
CMD="python tools/profile_run.py --fwd 2,6 --bwd 2 --iters 2"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:msda_ -s 3 -c 3 -f -o gpurun_out/prof_r1b $CMD > gpurun_out/ncu_full.log 2>&1
tail -5 gpurun_out/pytest_gpu.log; tail -3 gpurun_out/ncu_full.log
