#!/bin/bash
# round 2: full GPU suite, then decoder forward (configs[2], [3]) and the training step (configs[4]) with the
# inference / training kernels, then the decoder launch list
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
rm -f gpurun_out/configs_r2.jsonl
timeout 900 python tools/bench_configs.py --configs 3,4,5 --linear tf32x3 --fused --out gpurun_out/configs_r2.jsonl > gpurun_out/configs_r2.log 2>&1; echo "configs exit $?"
cut -c1-360 gpurun_out/configs_r2.jsonl
LINEAR=tf32x3 FUSED=1 bash tools/gpu_prof_decoder.sh
