#!/bin/bash
# round 2: the decoder glue kernels (row-bias GEMM, packed fused forward, GroupNorm with the conv bias,
# channel bias, transposes): their tests, then the decoder forward at configs[2] / [3] and its launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_linear_gpu.py tests/test_fused_gpu.py tests/test_pixel_decoder.py tests/test_modules.py tests/test_reference_stack_gpu.py -m gpu -q -x > gpurun_out/pytest_d.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_d.log
tail -15 gpurun_out/pytest_d.log
rm -f gpurun_out/configs_r2.jsonl
timeout 900 python tools/bench_configs.py --configs 3,4 --linear tf32x3 --fused --out gpurun_out/configs_r2.jsonl > gpurun_out/configs_r2.log 2>&1; echo "configs exit $?"
cut -c1-420 gpurun_out/configs_r2.jsonl
LINEAR=tf32x3 FUSED=1 bash tools/gpu_prof_decoder.sh
