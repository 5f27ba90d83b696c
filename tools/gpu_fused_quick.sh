#!/bin/bash
# fused-producer parity tests + op-level fused vs unfused timing (one short GPU call)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_fused_gpu.py -x -q -m gpu 2>&1 | tail -5
timeout 300 python tools/bench_configs.py --configs f --out gpurun_out/configs_fused_quick.jsonl 2>&1 | grep '^{' | cut -c1-400
