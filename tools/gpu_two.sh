#!/bin/bash
# 2-GPU box: the multi-device tests, bench.py at N=2, and the multi-rank configs with the inference / training kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "multi_device or linear or sharding" 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29602 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/two_n2.json 2> gpurun_out/two_n2.err; echo "N=2 exit $?"; head -c 300 gpurun_out/two_n2.json; echo
rm -f gpurun_out/configs_two.jsonl
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29702 tools/bench_configs.py --configs 3,5,q --linear tf32x3 --fused --out gpurun_out/configs_two.jsonl > gpurun_out/configs_two.log 2>&1; echo "configs exit $?"
cut -c1-260 gpurun_out/configs_two.jsonl
