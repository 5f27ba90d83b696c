#!/bin/bash
# 2-GPU pass: GPU tests, scaling bench (1 and 2 ranks), multi-rank configs
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --gpus 1 --steps 30 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_ref_n2.json 2> gpurun_out/bench_ref_n2.err
rm -f gpurun_out/configs_n2.jsonl
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/bench_configs.py --configs 2,3,5,q --out gpurun_out/configs_n2.jsonl > gpurun_out/configs_n2.log 2>&1; echo "configs exit $?" >> gpurun_out/configs_n2.log
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/bench_n2.json | head -c 1500; echo; tail -3 gpurun_out/bench_n2.err; cat gpurun_out/bench_ref_n2.json | head -c 600; echo; tail -6 gpurun_out/configs_n2.log
