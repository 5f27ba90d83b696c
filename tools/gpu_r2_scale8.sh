#!/bin/bash
# usage: tools/gpu_r2_scale8.sh N  (gpurun --gpus N): N-rank bench line, N-rank NCCL tests, query-sharded single-image latency
N=${1:-8}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_multi_gpu.py -m gpu -q -x > gpurun_out/pytest_multi.log 2>&1; echo "pytest multi exit $?"; tail -2 gpurun_out/pytest_multi.log
NCCL_DEBUG=INFO timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29508 bench.py --gpus $N --steps 30 --warmup 5 > gpurun_out/scale_n$N.json 2> gpurun_out/scale_n$N.err
echo "bench N=$N exit $?"; head -c 300 gpurun_out/scale_n$N.json; echo; grep -c "NCCL INFO.*nranks" gpurun_out/scale_n$N.err
timeout 300 python bench.py --gpus 1 --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/scale_n1.json 2> gpurun_out/scale_n1.err; echo "bench N=1 exit $?"; head -c 200 gpurun_out/scale_n1.json; echo
rm -f gpurun_out/configs_n$N.jsonl
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29701 tools/bench_configs.py --configs q,5 --linear tf32x3 --fused --out gpurun_out/configs_n$N.jsonl > gpurun_out/configs_n$N.log 2>&1; echo "configs exit $?"
cut -c1-420 gpurun_out/configs_n$N.jsonl
