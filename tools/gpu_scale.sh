#!/bin/bash
# Scaling check as the driver runs it: bench.py at N = 1, 2, 4, 8 on one box (torchrun for N > 1),
# reference arm once, plus the multi-rank configs at the largest N.
mkdir -p gpurun_out
NG=${1:-8}
for n in 1 2 4 8; do
  [ $n -gt $NG ] && break
  if [ $n -eq 1 ]; then
    timeout 600 python bench.py --gpus 1 --steps 30 --warmup 5 > gpurun_out/scale_n1.json 2> gpurun_out/scale_n1.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) bench.py --gpus $n --steps 30 --warmup 5 > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
  fi
  echo "N=$n exit $?"; head -c 400 gpurun_out/scale_n$n.json; echo
done
rm -f gpurun_out/configs_n$NG.jsonl
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29700 tools/bench_configs.py --configs 2,3,5,q --out gpurun_out/configs_n$NG.jsonl > gpurun_out/configs_n$NG.log 2>&1; echo "configs exit $?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29701 tools/bench_configs.py --configs 3,5,q --linear tf32x3 --fused --out gpurun_out/configs_n$NG.jsonl >> gpurun_out/configs_n$NG.log 2>&1; echo "configs (tf32x3, fused) exit $?"
cut -c1-300 gpurun_out/configs_n$NG.jsonl
