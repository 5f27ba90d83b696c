#!/bin/bash
# usage: tools/gpu_r2_multi.sh N   (run with gpurun --gpus N): N-rank NCCL test + bench at 1..N ranks
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -q -x > gpurun_out/pytest_multi.log 2>&1; echo "pytest multi exit $?"; tail -3 gpurun_out/pytest_multi.log
for n in 1 2 4 8; do
  if [ $n -le $N ]; then
    if [ $n -eq 1 ]; then
      timeout 900 python bench.py --gpus 1 > gpurun_out/scale_n1.json 2> gpurun_out/scale_n1.err
    else
      NCCL_DEBUG=INFO timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n)) bench.py --gpus $n > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
    fi
    echo "n=$n exit $?"; python -c "
import json; d=json.loads(open('gpurun_out/scale_n$n.json').read().strip().splitlines()[-1]); print($n, d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'])"
  fi
done
grep -c "NCCL INFO.*nranks" gpurun_out/scale_n$N.err
