#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/sweep_r2.jsonl
timeout 1200 python tools/sweep.py --out gpurun_out/sweep_r2.jsonl --variants "0,2,20,21,24,25,26,27" --caps 0 --orders 0 --workloads "cityscapes_512x1024_b8,cityscapes_1024x2048_b1,cityscapes_1024x2048_b8,kitti_384x1248_b16" > gpurun_out/sweep_r2.log 2>&1; echo "sweep exit $?"
python - <<'PY'
import json
for l in open('gpurun_out/sweep_r2.jsonl'):
    r=json.loads(l); print(r['workload'], r['mode'], r['variant'], 'fwd %.3f bwd %.3f ms'%(r['fwd_ms'], r['bwd_ms']))
PY
