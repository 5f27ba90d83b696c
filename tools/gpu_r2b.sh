#!/bin/bash
# usage: tools/gpu_r2b.sh "<variants>" [workloads] : variants parity + sweep
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "variants or golden" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
rm -f gpurun_out/sweep.jsonl
timeout 600 python tools/sweep.py --out gpurun_out/sweep.jsonl --variants "$1" --caps 0 --orders 0 --workloads "${2:-cityscapes_512x1024_b8}" > gpurun_out/sweep.log 2>&1; echo "sweep exit $?"
python - <<'PY'
import json
for l in open('gpurun_out/sweep.jsonl'):
    r=json.loads(l); print(r['workload'], r['mode'], r['variant'], 'fwd %.3f bwd %.3f ms'%(r['fwd_ms'], r['bwd_ms']))
PY
