#!/bin/bash
# round 2, call A: parity of the new backward kernel, variant sweep, bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_checked_gpu.py -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
rm -f gpurun_out/sweep.jsonl
timeout 600 python tools/sweep.py --out gpurun_out/sweep.jsonl --variants "2,20" --caps 0 --orders 0 > gpurun_out/sweep.log 2>&1; echo "sweep exit $?"
python - <<'PY'
import json
for l in open('gpurun_out/sweep.jsonl'):
    r=json.loads(l); print(r['workload'], r['mode'], r['variant'], 'fwd %.3f bwd %.3f ms'%(r['fwd_ms'], r['bwd_ms']))
PY
timeout 600 python bench.py --no-e2e --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; head -c 600 gpurun_out/bench.json
