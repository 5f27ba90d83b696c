#!/bin/bash
# round 2: full GPU suite + smoke + sweep of the backward dispatch + bench line
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -2 gpurun_out/smoke.log; tail -6 gpurun_out/pytest_gpu.log
rm -f gpurun_out/sweep.jsonl
timeout 600 python tools/sweep.py --out gpurun_out/sweep.jsonl --variants "${1:-0,2,20}" --caps 0 --orders 0 > gpurun_out/sweep.log 2>&1; echo "sweep exit $?"
python - <<'PY'
import json
for l in open('gpurun_out/sweep.jsonl'):
    r=json.loads(l); print(r['workload'], r['mode'], r['variant'], 'fwd %.3f bwd %.3f ms'%(r['fwd_ms'], r['bwd_ms']))
PY
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; head -c 400 gpurun_out/bench.json
