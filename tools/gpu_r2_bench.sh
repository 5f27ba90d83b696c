#!/bin/bash
# round 2: both bench arms, then (only after the plain command exited 0) the ncu launch list and the
# ncu --set full capture of the same command
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
timeout 900 python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench ref exit $?"
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:msda_ -s 9 -c 6 -f -o gpurun_out/prof_bench $CMD > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_list.log; tail -2 gpurun_out/ncu_full.log
python - <<'PY'
import json
for f in ('gpurun_out/bench.json','gpurun_out/bench_ref.json'):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d.get('value'), d.get('ms_per_step'), d.get('roofline',{}).get('frac'), d.get('e2e',{}).get('value'), d.get('cpu_baseline',{}).get('kind'), d.get('cpu_baseline',{}).get('cores'))
        print({k:(v.get('ms') if isinstance(v,dict) else v) for k,v in d.get('kernels',{}).items()})
    except Exception as e: print(f, 'ERR', e)
PY
