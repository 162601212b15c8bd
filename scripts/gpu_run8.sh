#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for w in target c3; do
timeout 600 python bench.py --workload $w --steps 50 --warmup 5 --no-cpu > gpurun_out/bench_${w}_n1.json 2> gpurun_out/bench_${w}_n1.err; python -c "
import json; d=json.load(open('gpurun_out/bench_${w}_n1.json')); print('$w n1', round(d['value'],1), round(d['ms_per_step'],4), round(d['e2e']['value'],1), d['phases_ms'], round(d['roofline']['frac'],3))"
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_target.csv python bench.py --steps 3 --warmup 2 --no-cpu > gpurun_out/ncu_launch.log 2>&1; echo "launch list rc=$?"
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/launches_target.csv')))
hdr=None; out=[]
for r in rows:
    if 'Kernel Name' in r: hdr=r; continue
    if hdr and len(r)==len(hdr):
        d=dict(zip(hdr,r)); out.append((d['Kernel Name'][:40], d['Metric Value']))
for o in out[-6:]: print(o)
PY
