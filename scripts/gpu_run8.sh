#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for rows in 0 1250000; do for w in target c4; do
timeout 600 python bench.py --workload $w --rows $rows --steps 60 --warmup 5 --no-cpu > gpurun_out/pl.json 2> gpurun_out/pl.err; python -c "
import json; d=json.load(open('gpurun_out/pl.json')); print('$w rows=$rows', 'step_ms', round(d['ms_per_step'],4), 'e2e_ms', round(d['e2e']['ms_per_step'],4), d['phases_ms'], 'frac', round(d['roofline']['frac'],3))"
done; done
for w in c3 c1; do
timeout 600 python bench.py --workload $w --steps 60 --warmup 5 --no-cpu > gpurun_out/pl.json 2> gpurun_out/pl.err; python -c "
import json; d=json.load(open('gpurun_out/pl.json')); print('$w', 'step_ms', round(d['ms_per_step'],4), 'e2e_ms', round(d['e2e']['ms_per_step'],4), d['phases_ms'], 'frac', round(d['roofline']['frac'],3))"
done
