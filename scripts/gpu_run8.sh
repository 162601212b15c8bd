#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for w in target c3 c5; do
timeout 600 python bench.py --workload $w --steps 50 --warmup 5 --no-cpu > gpurun_out/bench_${w}_n1.json 2> gpurun_out/bench_${w}_n1.err; python -c "
import json; d=json.load(open('gpurun_out/bench_${w}_n1.json')); print('$w n1', round(d['value'],1), round(d['ms_per_step'],4), round(d['e2e']['value'],1), d['phases_ms'], round(d['roofline']['frac'],3))"
done
timeout 600 python bench.py --workload target --rows 1250000 --steps 100 --warmup 10 --no-cpu > gpurun_out/b_shard.json 2> gpurun_out/b_shard.err; python -c "
import json; d=json.load(open('gpurun_out/b_shard.json')); print('shard', 'step_ms', round(d['ms_per_step'],4), 'e2e_ms', round(d['e2e']['ms_per_step'],4), d['phases_ms'], round(d['roofline']['frac'],3))"
