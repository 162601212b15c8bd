#!/bin/bash
# A/B of the K1 front-ends on one GPU (scan_mode 1 = autonomous warps, 2 = planner + workers)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_planner.py -x -q 2>&1 | tail -5
for w in target c3; do for m in 2; do
OTTERS_SCAN_MODE=$m timeout 600 python bench.py --workload $w --steps 50 --warmup 5 --no-cpu > gpurun_out/ab_${w}_$m.json 2> gpurun_out/ab_${w}_$m.err; python -c "
import json; d=json.load(open('gpurun_out/ab_${w}_$m.json')); print('$w mode=$m', 'step_ms', round(d['ms_per_step'],4), 'e2e_ms', round(d['e2e']['ms_per_step'],4), 'scan_ms', round(d['phases_ms']['scan'],4), 'frac', round(d['roofline']['frac'],3))"
done; done
for m in 2; do
OTTERS_SCAN_MODE=$m timeout 600 python bench.py --workload target --rows 1250000 --steps 100 --warmup 10 --no-cpu > gpurun_out/ab_shard_$m.json 2> gpurun_out/ab_shard_$m.err; python -c "
import json; d=json.load(open('gpurun_out/ab_shard_$m.json')); print('target 1.25M rows mode=$m', 'step_ms', round(d['ms_per_step'],4), 'e2e_ms', round(d['e2e']['ms_per_step'],4), 'scan_ms', round(d['phases_ms']['scan'],4), 'frac', round(d['roofline']['frac'],3))"
done
