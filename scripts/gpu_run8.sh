#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python bench.py --workload c2 --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "c2 rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_c2.json')); print(d['value'], d['ms_per_step'], d['e2e']['value']); print(d['roofline'])"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_target.csv python bench.py --steps 3 --warmup 2 --no-cpu > gpurun_out/ncu_launch.log 2>&1; echo "launch list rc=$?"
