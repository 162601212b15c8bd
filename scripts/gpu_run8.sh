#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_batched.py -q --tb=line -x 2>&1 | tail -12 > gpurun_out/pytest_batched.log; echo "pytest rc=${PIPESTATUS[0]}"; cat gpurun_out/pytest_batched.log
timeout 600 python bench.py --workload c2 --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "c2 rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_c2.json')); print(d['value'], d['ms_per_step']); print(d['roofline']); print(d['e2e'])"
