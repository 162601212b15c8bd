#!/bin/bash
# one GPU call: single-pass K2 numbers, c2 bench line, then the whole GPU suite
mkdir -p gpurun_out
OTTERS_BATCH_TRACE=1 timeout 120 python scripts/dbg_passes.py > gpurun_out/passes.log 2>&1
echo "dbg_passes rc=$?" >> gpurun_out/passes.log
timeout 120 python bench.py --workload c2 --steps 20 --warmup 3 > gpurun_out/bench_c2_auto.json 2> gpurun_out/bench_c2_auto.err
echo "bench rc=$?"
timeout 400 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"
tail -5 gpurun_out/pytest_gpu.log
grep -v "^\[otters batch\]" gpurun_out/passes.log | tail -25
