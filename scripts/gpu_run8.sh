#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_batched.py -q --tb=line 2>&1 | tail -30 > gpurun_out/pytest_batched.log; echo "pytest rc=${PIPESTATUS[0]}"; cat gpurun_out/pytest_batched.log
timeout 300 python scripts/dbg_batched.py > gpurun_out/dbg_batched.log 2>&1; echo "dbg rc=$?"; cat gpurun_out/dbg_batched.log | tail -20
