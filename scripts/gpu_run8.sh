#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/dbg_flake.py > gpurun_out/dbg_flake.log 2>&1; echo rc=$?
grep -B1 "verified=0" gpurun_out/dbg_flake.log | head -10; grep -c "verified=1" gpurun_out/dbg_flake.log
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4
