#!/bin/bash
mkdir -p gpurun_out
timeout 900 python scripts/sweep_tuning.py c4 > gpurun_out/sweep_c4.log 2>&1; echo rc=$?; sort -t= -k5 -n -r gpurun_out/sweep_c4.log | head -3
timeout 900 python scripts/sweep_tuning.py target > gpurun_out/sweep_target.log 2>&1; echo rc=$?
timeout 900 python scripts/sweep_tuning.py c3 > gpurun_out/sweep_c3.log 2>&1; echo rc=$?
timeout 900 python scripts/sweep_tuning.py c5 > gpurun_out/sweep_c5.log 2>&1; echo rc=$?
tail -3 gpurun_out/sweep_*.log
