#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/exp_e2e.py > gpurun_out/exp_e2e.log 2>&1; cat gpurun_out/exp_e2e.log
# launch list of the default bench (shares, cold-cache serialised)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_target.csv \
   python bench.py --steps 3 --warmup 2 --no-cpu > gpurun_out/ncu_launch.log 2>&1; echo "launch list rc=$?"
# full capture of the scan kernel (target + c4)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan_kernel -s 2 -c 2 -o gpurun_out/scan_target \
   python bench.py --steps 2 --warmup 2 --no-cpu > gpurun_out/ncu_full_target.log 2>&1; echo "full target rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan_kernel -s 2 -c 2 -o gpurun_out/scan_c4 \
   python bench.py --workload c4 --steps 2 --warmup 2 --no-cpu > gpurun_out/ncu_full_c4.log 2>&1; echo "full c4 rc=$?"
ls -la gpurun_out
