#!/bin/bash
# refresh the ncu captures of the small per-query kernels and of K1 on the target workload
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:select_kernel -s 3 -c 1 -o gpurun_out/select_target -f python bench.py --steps 2 --warmup 2 --no-cpu > gpurun_out/ncu_sel.log 2>&1; echo "select rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:prune_leafpar_kernel -s 3 -c 1 -o gpurun_out/prune_target -f python bench.py --steps 2 --warmup 2 --no-cpu > gpurun_out/ncu_prune.log 2>&1; echo "prune rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_kernel -s 3 -c 1 -o gpurun_out/scan_target -f python bench.py --steps 2 --warmup 2 --no-cpu > gpurun_out/ncu_scan.log 2>&1; echo "scan rc=$?"
