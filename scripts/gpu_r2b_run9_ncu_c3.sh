#!/bin/bash
# round 2, second session, run 9: ncu --set full of K1 on C3 with bf16 rows (256-byte rows: why 0.42 of the HBM peak).
mkdir -p gpurun_out/r2b9 /tmp/rep
O=gpurun_out/r2b9
B="python bench.py --no-cpu --no-parity --blocking --steps 3 --warmup 2"
timeout 150 ncu --set full --clock-control none --import-source on -k regex:scan_kernel -s 3 -c 1 -o /tmp/rep/r2b_scan_c3_bf16 -f $B --workload c3 --vector-format bf16 > $O/ncu.log 2>&1; echo "ncu rc=$?"
python scripts/ncu_summary.py /tmp/rep/r2b_scan_c3_bf16.ncu-rep > $O/r2b_scan_c3_bf16_summary.txt 2>&1
python scripts/ncu_hot.py /tmp/rep/r2b_scan_c3_bf16.ncu-rep 40 > $O/r2b_scan_c3_bf16_hot_sass.txt 2>&1
head -12 $O/r2b_scan_c3_bf16_summary.txt
