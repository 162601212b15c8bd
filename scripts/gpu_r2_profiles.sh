#!/bin/bash
# round 2 profiles: launch lists (ncu gpu__time_duration) and full captures of the dominant kernels at HEAD.
# The .ncu-rep files stay on the box (/tmp): only their text summaries (scripts/ncu_summary.py, ncu_hot.py) and two reports
# travel back (gpurun_out/ is limited to 64 MiB).
mkdir -p gpurun_out /tmp/rep
B="python bench.py --no-cpu --no-parity --blocking --steps 3 --warmup 2"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r2_launches_target.csv $B > gpurun_out/ncu_l1.log 2>&1; echo "launch list target rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r2_launches_c3.csv $B --workload c3 > gpurun_out/ncu_l2.log 2>&1; echo "launch list c3 rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r2_launches_c2.csv $B --workload c2 > gpurun_out/ncu_l3.log 2>&1; echo "launch list c2 rc=$?"
cap() { name=$1; kern=$2; shift 2
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$kern -s 3 -c 1 -o /tmp/rep/r2_$name -f $B "$@" > gpurun_out/ncu_$name.log 2>&1; echo "ncu $name rc=$?"
  python scripts/ncu_summary.py /tmp/rep/r2_$name.ncu-rep > gpurun_out/r2_${name}_summary.txt 2>&1
  python scripts/ncu_hot.py /tmp/rep/r2_$name.ncu-rep 40 > gpurun_out/r2_${name}_hot_sass.txt 2>&1
}
cap scan_target scan_planner_kernel
cap scan_c3 "scan_kernel" --workload c3
cap rowmask_target rowmask_kernel
cap prune_target prune_leafpar_kernel
cap scan_c5 scan_planner_kernel --workload c5
cap batch_c2 batch_kernel --workload c2
cap scan_c1 "scan_kernel" --workload c1
cp /tmp/rep/r2_scan_target.ncu-rep /tmp/rep/r2_scan_c3.ncu-rep gpurun_out/ 2>/dev/null
du -sh gpurun_out
