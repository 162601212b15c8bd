#!/bin/bash
# round 2, second session, run 6: role-by-role timing of K2's bf16 rung with the -DOTTERS_K2_EXPERIMENTS build of the library
# (otters_b200/libotters_b200_dbg.so, swapped in on the box only), kernel times from an ncu launch list.
mkdir -p gpurun_out/r2b6
O=gpurun_out/r2b6
cp otters_b200/libotters_b200_dbg.so otters_b200/libotters_b200.so
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:batch_kernel -c 12 --csv --log-file $O/roles_launches.csv python scripts/dbg_roles_bf16.py > $O/roles.log 2>&1; echo "roles rc=$?"
tail -3 $O/roles.log
grep batch_kernel $O/roles_launches.csv | awk -F'","' '{print $NF}'
OTTERS_K2_EPI_WARPS=4 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:batch_kernel -c 12 --csv --log-file $O/roles_launches_epi4.csv python scripts/dbg_roles_bf16.py > $O/roles_epi4.log 2>&1; echo "roles epi4 rc=$?"
grep batch_kernel $O/roles_launches_epi4.csv | awk -F'","' '{print $NF}'
