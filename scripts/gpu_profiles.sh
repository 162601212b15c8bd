#!/bin/bash
# c2: batched tcgen05 path — ncu full capture of batch_kernel + launch list; target launch list with the new staging
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:batch_kernel -s 2 -c 1 -o gpurun_out/batch_c2 -f python bench.py --workload c2 --steps 2 --warmup 2 --no-cpu > gpurun_out/ncu_c2.log 2>&1; echo "ncu rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_c2.csv python bench.py --workload c2 --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_launch_c2.log 2>&1; echo "launch list c2 rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_target.csv python bench.py --steps 3 --warmup 2 --no-cpu > gpurun_out/ncu_launch.log 2>&1; echo "launch list target rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:select_kernel -s 3 -c 1 -o gpurun_out/select_target -f python bench.py --steps 2 --warmup 2 --no-cpu > gpurun_out/ncu_sel.log 2>&1; echo "select rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:prune_kernel -s 3 -c 1 -o gpurun_out/prune_target -f python bench.py --steps 2 --warmup 2 --no-cpu > gpurun_out/ncu_prune.log 2>&1; echo "prune rc=$?"
for w in target c4 c3 c5 c1 c2; do
  timeout 600 python bench.py --workload $w --steps 30 --warmup 5 > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; echo "$w rc=$?"
done
