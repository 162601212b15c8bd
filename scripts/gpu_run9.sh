#!/bin/bash
# c2: batched tcgen05 path — bench line, ncu full capture, launch list
mkdir -p gpurun_out
timeout 600 python bench.py --workload c2 --steps 10 --warmup 3 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "c2 rc=$?"; cat gpurun_out/bench_c2.json; tail -3 gpurun_out/bench_c2.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:batch_kernel -s 2 -c 1 -o gpurun_out/batch_c2 python bench.py --workload c2 --steps 2 --warmup 2 --no-cpu > gpurun_out/ncu_c2.log 2>&1; echo "ncu rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_c2.csv python bench.py --workload c2 --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_launch_c2.log 2>&1; echo "launch list rc=$?"
