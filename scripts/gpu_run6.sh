#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 2 --steps 30 --warmup 3 --phase-timing > gpurun_out/bench_target_n2.json 2> gpurun_out/bench_target_n2.err; echo "n2 rc=$?"; grep phases gpurun_out/bench_target_n2.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:select_kernel -s 2 -c 1 -o gpurun_out/select_target python bench.py --steps 2 --warmup 2 --no-cpu > gpurun_out/ncu_sel.log 2>&1; echo "select rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:prune_kernel -s 2 -c 1 -o gpurun_out/prune_target python bench.py --steps 2 --warmup 2 --no-cpu > gpurun_out/ncu_prune.log 2>&1; echo "prune rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_kernel -s 2 -c 1 -o gpurun_out/scan_target python bench.py --steps 2 --warmup 2 --no-cpu > gpurun_out/ncu_target.log 2>&1; echo "target rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_kernel -s 2 -c 1 -o gpurun_out/scan_c4 python bench.py --workload c4 --steps 2 --warmup 2 --no-cpu > gpurun_out/ncu_c4.log 2>&1; echo "c4 rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_target.csv python bench.py --steps 3 --warmup 2 --no-cpu > gpurun_out/ncu_launch.log 2>&1; echo "launch list rc=$?"
