#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
for wl in target c4 c3 c5 c1; do
  timeout 600 python bench.py --workload $wl --steps 30 --warmup 3 --no-cpu > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err; echo "$wl rc=$?"; tail -3 gpurun_out/bench_$wl.err
  python -c "
import json,sys
d=json.loads(open('gpurun_out/bench_$wl.json').read().strip().splitlines()[-1])
print('$wl value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'scan_ms',d['roofline']['scan_ms'],'GB/s',round(d['roofline']['achieved']),'frac',round(d['roofline']['frac'],3), d['clocks'])
"
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_target.csv python bench.py --steps 3 --warmup 2 --no-cpu > gpurun_out/ncu_launch.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan_kernel -s 2 -c 1 -o gpurun_out/scan_c3 python bench.py --workload c3 --steps 2 --warmup 2 --no-cpu > gpurun_out/ncu_c3.log 2>&1; echo "c3 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan_kernel -s 2 -c 1 -o gpurun_out/scan_c5 python bench.py --workload c5 --steps 2 --warmup 2 --no-cpu > gpurun_out/ncu_c5.log 2>&1; echo "c5 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan_kernel -s 2 -c 1 -o gpurun_out/scan_target python bench.py --steps 2 --warmup 2 --no-cpu > gpurun_out/ncu_target.log 2>&1; echo "target rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rowmask_kernel -s 2 -c 1 -o gpurun_out/rowmask_target python bench.py --steps 2 --warmup 2 --no-cpu > gpurun_out/ncu_rm.log 2>&1; echo "rowmask rc=$?"
