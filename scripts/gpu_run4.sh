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
