#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests -m gpu -x -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu.log
for wl in target c3; do
timeout 600 python bench.py --workload $wl --steps 30 --warmup 3 --no-cpu > gpurun_out/bench_${wl}_n1.json 2> gpurun_out/bench_${wl}_n1.err; echo "$wl n1 rc=$?"; tail -3 gpurun_out/bench_${wl}_n1.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 2 --workload $wl --steps 30 --warmup 3 > gpurun_out/bench_${wl}_n2.json 2> gpurun_out/bench_${wl}_n2.err; echo "$wl n2 rc=$?"; tail -5 gpurun_out/bench_${wl}_n2.err
for n in 1 2; do python -c "
import json,sys
d=json.loads(open('gpurun_out/bench_${wl}_n$n.json').read().strip().splitlines()[-1])
print('$wl n=$n value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'scan_ms',d['roofline']['scan_ms'],'GB/s',round(d['roofline']['achieved']),'frac',round(d['roofline']['frac'],3), d['clocks'])
"; done
done
