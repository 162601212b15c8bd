#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tests/dist_check.py > gpurun_out/dist_check_n2.log 2>&1; echo "dist_check rc=$?"; grep -E "DIST_CHECK|Error" gpurun_out/dist_check_n2.log | tail -3
for w in target c1 c3; do
timeout 600 python bench.py --workload $w --steps 100 --warmup 10 --no-cpu > gpurun_out/b1.json 2> gpurun_out/b1.err; python -c "
import json; d=json.load(open('gpurun_out/b1.json')); print('$w n1', round(d['value'],1), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), round(d['e2e']['ms_per_step'],4), d['e2e']['h2d_bytes_per_step'], d['e2e']['d2h_bytes_per_step'])"
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 2 --workload target --steps 100 --warmup 10 > gpurun_out/bench_target_n2.json 2> gpurun_out/bench_target_n2.err; echo "target n2 rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_target_n2.json')); print(round(d['value'],1), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), round(d['e2e']['ms_per_step'],4))"
