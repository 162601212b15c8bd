#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tests/dist_check.py > gpurun_out/dist_check_n2.log 2>&1; echo "dist_check rc=$?"; grep -v "^\*\|OMP_NUM" gpurun_out/dist_check_n2.log | tail -4
for w in target c3 c1 c4; do
timeout 600 python bench.py --workload $w --steps 50 --warmup 5 --no-cpu > gpurun_out/bench_${w}_n1.json 2> gpurun_out/bench_${w}_n1.err; python -c "
import json; d=json.load(open('gpurun_out/bench_${w}_n1.json')); print('$w n1', d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['phases_ms'], d['roofline']['frac'])"
done
for w in target c4; do for ex in auto nccl; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 2 --workload $w --steps 50 --warmup 5 --exchange $ex > gpurun_out/bench_${w}_n2_$ex.json 2> gpurun_out/bench_${w}_n2_$ex.err; echo "$w n2 $ex rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_${w}_n2_$ex.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['phases_ms'])"
done; done
