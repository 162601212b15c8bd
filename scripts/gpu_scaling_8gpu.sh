#!/bin/bash
# 8-GPU box: strong-scaling points of the target workload and c4 (fused exchange)
mkdir -p gpurun_out
run() { # n workload exchange
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $1 --workload $2 --steps 100 --warmup 10 --exchange $3 > gpurun_out/bench_$2_n$1_$3.json 2> gpurun_out/bench_$2_n$1_$3.err; echo "$2 n$1 $3 rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_$2_n$1_$3.json')); print('value', round(d['value'],1), 'step_ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), 'e2e_ms', round(d['e2e']['ms_per_step'],4), d['phases_ms'], d['config']['parallelism'][-40:])"
}
run 8 target auto
run 8 target nccl
run 4 target auto
run 8 c4 auto
