#!/bin/bash
# round 2, second session, run 4 (two GPUs): the sharded paths with the current library — dist_check (oracle parity of the fused
# peer exchange and the NCCL path), the target at N = 2 with fp32 and with bf16 rows (digest must equal the N = 1 digest).
mkdir -p gpurun_out/r2b4
O=gpurun_out/r2b4
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29541 tests/dist_check.py > $O/dist_check_n2.log 2>&1; echo "dist_check rc=$?"; grep -E "DIST_CHECK|Error|assert" $O/dist_check_n2.log | tail -3
timeout 300 $TR --master-port 29544 bench.py --gpus 2 --workload target --steps 20 --warmup 5 > $O/bench_target_n2.json 2> $O/bench_target_n2.err; echo "target n2 rc=$?"
timeout 300 $TR --master-port 29545 bench.py --gpus 2 --workload target --vector-format bf16 --steps 20 --warmup 5 > $O/bench_target_bf16_n2.json 2> $O/bench_target_bf16_n2.err; echo "target bf16 n2 rc=$?"
python - <<PY
import json
for f in ('bench_target_n2','bench_target_bf16_n2'):
    try:
        d=json.load(open('$O/%s.json'%f)); p=d['parity_check']
        print(f, 'value=%.1f e2e=%.1f blocking=%.1f parity=%s digest=%s %s' % (d['value'], d['e2e']['value'], d['e2e']['blocking_value'], p['ok'], p['digest'][:12], p['failures']))
    except Exception as e:
        print(f, 'no line', e)
PY
