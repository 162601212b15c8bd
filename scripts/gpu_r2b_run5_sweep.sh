#!/bin/bash
# round 2, second session, run 5: K1 on bf16 rows — front-end / columns-per-slot / slots A/B on one box (bench.py flags only).
mkdir -p gpurun_out/r2b5
O=gpurun_out/r2b5
run() { name=$1; shift
  timeout 200 python bench.py --no-cpu --no-parity --steps 20 --warmup 5 --vector-format bf16 "$@" > $O/$name.json 2> $O/$name.err
  python - <<PY
import json
try:
    d=json.load(open('$O/$name.json')); r=d['roofline']
    print('%-28s value=%8.1f e2e=%8.1f blocking=%8.1f scan_ms=%.4f frac=%.3f' % ('$name', d['value'], d['e2e']['value'], d['e2e']['blocking_value'], r['scan_ms'], r['frac']))
except Exception as e:
    print('$name no line', e)
PY
}
run target_auto --workload target
run target_autonomous --workload target --scan-mode 1
run target_kc768 --workload target --tuning 0,0,768,0,0
run target_kc768_autonomous --workload target --tuning 0,0,768,0,0 --scan-mode 1
run target_kc256 --workload target --tuning 0,0,256,0,0
run target_autonomous_2slots --workload target --tuning 0,2,0,0,0 --scan-mode 1
run c5_auto --workload c5
run c5_kc768 --workload c5 --tuning 0,0,768,0,0
run c5_autonomous --workload c5 --scan-mode 1
run c3_auto --workload c3
run c3_2slots --workload c3 --tuning 0,2,0,0,0
run c3_planner --workload c3 --scan-mode 2
run c4_auto --workload c4
run c4_kc768 --workload c4 --tuning 0,0,768,0,0
