#!/bin/bash
# round 2 experiments on one GPU: what bounds K1 on narrow rows (C3) and on shard-sized stores (1.25M x 768 = one 8-way shard)?
mkdir -p gpurun_out
summ() { python - "$1" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
except Exception as e:
    print("  (no json:", e, ")"); sys.exit(0)
r, e, p = d["roofline"], d["e2e"], d.get("parity_check") or {}
print("  value=%.1f q/s (%.4f ms) e2e=%.1f blocking=%.1f roof=%.0f frac=%.3f scan_ms=%.4f parity=%s" % (
    d["value"], d["ms_per_step"], e["value"], e["blocking_value"], r["achieved"], r["frac"], r["scan_ms"], p.get("ok")))
PY
}
run() { name=$1; shift; timeout 600 python bench.py --no-cpu "$@" > gpurun_out/r2x_$name.json 2> gpurun_out/r2x_$name.err; echo "$name rc=$? [$*]"; summ gpurun_out/r2x_$name.json; grep -a "Error\|error" gpurun_out/r2x_$name.err | head -2 | cut -c1-200; }
run c3u --workload c3u --steps 50 --warmup 5
for t in 0,0,0,0,64 0,0,0,0,32 12,2,0,0,0 8,2,0,0,0 16,1,64,0,0; do run c3_t$t --workload c3 --steps 50 --warmup 5 --tuning $t; done
for t in 0,0,0,0,0 0,0,0,0,128 0,0,0,0,64; do run shard_t$t --rows 1250000 --steps 200 --warmup 20 --tuning $t; done
run shard_blocking --rows 1250000 --steps 200 --warmup 20 --blocking
timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan_kernel -s 3 -c 1 -o gpurun_out/r2_scan_c3 -f python bench.py --workload c3 --steps 2 --warmup 2 --no-cpu --no-parity --blocking > gpurun_out/ncu_c3.log 2>&1; echo "ncu c3 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan_kernel -s 3 -c 1 -o gpurun_out/r2_scan_shard -f python bench.py --rows 1250000 --steps 2 --warmup 2 --no-cpu --no-parity --blocking > gpurun_out/ncu_shard.log 2>&1; echo "ncu shard rc=$?"
