#!/bin/bash
# round 2: two co-resident CTAs per SM (ctas_per_sm = 2) vs one; large units when pipelined
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
run() { name=$1; shift; timeout 600 python bench.py --no-cpu "$@" > gpurun_out/r2v_$name.json 2> gpurun_out/r2v_$name.err; echo "$name rc=$? [$*]"; summ gpurun_out/r2v_$name.json; grep -a "Error\|error" gpurun_out/r2v_$name.err | head -2 | cut -c1-200; }
for t in 0,0,0,0,0 0,0,0,2,0 6,0,0,2,0 8,0,0,2,0 0,0,0,0,32 0,0,0,2,32; do
  run shard_$t --rows 1250000 --steps 200 --warmup 20 --tuning $t
done
for t in 0,0,0,0,0 0,0,0,2,0; do
  run target_$t --steps 40 --warmup 5 --tuning $t
  run c3_$t --workload c3 --steps 50 --warmup 5 --tuning $t
  run c4shard_$t --workload c4 --rows 1250000 --steps 200 --warmup 20 --tuning $t
done
