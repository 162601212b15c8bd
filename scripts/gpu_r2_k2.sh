#!/bin/bash
# round 2: K2 single-pass variants (k-blocks per stage, CTA pairs with direct barrier credit) + K1 slot variants on narrow rows
mkdir -p gpurun_out
for cfg in "1 1 0" "1 2 0" "2 1 0" "2 2 0" "2 1 1" "2 2 1"; do
  set -- $cfg
  OTTERS_K2_KPS=$2 OTTERS_K2_DIRECT=$3 timeout 180 python scripts/dbg_k2_variants.py $1 2>&1 | tail -2
done | tee gpurun_out/r2_k2_variants.log
summ() { python - "$1" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
except Exception as e:
    print("  (no json:", e, ")"); sys.exit(0)
r, e, p = d["roofline"], d["e2e"], d.get("parity_check") or {}
print("  value=%.1f q/s (%.4f ms) blocking=%.1f roof=%.0f frac=%.3f scan_ms=%.4f parity=%s" % (d["value"], d["ms_per_step"], e["blocking_value"], r["achieved"], r["frac"], r["scan_ms"], p.get("ok")))
PY
}
run() { name=$1; shift; timeout 600 python bench.py --no-cpu "$@" > gpurun_out/r2q_$name.json 2> gpurun_out/r2q_$name.err; echo "$name rc=$? [$*]"; summ gpurun_out/r2q_$name.json; grep -a "Error\|error" gpurun_out/r2q_$name.err | head -2 | cut -c1-200; }
for t in 0,0,0,0,0 12,2,0,0,0 8,3,0,0,0 10,2,0,0,0; do
  run c3u_$t --workload c3u --steps 50 --warmup 5 --tuning $t
  run c3_$t --workload c3 --steps 50 --warmup 5 --tuning $t
done
run c1 --workload c1 --steps 300 --warmup 30
