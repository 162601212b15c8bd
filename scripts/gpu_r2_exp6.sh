#!/bin/bash
# round 2: row predicate as its own kernel (K0b bitmask) vs fused into the scan's producer; shared-memory headroom for co-resident K0/K0b
mkdir -p gpurun_out
summ() { python - "$1" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
except Exception as e:
    print("  (no json:", e, ")"); sys.exit(0)
r, e, p = d["roofline"], d["e2e"], d.get("parity_check") or {}
print("  value=%.1f q/s (%.4f ms) e2e=%.1f blocking=%.1f roof=%.0f frac=%.3f phases=%s parity=%s" % (
    d["value"], d["ms_per_step"], e["value"], e["blocking_value"], r["achieved"], r["frac"], {k: round(v, 4) for k, v in d["phases_ms"].items()}, p.get("ok")))
PY
}
run() { name=$1; shift; timeout 600 python bench.py --no-cpu "$@" > gpurun_out/r2u_$name.json 2> gpurun_out/r2u_$name.err; echo "$name rc=$? [$*] MARGIN=$OTTERS_SMEM_MARGIN"; summ gpurun_out/r2u_$name.json; grep -a "Error\|error" gpurun_out/r2u_$name.err | head -2 | cut -c1-200; }
for m in 128 8192; do
  export OTTERS_SMEM_MARGIN=$m
  run c3_fused_$m --workload c3 --steps 50 --warmup 5
  run c3_unfused_$m --workload c3 --steps 50 --warmup 5 --unfused-predicate 1
  run c3_unfused_pl_$m --workload c3 --steps 50 --warmup 5 --unfused-predicate 1 --scan-mode 2
  run shard_fused_$m --rows 1250000 --steps 200 --warmup 20
  run shard_unfused_$m --rows 1250000 --steps 200 --warmup 20 --unfused-predicate 1
  run shard_unfused_auto1_$m --rows 1250000 --steps 200 --warmup 20 --unfused-predicate 1 --scan-mode 1
  run target_fused_$m --steps 40 --warmup 5
  run target_unfused_$m --steps 40 --warmup 5 --unfused-predicate 1
  run target_unfused_auto1_$m --steps 40 --warmup 5 --unfused-predicate 1 --scan-mode 1
done
