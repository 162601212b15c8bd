#!/bin/bash
# round 2: same-box A/B of the K1 producer variants (env switches): predicate form, row-list prefetch, shared-memory margin
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
run() { name=$1; shift; timeout 600 python bench.py --no-cpu --no-parity "$@" > gpurun_out/r2w_$name.json 2> gpurun_out/r2w_$name.err; echo "$name rc=$? [$*] PRED_SEQ=$OTTERS_PRED_SEQ NO_PREFETCH=$OTTERS_NO_PREFETCH MARGIN=$OTTERS_SMEM_MARGIN"; summ gpurun_out/r2w_$name.json; grep -a "Error\|error" gpurun_out/r2w_$name.err | head -2 | cut -c1-200; }
for cfg in "0 0 128" "1 0 128" "1 1 128" "0 1 128" "1 0 1024" "1 1 1024"; do
  set -- $cfg; export OTTERS_PRED_SEQ=$1 OTTERS_NO_PREFETCH=$2 OTTERS_SMEM_MARGIN=$3
  run c3_$1$2_$3 --workload c3 --steps 50 --warmup 5
  run target_$1$2_$3 --steps 40 --warmup 5
  run shard_$1$2_$3 --rows 1250000 --steps 200 --warmup 20
done
