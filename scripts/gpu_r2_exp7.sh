#!/bin/bash
# round 2: K0b v2 (row bitmask at HBM bandwidth) as the default predicate path — whole GPU suite, then bench lines
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest5.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_pytest5.log
summ() { python - "$1" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
except Exception as e:
    print("  (no json:", e, ")"); sys.exit(0)
r, e, p = d["roofline"], d["e2e"], d.get("parity_check") or {}
print("  value=%.1f q/s (%.4f ms) e2e=%.1f blocking=%.1f launches=%s roof=%.0f frac=%.3f phases=%s parity=%s %s" % (
    d["value"], d["ms_per_step"], e["value"], e["blocking_value"], d["gpu_launches"], r["achieved"], r["frac"], {k: round(v, 4) for k, v in d["phases_ms"].items()}, p.get("ok"), p.get("failures")))
PY
}
run() { name=$1; shift; timeout 600 python bench.py --no-cpu "$@" > gpurun_out/r2t_$name.json 2> gpurun_out/r2t_$name.err; echo "$name rc=$? [$*]"; summ gpurun_out/r2t_$name.json; grep -a "Error\|error" gpurun_out/r2t_$name.err | head -2 | cut -c1-200; }
run c3 --workload c3 --steps 50 --warmup 5
run c3_fused --workload c3 --steps 50 --warmup 5 --unfused-predicate 2
run target --steps 50 --warmup 5
run target_fused --steps 50 --warmup 5 --unfused-predicate 2
run shard --rows 1250000 --steps 200 --warmup 20
run shard_fused --rows 1250000 --steps 200 --warmup 20 --unfused-predicate 2
run c5 --workload c5 --steps 30 --warmup 5
run c5_fused --workload c5 --steps 30 --warmup 5 --unfused-predicate 2
