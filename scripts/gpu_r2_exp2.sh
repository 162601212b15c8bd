#!/bin/bash
# round 2: row-list prefetch in K1 — parity subset, then C3 / shard-sized / target lines
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_async.py tests/test_gpu_kats.py tests/test_gpu_planner.py -m gpu -x -q > gpurun_out/r2_pytest3.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest3.log
summ() { python - "$1" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
except Exception as e:
    print("  (no json:", e, ")"); sys.exit(0)
r, e, p = d["roofline"], d["e2e"], d.get("parity_check") or {}
print("  value=%.1f q/s (%.4f ms) e2e=%.1f blocking=%.1f roof=%.0f frac=%.3f scan_ms=%.4f parity=%s %s" % (
    d["value"], d["ms_per_step"], e["value"], e["blocking_value"], r["achieved"], r["frac"], r["scan_ms"], p.get("ok"), p.get("failures")))
PY
}
run() { name=$1; shift; timeout 600 python bench.py --no-cpu "$@" > gpurun_out/r2y_$name.json 2> gpurun_out/r2y_$name.err; echo "$name rc=$? [$*]"; summ gpurun_out/r2y_$name.json; grep -a "Error\|error" gpurun_out/r2y_$name.err | head -2 | cut -c1-200; }
run c3 --workload c3 --steps 50 --warmup 5
run c3_u64 --workload c3 --steps 50 --warmup 5 --tuning 0,0,0,0,64
run target --steps 50 --warmup 5
for t in 0,0,0,0,0 0,0,0,0,128 0,0,0,0,64; do run shard_t$t --rows 1250000 --steps 200 --warmup 20 --tuning $t; done
run c5 --workload c5 --steps 30 --warmup 5
run c1 --workload c1 --steps 300 --warmup 30
