#!/bin/bash
# round 2: two unit ids claimed ahead in K1 — parity subset + bench lines
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_async.py tests/test_gpu_kats.py tests/test_gpu_planner.py -m gpu -x -q > gpurun_out/r2_pytest10.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest10.log
summ() { python - "$1" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
except Exception as e:
    print("  (no json:", e, ")"); sys.exit(0)
r, e, p = d["roofline"], d["e2e"], d.get("parity_check") or {}
print("  value=%.1f q/s (%.4f ms) e2e=%.1f blocking=%.1f roof=%.0f frac=%.3f scan_ms=%.4f parity=%s" % (d["value"], d["ms_per_step"], e["value"], e["blocking_value"], r["achieved"], r["frac"], r["scan_ms"], p.get("ok")))
PY
}
run() { name=$1; shift; timeout 600 python bench.py --no-cpu "$@" > gpurun_out/r2o_$name.json 2> gpurun_out/r2o_$name.err; echo "$name rc=$? [$*]"; summ gpurun_out/r2o_$name.json; grep -a "Error\|error" gpurun_out/r2o_$name.err | head -2 | cut -c1-200; }
run c3 --workload c3 --steps 50 --warmup 5
run c3u --workload c3u --steps 50 --warmup 5
run target --steps 40 --warmup 5
run shard --rows 1250000 --steps 200 --warmup 20
run c4 --workload c4 --steps 30 --warmup 5
run c1 --workload c1 --steps 300 --warmup 30
