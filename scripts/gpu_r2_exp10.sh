#!/bin/bash
# round 2: dead-chunk skipping in K1 (atomicMax on the unit counter) — parity subset + A/B on the same box
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_async.py tests/test_gpu_kats.py tests/test_gpu_planner.py tests/test_gpu_per_query.py "tests/test_gpu_fullsize.py::test_fullsize_metastore_c3" -m gpu -x -q > gpurun_out/r2_pytest9.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest9.log
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
run() { name=$1; shift; timeout 600 python bench.py --no-cpu "$@" > gpurun_out/r2p_$name.json 2> gpurun_out/r2p_$name.err; echo "$name rc=$? [$*] NO_SKIP=$OTTERS_NO_SKIP"; summ gpurun_out/r2p_$name.json; grep -a "Error\|error" gpurun_out/r2p_$name.err | head -2 | cut -c1-200; }
for ns in "" 1; do
  if [ -n "$ns" ]; then export OTTERS_NO_SKIP=1; else unset OTTERS_NO_SKIP; fi
  run c3_skip$ns --workload c3 --steps 50 --warmup 5
  run target_skip$ns --steps 40 --warmup 5
  run shard_skip$ns --rows 1250000 --steps 200 --warmup 20
  run c5_skip$ns --workload c5 --steps 30 --warmup 5
done
