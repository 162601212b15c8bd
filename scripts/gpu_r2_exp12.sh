#!/bin/bash
# round 2: unit-claim prefetch depth A/B on one box (OTTERS_CLAIM_DEPTH = 1, 2, 4; unset = automatic)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_async.py tests/test_gpu_planner.py -m gpu -x -q > gpurun_out/r2_pytest11.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2_pytest11.log
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
run() { name=$1; shift; timeout 600 python bench.py --no-cpu "$@" > gpurun_out/r2n_$name.json 2> gpurun_out/r2n_$name.err; echo "$name rc=$? [$*] DEPTH=$OTTERS_CLAIM_DEPTH"; summ gpurun_out/r2n_$name.json; grep -a "Error\|error" gpurun_out/r2n_$name.err | head -2 | cut -c1-200; }
for d in 1 2 4; do
  export OTTERS_CLAIM_DEPTH=$d
  run c3_d$d --workload c3 --steps 50 --warmup 5
  run c3u_d$d --workload c3u --steps 50 --warmup 5
  run target_d$d --steps 40 --warmup 5
  run shard_d$d --rows 1250000 --steps 200 --warmup 20
  run c1_d$d --workload c1 --steps 300 --warmup 30
done
