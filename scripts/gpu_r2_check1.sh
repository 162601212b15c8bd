#!/bin/bash
# round 2, first GPU check: the whole GPU suite on the one-launch query path, then A/B bench lines of the target
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest1.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2_pytest1.log
summ() { python - "$1" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
except Exception as e:
    print("  (no json:", e, ")"); sys.exit(0)
r, e, p = d["roofline"], d["e2e"], d.get("parity_check") or {}
print("  value=%.1f q/s (%.4f ms) e2e=%.1f blocking=%.1f launches=%s roof=%.0f %s frac=%.3f scan_ms=%s parity=%s %s cpu=%s clocks=%s" % (
    d["value"], d["ms_per_step"], e["value"], e["blocking_value"], d["gpu_launches"], r["achieved"], r["unit"], r["frac"], r["scan_ms"],
    p.get("ok"), p.get("failures"), (d.get("cpu_baseline") or {}).get("value"), d["clocks"]["reasons"] if d.get("clocks") else None))
PY
}
run() { name=$1; shift; timeout 900 python bench.py "$@" > gpurun_out/r2_$name.json 2> gpurun_out/r2_$name.err; echo "$name rc=$?"; summ gpurun_out/r2_$name.json; tail -3 gpurun_out/r2_$name.err | cut -c1-300; }
run target --steps 50 --warmup 5
run target_sepsel --steps 50 --warmup 5 --no-cpu --separate-select 1
run target_sepboth --steps 50 --warmup 5 --no-cpu --separate-select 1 --separate-prune 1 --blocking
run c3 --workload c3 --steps 50 --warmup 5
run c1 --workload c1 --steps 200 --warmup 20
run c5 --workload c5 --steps 30 --warmup 5
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -1
