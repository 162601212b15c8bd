#!/bin/bash
# round 2: whole GPU suite again (prune kernel back as its own launch by default), C3 / target A/B of lazy pruning
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest2.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_pytest2.log
summ() { python - "$1" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
except Exception as e:
    print("  (no json:", e, ")"); sys.exit(0)
r, e, p = d["roofline"], d["e2e"], d.get("parity_check") or {}
print("  value=%.1f q/s (%.4f ms) e2e=%.1f blocking=%.1f launches=%s roof=%.0f %s frac=%.3f scan_ms=%s phases=%s parity=%s %s cpu=%s clocks=%s" % (
    d["value"], d["ms_per_step"], e["value"], e["blocking_value"], d["gpu_launches"], r["achieved"], r["unit"], r["frac"], r["scan_ms"], d.get("phases_ms"),
    p.get("ok"), p.get("failures"), (d.get("cpu_baseline") or {}).get("value"), d["clocks"]["reasons"] if d.get("clocks") else None))
PY
}
run() { name=$1; shift; timeout 900 python bench.py "$@" > gpurun_out/r2_$name.json 2> gpurun_out/r2_$name.err; echo "$name rc=$?"; summ gpurun_out/r2_$name.json; tail -3 gpurun_out/r2_$name.err | cut -c1-300; }
run b_target --steps 50 --warmup 5 --no-cpu
run b_target_lazy --steps 50 --warmup 5 --no-cpu --lazy-prune 1
run b_c3 --workload c3 --steps 50 --warmup 5 --no-cpu
run b_c3_lazy --workload c3 --steps 50 --warmup 5 --no-cpu --lazy-prune 1
run b_c3_sepsel --workload c3 --steps 50 --warmup 5 --no-cpu --separate-select 1
run b_c4 --workload c4 --steps 30 --warmup 5 --no-cpu
run b_c2 --workload c2 --steps 20 --warmup 3 --no-cpu
