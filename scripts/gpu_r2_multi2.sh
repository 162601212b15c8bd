#!/bin/bash
# round 2, N GPUs (N = $1): dist_check, then the target and C4 lines (default path) — scaling evidence
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29733"
timeout 900 $TR tests/dist_check.py > gpurun_out/r2_dist_check_n$N.log 2>&1; echo "dist_check rc=$?"; grep -a "DIST_CHECK_OK\|Error\|error" gpurun_out/r2_dist_check_n$N.log | head -5
summ() { python - "$1" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
except Exception as e:
    print("  (no json:", e, ")"); sys.exit(0)
r, e, p = d["roofline"], d["e2e"], d.get("parity_check") or {}
print("  value=%.1f q/s (%.4f ms) e2e=%.1f blocking=%.1f launches=%s roof=%.0f frac=%.3f phases=%s parity=%s %s digest=%s clocks=%s" % (
    d["value"], d["ms_per_step"], e["value"], e["blocking_value"], d["gpu_launches"], r["achieved"], r["frac"], {k: round(v, 4) for k, v in d["phases_ms"].items()},
    p.get("ok"), p.get("failures"), (p.get("digest") or "")[:12], d["clocks"]["reasons"] if d.get("clocks") else None))
PY
}
run() { name=$1; shift; timeout 900 $TR bench.py --gpus $N "$@" > gpurun_out/r2_$name.json 2> gpurun_out/r2_$name.err; echo "$name rc=$?"; summ gpurun_out/r2_$name.json; grep -a "Error\|error" gpurun_out/r2_$name.err | head -3 | cut -c1-300; }
run v2_target_n$N --steps 200 --warmup 20
run v2_target_n${N}_cps2 --steps 200 --warmup 20 --tuning 0,0,0,2,0
run v2_c4_n$N --workload c4 --steps 100 --warmup 10
run v2_c3_n$N --workload c3 --steps 200 --warmup 20
run v2_c5_n$N --workload c5 --steps 100 --warmup 10
