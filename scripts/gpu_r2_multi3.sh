#!/bin/bash
# round 2, 2 GPUs: large-k exchange merge (C5 shard size of the 8-GPU run), dist_check
N=2
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
print("  value=%.1f q/s (%.4f ms) e2e=%.1f blocking=%.1f launches=%s roof=%.0f frac=%.3f phases=%s parity=%s %s" % (
    d["value"], d["ms_per_step"], e["value"], e["blocking_value"], d["gpu_launches"], r["achieved"], r["frac"], {k: round(v, 4) for k, v in d["phases_ms"].items()},
    p.get("ok"), p.get("failures")))
PY
}
run() { name=$1; shift; timeout 900 $TR bench.py --gpus $N "$@" > gpurun_out/r2_$name.json 2> gpurun_out/r2_$name.err; echo "$name rc=$?"; summ gpurun_out/r2_$name.json; grep -a "Error\|error" gpurun_out/r2_$name.err | head -3 | cut -c1-300; }
run v3_c5_shardsize_n2 --workload c5 --rows 1250000 --steps 100 --warmup 10
run v3_c5_n2 --workload c5 --steps 50 --warmup 5
run v3_target_n2 --steps 100 --warmup 10
