#!/bin/bash
# round-2 record run on N GPUs (N = $1): dist_check under torchrun, target (+ C4, C5 at N = 8) bench lines, reference arm under torchrun
N=${1:-8}
mkdir -p gpurun_out/r2_bench
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29733"
timeout 900 $TR tests/dist_check.py > gpurun_out/r2_dist_check_n$N.log 2>&1; echo "dist_check rc=$?"; grep -a "DIST_CHECK_OK\|Error\|error" gpurun_out/r2_dist_check_n$N.log | head -5
summ() { python - "$1" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
except Exception as e:
    print("  (no json:", e, ")"); sys.exit(0)
r, e, p = d["roofline"], d["e2e"], d.get("parity_check") or {}
print("  value=%.1f q/s (%.4f ms) e2e=%.1f blocking=%.1f launches=%s frac=%.3f phases=%s parity=%s %s digest=%s clocks=%s" % (
    d["value"], d["ms_per_step"], e["value"], e["blocking_value"], d["gpu_launches"], r["frac"], {k: round(v, 4) for k, v in d["phases_ms"].items()},
    p.get("ok"), p.get("failures"), (p.get("digest") or "")[:12], d["clocks"]["reasons"] if d.get("clocks") else None))
PY
}
run() { name=$1; shift; timeout 900 $TR bench.py --gpus $N "$@" > gpurun_out/r2_bench/$name.json 2> gpurun_out/r2_bench/$name.err; echo "$name rc=$?"; summ gpurun_out/r2_bench/$name.json; grep -a "Error\|error" gpurun_out/r2_bench/$name.err | head -3 | cut -c1-300; }
run bench_target_n$N --steps 20 --warmup 5
run bench_target_n${N}_200 --steps 200 --warmup 20
if [ "$N" = "8" ]; then
  run bench_c4_n$N --workload c4 --steps 100 --warmup 10
  run bench_c5_n$N --workload c5 --steps 100 --warmup 10
  run bench_target_n${N}_nccl --steps 200 --warmup 20 --exchange nccl
  timeout 600 $TR bench.py --gpus $N --impl reference --steps 20 --warmup 5 > gpurun_out/r2_bench/bench_ref_n$N.json 2> /dev/null; echo "ref rc=$?"
  python -c "
import json; d=json.load(open('gpurun_out/r2_bench/bench_ref_n$N.json')); print('reference arm under torchrun: value=%.2f q/s cores=%d' % (d['value'], d['cpu_baseline']['cores']))"
fi
