#!/bin/bash
# round 2, second session, run 1: the new paths first (bf16 rung of K2, bf16 row stores, row ordering), then bench lines,
# the whole GPU suite and two ncu captures.  Everything is bounded by its own timeout.
mkdir -p gpurun_out/r2b /tmp/rep
O=gpurun_out/r2b
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/gpu.txt
timeout 400 python -m pytest tests/test_gpu_bf16_store.py tests/test_gpu_reorder.py -q > $O/pytest_new.log 2>&1; echo "new tests rc=$?"; tail -3 $O/pytest_new.log
timeout 600 python -m pytest tests/test_gpu_batched.py -q > $O/pytest_batched.log 2>&1; echo "batched rc=$?"; tail -3 $O/pytest_batched.log
one() { name=$1; shift
  timeout 400 python bench.py "$@" > $O/bench_$name.json 2> $O/bench_$name.err; echo "bench $name rc=$?"
  python - <<PY
import json
try:
    d=json.load(open('$O/bench_$name.json')); r=d['roofline']; p=d.get('parity_check') or {}
    print('  value=%.1f step=%.4f ms e2e=%.1f blocking=%.1f roof=%.1f %s frac=%.3f scan_ms=%s parity=%s %s passes=%s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['blocking_value'], r['achieved'], r['unit'], r['frac'], r.get('scan_ms'), p.get('ok'), p.get('failures'), r.get('mma_passes')))
except Exception as e:
    print('  no line:', e)
PY
}
one c2 --workload c2 --steps 20 --warmup 5
one c2_tf32 --workload c2 --steps 20 --warmup 5 --batch-passes 1 --no-cpu
one target_bf16 --workload target --vector-format bf16 --steps 20 --warmup 5
one target --workload target --steps 20 --warmup 5 --no-cpu
one c4_bf16 --workload c4 --vector-format bf16 --steps 20 --warmup 5 --no-cpu
one c3_bf16 --workload c3 --vector-format bf16 --steps 20 --warmup 5 --no-cpu
timeout 900 python -m pytest tests -m gpu -q -x > $O/pytest_full.log 2>&1; echo "full suite rc=$?"; tail -3 $O/pytest_full.log
B="python bench.py --no-cpu --no-parity --blocking --steps 3 --warmup 2"
cap() { name=$1; kern=$2; shift 2
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$kern -s 3 -c 1 -o /tmp/rep/r2b_$name -f $B "$@" > $O/ncu_$name.log 2>&1; echo "ncu $name rc=$?"
  python scripts/ncu_summary.py /tmp/rep/r2b_$name.ncu-rep > $O/r2b_${name}_summary.txt 2>&1
  python scripts/ncu_hot.py /tmp/rep/r2b_$name.ncu-rep 40 > $O/r2b_${name}_hot_sass.txt 2>&1
}
cap batch_c2_bf16 batch_kernel --workload c2
cap scan_target_bf16 scan_planner_kernel --workload target --vector-format bf16
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/r2b_launches_c2.csv $B --workload c2 > $O/ncu_l_c2.log 2>&1; echo "launch list c2 rc=$?"
timeout 120 python __graft_entry__.py --smoke 2>&1 | tail -1
du -sh gpurun_out
