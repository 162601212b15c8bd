#!/bin/bash
# round 2, second session, run 3 (record run at HEAD): new K2-on-bf16-store test first, the whole GPU suite, bench lines.
mkdir -p gpurun_out/r2b3
O=gpurun_out/r2b3
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/gpu.txt
timeout 300 python -m pytest tests/test_gpu_bf16_store.py -q > $O/pytest_bf16.log 2>&1; echo "bf16 store tests rc=$?"; tail -3 $O/pytest_bf16.log
timeout 900 python -m pytest tests -m gpu -q -x > $O/pytest_full.log 2>&1; echo "full suite rc=$?"; tail -4 $O/pytest_full.log
one() { name=$1; shift
  timeout 300 python bench.py "$@" > $O/bench_$name.json 2> $O/bench_$name.err; echo "bench $name rc=$?"
  python - <<PY
import json
try:
    d=json.load(open('$O/bench_$name.json')); r=d['roofline']; p=d.get('parity_check') or {}; c=d.get('cpu_baseline') or {}
    print('  value=%.1f step=%.4f ms e2e=%.1f blocking=%.1f roof=%.1f %s frac=%.3f scan_ms=%s parity=%s %s passes=%s cpu=%s clocks=%s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['blocking_value'], r['achieved'], r['unit'], r['frac'], r.get('scan_ms'), p.get('ok'), p.get('failures'), r.get('mma_passes'), c.get('value'), d['clocks'].get('reasons')))
except Exception as e:
    print('  no line:', e)
PY
}
one target --workload target --steps 20 --warmup 5
one c2 --workload c2 --steps 20 --warmup 5
one c2_bf16store --workload c2 --vector-format bf16 --steps 20 --warmup 5 --no-cpu
one target_bf16 --workload target --vector-format bf16 --steps 20 --warmup 5 --no-cpu
one c3 --workload c3 --steps 20 --warmup 5 --no-cpu
one c1 --workload c1 --steps 50 --warmup 5 --no-cpu
one c5 --workload c5 --steps 20 --warmup 5 --no-cpu
one c5_bf16 --workload c5 --vector-format bf16 --steps 20 --warmup 5 --no-cpu
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_ref.json 2> $O/bench_ref.err; echo "ref rc=$?"; python -c "
import json; d=json.load(open('$O/bench_ref.json')); print('  ref value=%.2f q/s cores=%d' % (d['value'], d['cpu_baseline']['cores']))"
timeout 120 python __graft_entry__.py --smoke 2>&1 | tail -1
du -sh gpurun_out
