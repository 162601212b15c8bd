#!/bin/bash
# round 2, second session, run 2: K2 with eight epilogue warps (quick check first; falls back to four for the rest of the run
# if that check fails), persistence tests, the whole GPU suite, K2 A/B lines (epilogue warps, k-blocks per stage), ncu of K2.
mkdir -p gpurun_out/r2b2 /tmp/rep
O=gpurun_out/r2b2
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/gpu.txt
timeout 200 python -m pytest tests/test_gpu_batched.py -q -x -k "many_queries or separated or special or vec_filter" > $O/pytest_k2_quick.log 2>&1; rc=$?
echo "k2 quick (8 epilogue warps) rc=$rc"; tail -3 $O/pytest_k2_quick.log
if [ $rc -ne 0 ]; then export OTTERS_K2_EPI_WARPS=4; echo "FALLING BACK to 4 epilogue warps for the rest of this run"; fi
timeout 300 python -m pytest tests/test_gpu_persist.py tests/test_gpu_bf16_store.py tests/test_gpu_reorder.py -q > $O/pytest_new.log 2>&1; echo "new tests rc=$?"; tail -3 $O/pytest_new.log
one() { name=$1; shift
  timeout 300 python bench.py "$@" > $O/bench_$name.json 2> $O/bench_$name.err; echo "bench $name rc=$?"
  python - <<PY
import json
try:
    d=json.load(open('$O/bench_$name.json')); r=d['roofline']; p=d.get('parity_check') or {}
    print('  value=%.1f step=%.4f ms e2e=%.1f blocking=%.1f roof=%.1f %s frac=%.3f scan_ms=%s parity=%s %s passes=%s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['blocking_value'], r['achieved'], r['unit'], r['frac'], r.get('scan_ms'), p.get('ok'), p.get('failures'), r.get('mma_passes')))
except Exception as e:
    print('  no line:', e)
PY
}
one c2 --workload c2 --steps 20 --warmup 5 --no-cpu
OTTERS_K2_EPI_WARPS=4 one c2_epi4 --workload c2 --steps 20 --warmup 5 --no-cpu --no-parity
OTTERS_K2_KPS=2 one c2_kps2 --workload c2 --steps 20 --warmup 5 --no-cpu --no-parity
OTTERS_K2_KPS=1 one c2_kps1 --workload c2 --steps 20 --warmup 5 --no-cpu --no-parity
one c2_tf32 --workload c2 --steps 20 --warmup 5 --no-cpu --no-parity --batch-passes 1
one c2_3x --workload c2 --steps 10 --warmup 3 --no-cpu --no-parity --batch-passes 3
timeout 900 python -m pytest tests -m gpu -q > $O/pytest_full.log 2>&1; echo "full suite rc=$?"; tail -4 $O/pytest_full.log
B="python bench.py --no-cpu --no-parity --blocking --steps 3 --warmup 2"
cap() { name=$1; kern=$2; shift 2
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$kern -s 3 -c 1 -o /tmp/rep/r2b_$name -f $B "$@" > $O/ncu_$name.log 2>&1; echo "ncu $name rc=$?"
  python scripts/ncu_summary.py /tmp/rep/r2b_$name.ncu-rep > $O/r2b_${name}_summary.txt 2>&1
  python scripts/ncu_hot.py /tmp/rep/r2b_$name.ncu-rep 40 > $O/r2b_${name}_hot_sass.txt 2>&1
}
cap batch_c2_bf16_epi8 batch_kernel --workload c2
du -sh gpurun_out
