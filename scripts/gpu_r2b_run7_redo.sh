#!/bin/bash
# round 2, second session, run 7: K2 lean epilogue — chunks that hold a candidate redone column by column (OTTERS_K2_REDO=0) or
# on the general straight-line path (=1): same-box A/B on c2, then the unfiltered batched tests with =1.
mkdir -p gpurun_out/r2b7
O=gpurun_out/r2b7
run() { name=$1; shift
  timeout 200 python bench.py --workload c2 --no-cpu --steps 20 --warmup 5 "$@" > $O/$name.json 2> $O/$name.err
  python - <<PY
import json
try:
    d=json.load(open('$O/$name.json')); r=d['roofline']; p=d.get('parity_check') or {}
    print('%-14s value=%9.1f step=%.4f scan_ms=%.4f frac=%.3f parity=%s passes=%s fallbacks=%s' % ('$name', d['value'], d['ms_per_step'], r['scan_ms'], r['frac'], p.get('ok'), r.get('mma_passes'), r.get('fallbacks')))
except Exception as e:
    print('$name no line', e)
PY
}
OTTERS_K2_REDO=0 run c2_redo0
OTTERS_K2_REDO=1 run c2_redo1
OTTERS_K2_REDO=0 run c2_redo0_b --no-parity
OTTERS_K2_REDO=1 run c2_redo1_b --no-parity
OTTERS_K2_REDO=1 timeout 280 python -m pytest tests/test_gpu_batched.py -q -x -k "not vec_filter and not few_survivors and not cta_pair_3x and not single_cta_3x" > $O/pytest_batched_redo1.log 2>&1; echo "batched (redo=1) rc=$?"; tail -3 $O/pytest_batched_redo1.log
