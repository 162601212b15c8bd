#!/bin/bash
# round 2, second session, run 10: K1 split into two translation units (fp32 / bf16 rows) — same kernels, sanity + speed check.
mkdir -p gpurun_out/r2b10
O=gpurun_out/r2b10
timeout 100 python -m pytest tests/test_gpu_parity.py tests/test_gpu_bf16_store.py tests/test_gpu_kats.py -q -x > $O/pytest.log 2>&1; echo "parity tests rc=$?"; tail -2 $O/pytest.log
timeout 100 python bench.py --no-cpu --steps 20 --warmup 5 > $O/bench_default.json 2> $O/bench_default.err; echo "default bench rc=$?"
python -c "
import json; d=json.load(open('$O/bench_default.json')); print('value=%.1f e2e=%.1f frac=%.3f parity=%s' % (d['value'], d['e2e']['value'], d['roofline']['frac'], d['parity_check']['ok']))"
