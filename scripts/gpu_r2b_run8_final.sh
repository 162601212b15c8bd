#!/bin/bash
# round 2, second session, run 8: the test files touched after the record run (oracle-side bf16 rounding), smoke, default bench.
mkdir -p gpurun_out/r2b8
O=gpurun_out/r2b8
timeout 200 python -m pytest tests/test_gpu_bf16_store.py tests/test_gpu_reorder.py tests/test_gpu_persist.py tests/test_sharded.py -m gpu -q > $O/pytest_touched.log 2>&1; echo "touched tests rc=$?"; tail -2 $O/pytest_touched.log
timeout 100 python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 200 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "default bench rc=$?"
python -c "
import json; d=json.load(open('$O/bench_default.json')); print('value=%.1f e2e=%.1f frac=%.3f parity=%s cpu=%.2f' % (d['value'], d['e2e']['value'], d['roofline']['frac'], d['parity_check']['ok'], d['cpu_baseline']['value']))"
