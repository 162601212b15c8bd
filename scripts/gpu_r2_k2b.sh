#!/bin/bash
# round 2: K2 defaults (single pass = CTA pairs, direct barrier credit, 2 k-blocks per stage) — batched tests + more variants
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_batched.py tests/test_gpu_fullsize.py::test_fullsize_batched_config2 -m gpu -x -q > gpurun_out/r2_pytest8.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest8.log
for cfg in "1 1 0" "2 2 1" "2 3 1" "1 2 0" "1 3 0"; do
  set -- $cfg
  OTTERS_K2_KPS=$2 OTTERS_K2_DIRECT=$3 timeout 180 python scripts/dbg_k2_variants.py $1 2>&1 | tail -1
done | tee gpurun_out/r2_k2_variants2.log
timeout 600 python bench.py --no-cpu --workload c2 --steps 20 --warmup 3 > gpurun_out/r2q_c2.json 2> gpurun_out/r2q_c2.err; python -c "
import json; d=json.load(open('gpurun_out/r2q_c2.json')); r=d['roofline']; print('c2 value=%.0f q/s step=%.3f ms kernel=%.3f ms frac=%.3f used=%s fallbacks=%s parity=%s' % (d['value'], d['ms_per_step'], r['scan_ms'], r['frac'], r['tensor_path_used'], r['fallbacks'], d['parity_check']['ok']))"
