#!/bin/bash
# round 2: full GPU suite (per-query batches, device gather, device-side build), then build-time A/B on the target's 10M-row metadata
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest7.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2_pytest7.log
for h in 0 1; do
  OTTERS_BUILD_HOST=$h timeout 900 python bench.py --no-cpu --no-parity --steps 5 --warmup 2 > gpurun_out/r2_build_host$h.json 2> gpurun_out/r2_build_host$h.err
  python -c "
import json; d=json.load(open('gpurun_out/r2_build_host$h.json')); print('OTTERS_BUILD_HOST=$h store_build_s=%.2f value=%.1f' % (d['store_build_s'], d['value']))"
done
