#!/bin/bash
# A/B of the K2 k-block size: 32 fp32 columns (128-byte swizzle, 2 stages) vs 16 (64-byte swizzle, 4 stages)
mkdir -p gpurun_out
cp otters_b200/libotters_b200.so /tmp/lib_orig.so
for v in bk16 bk32; do
  cp scripts/tmp/lib_$v.so otters_b200/libotters_b200.so
  echo "== $v"
  timeout 600 python -m pytest tests/test_gpu_batched.py -x -q 2>&1 | tail -2
  timeout 300 python scripts/dbg_c2_time.py 2>&1 | grep "dbg= 0\|dbg= 4"
done
cp /tmp/lib_orig.so otters_b200/libotters_b200.so
