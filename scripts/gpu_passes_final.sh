#!/bin/bash
# batched tests again after the test fixes, then one ncu capture of the single-pass batch_kernel on c2
mkdir -p gpurun_out
timeout 170 python -m pytest tests/test_gpu_batched.py -q > gpurun_out/pytest_batched.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/pytest_batched.log
timeout 150 ncu --set full --clock-control none --import-source on -k regex:batch_kernel -s 2 -c 1 -o gpurun_out/batch_c2_1x -f python bench.py --workload c2 --steps 2 --warmup 2 --no-cpu > gpurun_out/ncu_c2_1x.log 2>&1
echo "ncu rc=$?"
