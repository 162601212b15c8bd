#!/bin/bash
# sanity after container restore: gpu tests + smoke + bench lines for target / c4 / c1 / c3 / c5 and the reference arm
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench_target.json 2> gpurun_out/bench_target.err; echo "target rc=$?"; cat gpurun_out/bench_target.json
for w in c4 c1 c3 c5; do
  timeout 600 python bench.py --workload $w --steps 30 --warmup 3 > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; echo "$w rc=$?"; cat gpurun_out/bench_$w.json
done
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; cat gpurun_out/bench_ref.json
nvidia-smi -L
