#!/bin/bash
mkdir -p gpurun_out
for wl in target c4 c1; do
  timeout 600 python bench.py --workload $wl --steps 20 --warmup 3 > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err; echo "$wl rc=$?"
  tail -c 3000 gpurun_out/bench_$wl.json; tail -5 gpurun_out/bench_$wl.err
done
