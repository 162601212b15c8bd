#!/bin/bash
# round-1 record run on one GPU: every workload's bench line (with CPU baseline), reference arm, launch list, planner capture
mkdir -p gpurun_out
for w in target c4 c3 c5 c1 c2; do
  timeout 600 python bench.py --workload $w --steps 50 --warmup 5 > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; echo "$w rc=$?"
  python -c "
import json; d=json.load(open('gpurun_out/bench_$w.json')); r=d['roofline']; c=d['cpu_baseline']
print('  value=%.1f step=%.4f ms e2e=%.1f roof=%.1f %s frac=%.3f cpu=%.3f (%d cores) %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], r['achieved'], r['unit'], r['frac'], c['value'], c['cores'], d['clocks']['reasons']))"
done
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_target.csv python bench.py --steps 3 --warmup 2 --no-cpu > gpurun_out/ncu_launch.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan_planner_kernel -s 2 -c 1 -o gpurun_out/scan_planner_c4 -f python bench.py --workload c4 --steps 2 --warmup 2 --no-cpu > gpurun_out/ncu_pl.log 2>&1; echo "planner ncu rc=$?"
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -1
