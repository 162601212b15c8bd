#!/bin/bash
# round-2 record run on one GPU: the whole GPU suite, every workload's bench line (with CPU baseline and parity check), the
# reference arm (also under torchrun's OMP_NUM_THREADS=1), smoke
mkdir -p gpurun_out/r2_bench
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_final.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest_final.log
for w in target c4 c3 c5 c1 c2; do
  timeout 900 python bench.py --workload $w --steps 50 --warmup 5 > gpurun_out/r2_bench/bench_$w.json 2> gpurun_out/r2_bench/bench_$w.err; echo "$w rc=$?"
  python -c "
import json; d=json.load(open('gpurun_out/r2_bench/bench_$w.json')); r=d['roofline']; c=d['cpu_baseline']; p=d['parity_check']
print('  value=%.1f step=%.4f ms e2e=%.1f blocking=%.1f roof=%.1f %s frac=%.3f cpu=%.3f (%d cores) parity=%s %s clocks=%s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['blocking_value'], r['achieved'], r['unit'], r['frac'], c['value'], c['cores'], p['ok'], p['failures'], d['clocks']['reasons']))"
done
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2_bench/bench_ref.json 2> gpurun_out/r2_bench/bench_ref.err; echo "ref rc=$?"
OMP_NUM_THREADS=1 RANK=0 WORLD_SIZE=2 LOCAL_RANK=0 timeout 900 python bench.py --impl reference --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_bench/bench_ref_torchrun_env.json 2> /dev/null; echo "ref (torchrun env) rc=$?"
python -c "
import json
for f in ('bench_ref','bench_ref_torchrun_env'):
    d=json.load(open('gpurun_out/r2_bench/%s.json'%f)); print(f, 'value=%.2f q/s cores=%d ms_per_step=%.1f' % (d['value'], d['cpu_baseline']['cores'], d['ms_per_step']))"
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -1
