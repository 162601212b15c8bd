"""Single-pass K2 variants on BASELINE config 2 (1M x 768, 1024 queries, top-100): k-blocks per stage (OTTERS_K2_KPS) and, for CTA
pairs, loads that credit the leader's barrier directly (OTTERS_K2_DIRECT).  One process per configuration (the switches are read
once); every run is compared with the 3xTF32 result bytes.  usage: python scripts/dbg_k2_variants.py <cta_group>"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import otters_b200 as ob
from bench_workloads import synth_fill_np
cg = int(sys.argv[1]) if len(sys.argv) > 1 else 1
rows, dim, nq = int(os.environ.get("ROWS", 1_000_000)), int(os.environ.get("DIM", 768)), 1024
ctx = ob.default_context(0)
s = ob.VecStore(dim, ctx); s.add_synthetic(0, rows, 0x7735)
q = synth_fill_np(0, nq, dim, 0xBEEF)
ctx.set_tuning(batch_mode=1, batch_passes=3)
ref = s.query(q, ob.Metric.DotProduct).take(100).collect_arrays()
ctx.set_tuning(batch_mode=1, batch_cta_group=cg, batch_passes=1, timing=1)
ts = []
for _ in range(4):
    got = s.query(q, ob.Metric.DotProduct).take(100).collect_arrays()
    ts.append(ctx.last_work()["scan_ms"])
w = ctx.last_work()
same = all(np.array_equal(a, b) for a, b in zip(got, ref))
print(f"single-pass cg={cg} KPS={os.environ.get('OTTERS_K2_KPS', '1')} DIRECT={os.environ.get('OTTERS_K2_DIRECT', '0')} kernel_ms={min(ts):.3f} "
      f"used={w['batch_used']} passes={w['batch_passes']} fallback={w['batch_fallback']} same_as_3x={same}", flush=True)
