"""Timing experiments on the c2 shape: which pipeline role bounds the batched kernel (OTTERS_BATCH_DBG bits)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import otters_b200 as ob
from bench import synth_fill_np
rows, dim, nq = int(os.environ.get("ROWS", 1_000_000)), 768, 1024
ctx = ob.default_context(0)
s = ob.VecStore(dim, ctx); s.add_synthetic(0, rows, 0x7735)
q = synth_fill_np(0, nq, dim, 0xBEEF)
for cg in (1, 2):
    for dbg in (0, 4):
        os.environ["OTTERS_BATCH_DBG"] = str(dbg)
        ctx.set_tuning(batch_mode=1, batch_cta_group=cg, timing=1)
        ts = []
        for _ in range(3):
            try:
                s.query(q, ob.Metric.DotProduct).take(100).collect_arrays()
            except Exception as e:
                print("err", e)
            ts.append(ctx.last_work()["scan_ms"])
        w = ctx.last_work()
        print(f"cg={cg} dbg={dbg:2d} kernel_ms={min(ts):.3f} used={w['batch_used']} fallback={w['batch_fallback']} max_err={w['batch_max_err']:.3e} delta={w['batch_delta']:.3e}", flush=True)
