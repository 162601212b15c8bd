"""K2 on the c2 shape: kernel time and certificate of the single-pass tf32 selection vs the 3xTF32 split, single CTAs and pairs."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import otters_b200 as ob
from bench import synth_fill_np
rows, dim, nq = int(os.environ.get("ROWS", 1_000_000)), 768, 1024
ctx = ob.default_context(0)
s = ob.VecStore(dim, ctx); s.add_synthetic(0, rows, 0x7735)
q = synth_fill_np(0, nq, dim, 0xBEEF)
ref = None
for metric in (ob.Metric.DotProduct, ob.Metric.Cosine, ob.Metric.Euclidean):
    for cg in (1, 2):
        for passes in ((3, 0) if os.environ.get("QUICK") else (3, 1, 0)):
            ctx.set_tuning(batch_mode=1, batch_cta_group=cg, batch_passes=passes, timing=1)
            ts = []
            for _ in range(3):
                got = s.query(q, metric).take(100).collect_arrays()
                ts.append(ctx.last_work()["scan_ms"])
            w = ctx.last_work()
            key = metric
            if ref is None or ref[0] != key:
                ref = (key, got)
            same = all(np.array_equal(a, b) for a, b in zip(got, ref[1]))
            print(f"{metric.name:10s} cg={cg} passes={passes} kernel_ms={min(ts):.3f} used={w['batch_used']} accepted_passes={w['batch_passes']} "
                  f"attempts={w['batch_attempts']} fallback={w['batch_fallback']} max_err={w['batch_max_err']:.3e} delta={w['batch_delta']:.3e} "
                  f"select_ms={w['select_ms']:.3f} same_as_3x={same}", flush=True)
