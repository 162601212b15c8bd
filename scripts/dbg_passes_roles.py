"""Which role bounds the single-pass K2 kernel on the c2 shape (OTTERS_BATCH_DBG bits: 1 no loads, 4 no epilogue, 8 no MMAs)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import otters_b200 as ob
from bench import synth_fill_np
rows, dim, nq = int(os.environ.get("ROWS", 1_000_000)), 768, 1024
ctx = ob.default_context(0)
s = ob.VecStore(dim, ctx); s.add_synthetic(0, rows, 0x7735)
q = synth_fill_np(0, nq, dim, 0xBEEF)
for cg in (1, 2):
    for dbg in (0, 4, 5, 12, 13):
        os.environ["OTTERS_BATCH_DBG"] = str(dbg)
        ctx.set_tuning(batch_mode=1, batch_cta_group=cg, batch_passes=1, timing=1)
        ts = []
        for _ in range(3):
            try:
                s.query(q, ob.Metric.DotProduct).take(100).collect_arrays()
            except Exception as e:
                print("err", e)
            ts.append(ctx.last_work()["scan_ms"])
        print(f"single-pass cg={cg} dbg={dbg:2d} kernel_ms={min(ts):.3f}", flush=True)
