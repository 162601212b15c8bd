"""Role-by-role timing of K2's bf16 rung on the c2 shape, for a library built with -DOTTERS_K2_EXPERIMENTS
(OTTERS_BATCH_DBG bits: 1 no TMA loads, 4 no epilogue, 8 no MMAs; results are garbage and never accepted, so every call falls
back to the streaming kernel afterwards).  Run under `ncu --metrics gpu__time_duration.sum -k regex:batch_kernel`: the launch
list carries the kernel times, one batch per setting, in the order printed here."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import otters_b200 as ob
from bench_workloads import synth_fill_np
rows, dim, nq = int(os.environ.get("ROWS", 1_000_000)), 768, int(os.environ.get("NQ", 1024))
ctx = ob.default_context(0)
s = ob.VecStore(dim, ctx); s.add_synthetic(0, rows, 0x7735)
q = synth_fill_np(0, nq, dim, 0xBEEF)
ctx.set_tuning(batch_mode=1, batch_passes=2)
order = []
for dbg in (0, 0, 4, 5, 12, 13, 1, 8):
    os.environ["OTTERS_BATCH_DBG"] = str(dbg)
    try:
        s.query(q, ob.Metric.DotProduct).take(100).collect_arrays()
    except Exception as e:
        print("err", e)
    order.append(dbg)
    print(f"dbg={dbg:2d} done", flush=True)
print("ORDER", order)
