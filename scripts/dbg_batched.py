"""Debug helper: prints the batched path's telemetry for a few shapes and compares with the oracle."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from helpers import ob, ora

ctx = ob.default_context(0)
import sys as _s
cg = int(_s.argv[1]) if len(_s.argv) > 1 else 0
ctx.set_tuning(batch_mode=1, batch_cta_group=cg)
print("cta_group", cg)
for (n, dim, nq, k, metric) in [(300, 24, 2, 30, ob.Metric.DotProduct), (300, 32, 2, 30, ob.Metric.DotProduct), (2000, 64, 8, 30, ob.Metric.DotProduct),
                                (2000, 64, 8, 30, ob.Metric.Cosine), (2000, 64, 8, 30, ob.Metric.Euclidean), (5000, 768, 64, 100, ob.Metric.DotProduct)]:
    v = ora.synth_fill(0, n, dim, 0x7735 + n)
    q = ora.synth_fill(0, nq, dim, 0xBEEF + nq)
    s = ob.VecStore(dim); s.add_vectors(v)
    tt = ob.TakeType.Min if metric == ob.Metric.Euclidean else ob.TakeType.Max
    got = s.query(q, metric).take(k).collect_arrays()
    w = ctx.last_work()
    want = ora.vecstore_query(v, q, metric, tt, k, None, None, ora.CANONICAL)
    same = np.array_equal(got[0], want[0]) and np.array_equal(got[1].view(np.uint32), want[1].view(np.uint32))
    print(f"n={n} dim={dim} nq={nq} k={k} {metric.name}: used={w['batch_used']} fallback={w['batch_fallback']} max_err={w['batch_max_err']:.3e} "
          f"delta={w['batch_delta']:.3e} scan_ms={w['scan_ms']:.3f} rows_scored={w['rows_scored']} launches={w['kernel_launches']} same={same}", flush=True)
