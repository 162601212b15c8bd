import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from helpers import ob, ora
os.environ["OTTERS_BATCH_TRACE"] = "1"
ctx = ob.default_context(0)
n, dim, nq = 4097, 100, 33
v = ora.synth_fill(0, n, dim, 0x7735 + n); q = ora.synth_fill(0, nq, dim, 0xBEEF + nq)
s = ob.VecStore(dim); s.add_vectors(v)
for cg in (2, 1):
    ctx.set_tuning(batch_mode=1, batch_cta_group=cg)
    for rep in range(6):
        for k in (1, 30, 700, 1024):
            for call in ("take_max", "take_min"):
                print(f"cg={cg} rep={rep} k={k} {call}", file=sys.stderr, flush=True)
                getattr(s.query(q, ob.Metric.DotProduct), call)(k).collect_arrays()
