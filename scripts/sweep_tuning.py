"""Sweeps the scan-kernel tuning knobs on one resident store (results never change; only time does)."""
import ctypes as C, sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import otters_b200 as ob
from otters_b200 import _ffi
from otters_b200.meta import FilterPack
import bench

wl_name = sys.argv[1] if len(sys.argv) > 1 else "c4"
wl = dict(bench.WORKLOADS[wl_name])
if len(sys.argv) > 2: wl["rows"] = int(sys.argv[2])
rows, dim, chunk, k = wl["rows"], wl["dim"], wl["chunk"], wl["k"]
metric = getattr(ob.Metric, wl["metric"]); tt = ob.TakeType.Min if metric == ob.Metric.Euclidean else ob.TakeType.Max
dev = torch.device("cuda", 0); st = torch.cuda.Stream(device=dev); torch.cuda.set_stream(st)
ctx = ob.Context(0, st.cuda_stream)
fp = None
if wl["meta"]:
    cols = bench.meta_columns(ob, 0, rows, chunk)
    store = ob.MetaStore.from_columns(cols).with_synthetic_vectors(rows, dim, bench.DATA_SEED, 0).with_chunk_size(chunk).with_context(ctx).build()
    expr, _ = bench.meta_expr(ob, rows); fp = FilterPack(expr.compile(store.schema()), store.column_index())
else:
    store = ob.VecStore(dim, ctx); store.add_synthetic(0, rows, bench.DATA_SEED)
queries = bench.synth_fill_np(0, 16, dim, bench.QUERY_SEED)
idx, sc = np.zeros(k, np.uint64), np.zeros(k, np.float32); qs = _ffi.QueryStats()

def query(i):
    vq = _ffi.VecQuery(); q = queries[i % 16]
    vq.queries = q.ctypes.data_as(_ffi.c_f32p); vq.nq, vq.dim, vq.metric, vq.take_type, vq.k = 1, dim, int(metric), int(tt), k
    n = C.c_uint64()
    if wl["meta"]:
        rc = _ffi.otters_metastore_query(store.handle, C.byref(vq), fp.byref(), idx.ctypes.data_as(_ffi.c_u64p), sc.ctypes.data_as(_ffi.c_f32p), None, k, C.byref(n), C.byref(qs))
    else:
        rc = _ffi.otters_vecstore_query(store._handle(), C.byref(vq), idx.ctypes.data_as(_ffi.c_u64p), sc.ctypes.data_as(_ffi.c_f32p), None, k, C.byref(n))
    assert rc == 0, _ffi.last_error()
    return idx[:n.value].copy()

ref = None
grid = [(0,0,0,0,0)]
for kc in (768, 384, 256, 192, 128, 1536, 512):
    if kc > ((dim + 7)//8*8) and kc != 768: continue
    for w in (2, 4, 6, 8, 12, 16):
        for s in (1, 2, 3, 4):
            grid.append((w, s, kc, 0, 0))
seen = set()
for t in grid:
    try:
        ctx.set_tuning(*t)
        r = query(0)
    except ob.OttersError as e:
        continue
    if ref is None: ref = r
    assert np.array_equal(r, ref), f"tuning {t} changed the result"
    ms = []
    for i in range(6):
        query(i + 1); ms.append(ctx.last_work()["scan_ms"])
    w = ctx.last_work()
    gbs = w["scan_bytes"] / (np.median(ms) * 1e-3) / 1e9
    print(f"{wl_name} tuning={t} scan_ms={np.median(ms):.3f} min={min(ms):.3f} GB/s={gbs:.0f} select_ms={w['select_ms']:.3f}", flush=True)
