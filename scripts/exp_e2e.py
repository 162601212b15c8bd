"""Why is the scan kernel ~10 % slower when every query is followed by a host sync?  (experiment)"""
import ctypes as C, sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import otters_b200 as ob
from otters_b200 import _ffi
import bench

rows, dim, k = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000, 768, 100
torch.cuda.set_device(0)
ctx = ob.Context(0, torch.cuda.current_stream().cuda_stream)
store = ob.VecStore(dim, ctx)
store.add_synthetic(0, rows, bench.DATA_SEED)
queries = bench.synth_fill_np(0, 16, dim, bench.QUERY_SEED)
idx, sc = np.zeros(k, np.uint64), np.zeros(k, np.float32)
rec = torch.empty((k, 16), dtype=torch.uint8, device="cuda")

def vq_of(i):
    vq = _ffi.VecQuery(); q = queries[i % 16]
    vq.queries = q.ctypes.data_as(_ffi.c_f32p); vq.nq, vq.dim, vq.metric, vq.take_type, vq.k = 1, dim, 1, 0, k
    return vq

def sync_query(i):
    n = C.c_uint64(); vq = vq_of(i)
    assert _ffi.otters_vecstore_query(store._handle(), C.byref(vq), idx.ctypes.data_as(_ffi.c_u64p), sc.ctypes.data_as(_ffi.c_f32p), None, k, C.byref(n)) == 0

def async_query(i):
    vq = vq_of(i)
    assert _ffi.otters_query_local_device(store._handle(), None, C.byref(vq), None, 0, C.c_void_p(rec.data_ptr()), None) == 0

for i in range(3): sync_query(i)
for name, fn, post in [("sync", sync_query, None), ("sync+sleep1ms", sync_query, lambda: time.sleep(0.001)),
                       ("async", async_query, None), ("async+sync_each", async_query, torch.cuda.synchronize)]:
    torch.cuda.synchronize(); t0 = time.perf_counter(); sm = []
    for i in range(20):
        fn(i)
        if post: post()
        if name.startswith("sync"): sm.append(ctx.last_work()["scan_ms"])
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 20 * 1e3
    w = ctx.last_work()
    print(f"{name:18s} {dt:8.3f} ms/step  scan_ms(last)={w['scan_ms']:.3f} select_ms={w['select_ms']:.3f} mean_scan={np.mean(sm) if sm else float('nan'):.3f}")
