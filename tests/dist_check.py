"""Run under torchrun with >= 2 GPUs: sharded VecStore and MetaStore search over NCCL vs the oracle."""
import ctypes as C
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import otters_b200 as ob  # noqa: E402
from helpers import assert_same_results  # noqa: E402
from oracle import oracle as ora  # noqa: E402
from otters_b200 import _ffi  # noqa: E402
from otters_b200.meta import FilterPack  # noqa: E402
from otters_b200.sharded import CudaShard, cyclic_global_rows, shard_range  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
ctx = ob.Context(local, stream.cuda_stream)

n, dim, cs, k = 20000, 96, 128, 50
vectors = ora.synth_fill(0, n, dim, 7)
queries = ora.synth_fill(0, 2, dim, 8)
r0, r1 = shard_range(n, cs, world, rank)
rng = np.random.default_rng(3)
val = (np.arange(n) // cs % 5 * 10 + rng.integers(0, 10, n)).astype(np.int32)
nulls = rng.random(n) < 0.02
col = ob.Column.from_numpy("val", ob.DataType.Int32, val, nulls)
col_shard = ob.Column.from_numpy("val", ob.DataType.Int32, val[r0:r1], nulls[r0:r1])
expr = ob.col("val").gte(20) & ob.col("val").neq(33)

for metric, tt in ((ob.Metric.Cosine, ob.TakeType.Max), (ob.Metric.Euclidean, ob.TakeType.Min)):
    for nq in (1, 2):
        vq = _ffi.VecQuery()
        q = np.ascontiguousarray(queries[:nq])
        vq.queries = q.ctypes.data_as(_ffi.c_f32p)
        vq.nq, vq.dim, vq.metric, vq.take_type, vq.k = nq, dim, int(metric), int(tt), k
        # VecStore shard
        vs = ob.VecStore(dim, ctx)
        vs.add_vectors(vectors[r0:r1])
        got = CudaShard(vs, r0, k).search(vq, None, k, tt == ob.TakeType.Max)
        want = ora.vecstore_query(vectors, q, metric, tt, k)
        assert_same_results(got, want, f"vec {metric.name} nq={nq} rank {rank}")
        # MetaStore shard
        ms = ob.MetaStore.from_columns([col_shard]).with_vectors(vectors[r0:r1]).with_chunk_size(cs).with_context(ctx).build()
        fp = FilterPack(expr.compile(ms.schema()), ms.column_index())
        shard = CudaShard(ms, r0, k)
        gathered, st = shard.enqueue(vq, fp, k, want_stats=True)
        got = shard.merge(gathered, k, tt == ob.TakeType.Max)
        ost = ora.MetaStore(vectors, [col], cs)
        ofp = ora.FilterPack.from_compiled(expr.compile(ms.schema()), ms.column_index())
        oi, os_, oq, ostats = ost.query(q, metric, tt, k, None, ofp)
        assert_same_results(got, (oi, os_, oq), f"meta {metric.name} nq={nq} rank {rank}")
        tot = torch.tensor([st.evaluated_chunks, st.vectors_compared, st.total_chunks], dtype=torch.int64, device="cuda")
        dist.all_reduce(tot)
        assert tot.tolist() == [ostats["evaluated_chunks"], ostats["vectors_compared"], ostats["total_chunks"]], (tot.tolist(), ostats)
        # block-cyclic sharding: blocks of cs rows dealt round-robin; global ids come back through otters_shard_map
        g = cyclic_global_rows(n, cs, world, rank)
        colc = ob.Column.from_numpy("val", ob.DataType.Int32, val[g], nulls[g])
        msc = ob.MetaStore.from_columns([colc]).with_vectors(vectors[g]).with_chunk_size(cs).with_context(ctx).build()
        shardc = CudaShard(msc, 0, k, block_rows=cs)
        gathered, st = shardc.enqueue(vq, fp, k, want_stats=True)
        got = shardc.merge(gathered, k, tt == ob.TakeType.Max)
        assert_same_results(got, (oi, os_, oq), f"cyclic meta {metric.name} nq={nq} rank {rank}")
        tot = torch.tensor([st.evaluated_chunks, st.vectors_compared, st.total_chunks], dtype=torch.int64, device="cuda")
        dist.all_reduce(tot)
        assert tot.tolist() == [ostats["evaluated_chunks"], ostats["vectors_compared"], ostats["total_chunks"]], (tot.tolist(), ostats)
        vsc = ob.VecStore(dim, ctx)
        vsc.add_vectors(vectors[g])
        shv = CudaShard(vsc, 0, k, block_rows=cs)
        got = shv.search(vq, None, k, tt == ob.TakeType.Max)
        assert_same_results(got, want, f"cyclic vec {metric.name} nq={nq} rank {rank}")
        # fused exchange over peer memory (select + NVLink stores + flags + merge in one kernel), when the box maps it
        if shv.enable_peer_exchange():
            fused_ok = True
            for rep in range(5):  # wraps around the OTTERS_EXCHANGE_SLOTS-deep areas
                got, _ = shv.search_fused(vq, None, k)
                got = tuple(a.copy() for a in got)  # views of buffers the next call overwrites
                assert_same_results(got, want, f"fused vec {metric.name} nq={nq} rank {rank} rep {rep}")
            assert shardc.enable_peer_exchange()
            got, st = shardc.search_fused(vq, fp, k, want_stats=True)
            assert_same_results(got, (oi, os_, oq), f"fused meta {metric.name} nq={nq} rank {rank}")
            tot = torch.tensor([st.evaluated_chunks, st.vectors_compared, st.total_chunks], dtype=torch.int64, device="cuda")
            dist.all_reduce(tot)
            assert tot.tolist() == [ostats["evaluated_chunks"], ostats["vectors_compared"], ostats["total_chunks"]], (tot.tolist(), ostats)
            # two queries in flight per rank (otters_query_submit / otters_query_wait), exchange attached: the areas are
            # OTTERS_EXCHANGE_SLOTS deep, ranks drift apart by up to two queries
            if nq == 1:
                qs8 = ora.synth_fill(0, 9, dim, 9)
                vqs = []
                for i in range(9):
                    v8 = _ffi.VecQuery()
                    v8.queries = qs8[i:i + 1].ctypes.data_as(_ffi.c_f32p)
                    v8.nq, v8.dim, v8.metric, v8.take_type, v8.k = 1, dim, int(metric), int(tt), k
                    vqs.append(v8)
                for sh, flt, is_meta in ((shv, None, False), (shardc, fp, True)):
                    tickets = [sh.submit(vqs[0], flt)]
                    for i in range(9):
                        if i + 1 < 9:
                            tickets.append(sh.submit(vqs[i + 1], flt))
                        got, st = sh.wait(tickets[i], k, want_stats=is_meta)
                        got = tuple(a.copy() for a in got)
                        if is_meta:
                            oi8, os8, oq8, ostats8 = ost.query(qs8[i:i + 1], metric, tt, k, None, ofp)
                            assert_same_results(got, (oi8, os8, oq8), f"pipelined meta {metric.name} query {i} rank {rank}")
                            tot = torch.tensor([st.evaluated_chunks, st.vectors_compared], dtype=torch.int64, device="cuda")
                            dist.all_reduce(tot)
                            assert tot.tolist() == [ostats8["evaluated_chunks"], ostats8["vectors_compared"]], (tot.tolist(), ostats8)
                        else:
                            assert_same_results(got, ora.vecstore_query(vectors, qs8[i:i + 1], metric, tt, k),
                                                f"pipelined vec {metric.name} query {i} rank {rank}")
                pipelined_ok = True
            # a rank whose shard is empty still takes part
            vse = ob.VecStore(dim, ctx)
            if rank == 0:
                vse.add_vectors(vectors[:500])
            she = CudaShard(vse, 0, k)
            assert she.enable_peer_exchange()
            got, _ = she.search_fused(vq, None, k)
            assert_same_results(got, ora.vecstore_query(vectors[:500], q, metric, tt, k), f"fused empty shard {metric.name} nq={nq}")
        elif rank == 0:
            print("peer exchange unavailable:", shv.peer_error)
dist.barrier()
if rank == 0:
    print("DIST_CHECK_OK", "fused_exchange=%s" % ("fused_ok" in globals()), "pipelined=%s" % ("pipelined_ok" in globals()))
dist.destroy_process_group()
