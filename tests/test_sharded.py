"""Row-sharded search: partitioning, the fixed-size record format and the gather+merge plumbing.

CPU part (gloo, world_size 2): every rank searches its shard with the oracle, the ranks all-gather their k
records and merge — must equal the single-store answer.  GPU part: the CUDA shard path at world size 1, and
(when >= 2 GPUs are visible) the NCCL path under torchrun."""
import os
import subprocess
import sys

import numpy as np
import pytest

from helpers import ROOT, assert_same_results, ob, ora
from otters_b200.sharded import (EMPTY_ROW, RECORD_DTYPE, ShardedSearcher, cyclic_global_rows, cyclic_local_rows,
                                 merge_records_host, shard_range)


def test_shard_range_is_chunk_aligned_partition():
    for n, cs, world in [(10_000_000, 1024, 8), (1000, 96, 3), (5, 1024, 4), (0, 16, 2), (1025, 1024, 2), (777, 1, 5)]:
        prev = 0
        for r in range(world):
            r0, r1 = shard_range(n, cs, world, r)
            assert r0 == prev and r0 <= r1 <= n
            assert r0 % cs == 0 or r0 == n
            prev = r1
        assert prev == n


def test_cyclic_sharding_partitions_rows():
    for n, b, w in [(10_000_000, 1024, 8), (1000, 96, 3), (5, 1024, 4), (0, 16, 2), (1025, 1024, 2), (777, 1, 5), (2048, 1024, 2)]:
        parts = [cyclic_global_rows(n, b, w, r) for r in range(w)]
        assert [len(p) for p in parts] == [cyclic_local_rows(n, b, w, r) for r in range(w)]
        allrows = np.concatenate(parts) if n else np.zeros(0, np.int64)
        assert len(allrows) == n and len(np.unique(allrows)) == n
        assert all(np.all(np.diff(p) > 0) for p in parts)  # local order == global order within a shard
        if n >= b * w * 4:
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= b


def test_merge_records_host_order():
    rec = np.zeros(6, RECORD_DTYPE)
    rec["row"] = [5, 3, 3, EMPTY_ROW, 9, 1]
    rec["score"] = [1.0, 2.0, 2.0, 0.0, 2.0, -0.0]
    rec["qid"] = [0, 1, 0, 0, 0, 0]
    row, score, qid = merge_records_host(rec, 4, True)
    assert list(row) == [3, 3, 9, 5] and list(qid) == [0, 1, 0, 0]
    row, score, qid = merge_records_host(rec, 2, False)
    assert list(row) == [1, 5]


def _gloo_worker(rank, world, port, n, dim, k, out_dir):
    import torch
    import torch.distributed as dist

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    vectors = ora.synth_fill(0, n, dim, 99)
    queries = ora.synth_fill(0, 2, dim, 100)
    r0, r1 = shard_range(n, 64, world, rank)

    def local(kk):
        idx, score, qid = ora.vecstore_query(vectors[r0:r1], queries, ob.Metric.Euclidean, ob.TakeType.Min, kk)
        rec = np.zeros(kk, RECORD_DTYPE)
        rec["row"] = EMPTY_ROW
        rec["row"][: len(idx)] = idx + r0
        rec["score"][: len(idx)] = score
        rec["qid"][: len(idx)] = qid
        return rec

    def gather(rec):
        t = torch.from_numpy(rec.view(np.uint8).reshape(-1, 16).copy())
        out = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        return torch.cat(out).numpy().reshape(-1).view(RECORD_DTYPE)

    row, score, qid = ShardedSearcher(world, rank, gather, lambda g, kk: merge_records_host(g, kk, False)).search(local, k)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), row=row, score=score, qid=qid)
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_plumbing_gloo_world2(tmp_path):
    import torch.multiprocessing as mp

    n, dim, k, world = 1000, 24, 37, 2
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_gloo_worker, args=(world, port, n, dim, k, str(tmp_path)), nprocs=world, join=True)
    vectors = ora.synth_fill(0, n, dim, 99)
    queries = ora.synth_fill(0, 2, dim, 100)
    want = ora.vecstore_query(vectors, queries, ob.Metric.Euclidean, ob.TakeType.Min, k)
    for r in range(world):
        z = np.load(tmp_path / f"rank{r}.npz")
        assert_same_results((z["row"], z["score"], z["qid"]), want, f"rank {r}")


@pytest.mark.gpu
def test_cuda_shard_world1_matches_plain_query(ctx):
    import ctypes as C

    from otters_b200 import _ffi
    from otters_b200.sharded import CudaShard

    n, dim, k = 5000, 48, 25
    v = ora.synth_fill(0, n, dim, 5)
    q = ora.synth_fill(0, 1, dim, 6)
    store = ob.VecStore(dim)
    store.add_vectors(v)
    shard = CudaShard(store, row_base=1000, k_max=k)
    vq = _ffi.VecQuery()
    vq.queries = q.ctypes.data_as(_ffi.c_f32p)
    vq.nq, vq.dim, vq.metric, vq.take_type, vq.k = 1, dim, int(ob.Metric.Cosine), 1, k
    got = shard.search(vq, None, k, True)
    want = ora.vecstore_query(v, q, ob.Metric.Cosine, ob.TakeType.Max, k)
    assert_same_results((got[0] - 1000, got[1], got[2]), want)


@pytest.mark.gpu
def test_nccl_sharded_search_world2():
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29711", os.path.join(ROOT, "tests", "dist_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "DIST_CHECK_OK" in r.stdout


@pytest.mark.gpu
def test_synthetic_sharded_generator_matches_global_rows(ctx):
    n, dim, b, w = 5000, 24, 64, 3
    full = ora.synth_fill(0, n, dim, 77)
    q = ora.synth_fill(0, 1, dim, 78)
    for r in range(w):
        g = cyclic_global_rows(n, b, w, r)
        s = ob.VecStore(dim)
        s.add_synthetic_sharded(w, r, b, len(g), 77)
        assert np.array_equal(s.inv_norms().view(np.uint32), ora.inv_norms(full[g]).view(np.uint32))
        got = s.query(q[0], ob.Metric.DotProduct).take(20).collect_arrays()
        want = ora.vecstore_query(full[g], q, ob.Metric.DotProduct, ob.TakeType.Max, 20)
        assert_same_results(got, want, f"rank {r}")
