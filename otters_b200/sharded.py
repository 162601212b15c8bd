"""Row-sharded multi-GPU search (one process per GPU, torch.distributed for the plumbing).

Rows are split either into contiguous chunk-aligned ranges or block-cyclically (blocks of chunk_size rows
dealt round-robin, which keeps shards balanced under range filters); every rank holds its shard in HBM,
answers the query locally (``otters_query_local_device`` leaves k fixed-size records on the device), the
ranks all-gather the records (NCCL over NVLink; k*16 bytes per rank) and every rank runs the same final
merge kernel (``otters_topk_merge_device``).  Global row id = shard base + local row.
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, Optional, Tuple

import numpy as np

RECORD_DTYPE = np.dtype([("row", "<u8"), ("score", "<f4"), ("qid", "<u4")])  # otters_topk_record
EMPTY_ROW = np.uint64(0xFFFFFFFFFFFFFFFF)


def shard_range(n_rows: int, chunk_size: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous row range of `rank`, aligned to chunk boundaries so zonemaps and stats shard with the rows."""
    chunk_size = max(int(chunk_size), 1)
    n_chunks = (n_rows + chunk_size - 1) // chunk_size
    per = (n_chunks + world - 1) // world
    r0 = min(rank * per * chunk_size, n_rows)
    r1 = min((rank + 1) * per * chunk_size, n_rows)
    return r0, r1


def cyclic_local_rows(n_rows: int, block_rows: int, world: int, rank: int) -> int:
    """Rows held by `rank` when blocks of `block_rows` rows are dealt round-robin (block b -> rank b % world)."""
    block_rows = max(int(block_rows), 1)
    n_blocks = (n_rows + block_rows - 1) // block_rows
    mine = (n_blocks - rank + world - 1) // world if n_blocks > rank else 0
    if mine == 0:
        return 0
    last_block = (mine - 1) * world + rank
    tail = n_rows - last_block * block_rows  # rows in my last block (it may be the short global tail)
    return (mine - 1) * block_rows + min(block_rows, tail)


def cyclic_global_rows(n_rows: int, block_rows: int, world: int, rank: int) -> np.ndarray:
    """Global row ids of the local rows 0..n_local-1 of `rank` (same formula as otters_shard_map)."""
    n_local = cyclic_local_rows(n_rows, block_rows, world, rank)
    local = np.arange(n_local, dtype=np.int64)
    return (local // block_rows * world + rank) * block_rows + local % block_rows


def merge_records_host(records: np.ndarray, k: int, take_max: bool):
    """Reference merge of gathered records on the host (used by the CPU/gloo tests of the plumbing only;
    the product merges on the device).  Order: better score, lower row, lower query id."""
    rec = records[records["row"] != EMPTY_ROW]
    score = rec["score"].astype(np.float32) + np.float32(0.0)
    key = -score if take_max else score
    order = np.lexsort((rec["qid"], rec["row"], key))[:k]
    return rec["row"][order], rec["score"][order], rec["qid"][order]


class ShardedSearcher:
    """Local search -> all-gather -> merge.  `local_fn(k) -> records` and `merge_fn(gathered, k)` are
    injected so the same plumbing runs on CUDA/NCCL (product) and on CPU/gloo (tests)."""

    def __init__(self, world: int, rank: int, gather_fn: Callable, merge_fn: Callable):
        self.world, self.rank = world, rank
        self._gather = gather_fn
        self._merge = merge_fn

    def search(self, local_fn: Callable, k: int):
        local = local_fn(k)
        gathered = self._gather(local)
        return self._merge(gathered, k)


class CudaShard:
    """Product implementation: a VecStore or MetaStore shard on this rank's GPU."""

    def __init__(self, store, row_base: int, k_max: int, group=None, block_rows: int = 0):
        import torch
        import torch.distributed as dist

        from . import _ffi
        from .meta import MetaStore

        self._torch, self._dist, self._ffi = torch, dist, _ffi
        self.store = store
        self.is_meta = isinstance(store, MetaStore)
        if self.is_meta and store.row_order() is not None:
            from .types import OttersError

            # the shard map turns store positions into global row ids on the device; a row order would need its own id map there
            raise OttersError("row-sharded search does not support stores built with_row_order")
        self.ctx = store.ctx
        self.row_base = int(row_base)
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        # block_rows > 0: the shard holds every world-th block of block_rows rows (block-cyclic); else contiguous
        self.map = _ffi.ShardMap(self.row_base, self.world if block_rows else 1, self.rank if block_rows else 0, int(block_rows))
        dev = torch.device("cuda", self.ctx.device)
        self.local = torch.empty((k_max, 16), dtype=torch.uint8, device=dev)
        self.gathered = torch.empty((self.world * k_max, 16), dtype=torch.uint8, device=dev)
        self.k_max = k_max
        self.peer = None      # _ffi.PeerExchange once enable_peer_exchange() succeeded
        self.peer_error = None
        self._seq = 0
        # The NCCL path mixes work of two producers: the library enqueues on the context's stream, torch.distributed on
        # torch's current stream.  They are ordered for free only when both are the SAME stream (build the Context from
        # torch.cuda.current_stream().cuda_stream); otherwise enqueue()/merge() fall back to host-side synchronisation.

    # ---- fused exchange over peer memory -------------------------------------------------------------------
    def enable_peer_exchange(self) -> bool:
        """Maps every rank's record/flag areas into every process (torch symmetric memory: CUDA VMM handles
        exchanged over the process group) so that ``otters_query_exchange`` can store records straight into the
        peers' HBM over NVLink.  Returns False (and keeps the NCCL all-gather path) when the mapping is unavailable."""
        if self.world < 2 or self.world > 8:
            return False
        torch, dist, ffi = self._torch, self._dist, self._ffi
        try:
            import torch.distributed._symmetric_memory as symm

            flag_bytes = 256  # EXCHANGE_SLOTS * world uint32, padded
            rec_bytes = ffi.EXCHANGE_SLOTS * self.world * self.k_max * 16
            dev = torch.device("cuda", self.ctx.device)
            buf = symm.empty(flag_bytes + rec_bytes, dtype=torch.uint8, device=dev)
            group = self.group if self.group is not None else dist.group.WORLD
            try:
                hdl = symm.rendezvous(buf, group=group)
            except TypeError:
                hdl = symm.rendezvous(buf, group.group_name)
            buf.zero_()
            torch.cuda.synchronize()
            dist.barrier(group=self.group)
            ptrs = [int(p) for p in hdl.buffer_ptrs]
            if len(ptrs) != self.world or any(p == 0 for p in ptrs):
                raise RuntimeError("symmetric memory returned no peer pointers")
            self._symm = (buf, hdl)  # keep the mapping alive
            self._peer_rec = (C.c_void_p * self.world)(*[p + flag_bytes for p in ptrs])
            self._peer_flg = (C.c_void_p * self.world)(*ptrs)
            self.peer = ffi.PeerExchange(self.world, self.rank, self.k_max, C.cast(self._peer_rec, C.POINTER(C.c_void_p)),
                                         C.cast(self._peer_flg, C.POINTER(C.c_void_p)))
        except Exception as e:  # no symmetric memory on this box / build: NCCL path stays
            self.peer = None
            self.peer_error = f"{type(e).__name__}: {e}"
        # every rank must take the same path
        ok = torch.tensor([1 if self.peer is not None else 0], dtype=torch.int32, device=torch.device("cuda", self.ctx.device))
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)
        if int(ok.item()) == 0:
            self.peer = None
        return self.peer is not None

    def search_fused(self, vq, fp, k: int, fetch: bool = True, want_stats: bool = False):
        """Local search with the exchange and the merge fused into the selection kernel (no NCCL call).
        Returns ((rows, scores, query ids), stats); the arrays are views of buffers that the next call overwrites."""
        ffi = self._ffi
        self._seq += 1
        cache = getattr(self, "_fused_cache", None)
        if cache is None or cache[0] < k:
            # everything that does not change from query to query is bound once: this call sits on the critical path
            out = (np.zeros(k, np.uint64), np.zeros(k, np.float32), np.zeros(k, np.uint32))
            cache = self._fused_cache = (
                k, out, out[0].ctypes.data_as(ffi.c_u64p), out[1].ctypes.data_as(ffi.c_f32p), out[2].ctypes.data_as(ffi.c_u32p),
                C.c_uint64(0), ffi.QueryStats(), C.byref(self.map), C.byref(self.peer))
            if not self.is_meta:
                self.store._flush()
        _, out, p_idx, p_score, p_qid, out_len, st, map_ref, peer_ref = cache
        if self.is_meta:
            vs, ms, flt = None, self.store.handle, (fp.byref() if fp else None)
        else:
            vs, ms, flt = self.store._handle(), None, None
        st_ref = C.byref(st) if want_stats else None
        if fetch:
            rc = ffi.otters_query_exchange(vs, ms, C.byref(vq), flt, map_ref, peer_ref, self._seq, p_idx, p_score, p_qid, k,
                                           C.byref(out_len), st_ref)
        else:
            rc = ffi.otters_query_exchange(vs, ms, C.byref(vq), flt, map_ref, peer_ref, self._seq, None, None, None, 0, None, st_ref)
        if rc != 0:
            from .types import OttersError

            raise OttersError(ffi.last_error())
        if not fetch:
            return None, st
        m = min(out_len.value, k)
        return (out[0][:m], out[1][:m], out[2][:m]), st

    def submit(self, vq, fp) -> int:
        """Non-blocking search (``otters_query_submit``): enqueues the query on one of the context's two lanes — with the fused
        peer exchange when it is enabled and world > 1 — and returns a ticket for ``wait``.  Keep at most two tickets outstanding."""
        ffi = self._ffi
        self._seq += 1
        if self.is_meta:
            vs, ms, flt = None, self.store.handle, (fp.byref() if fp else None)
        else:
            self.store._flush()
            vs, ms, flt = self.store._handle(), None, None
        ticket = C.c_uint64(0)
        rc = ffi.otters_query_submit(vs, ms, C.byref(vq), flt, C.byref(self.map), C.byref(self.peer) if self.peer is not None else None,
                                     self._seq, C.byref(ticket))
        if rc != 0:
            from .types import OttersError

            raise OttersError(ffi.last_error())
        return ticket.value

    def wait(self, ticket: int, k: int, want_stats: bool = False):
        """Blocks until the query behind ``ticket`` has finished; returns ((rows, scores, query ids), stats) like search_fused."""
        ffi = self._ffi
        cache = getattr(self, "_wait_cache", None)
        if cache is None or cache[0] < k:
            out = (np.zeros(k, np.uint64), np.zeros(k, np.float32), np.zeros(k, np.uint32))
            cache = self._wait_cache = (k, out, out[0].ctypes.data_as(ffi.c_u64p), out[1].ctypes.data_as(ffi.c_f32p),
                                        out[2].ctypes.data_as(ffi.c_u32p), C.c_uint64(0), ffi.QueryStats())
        _, out, p_idx, p_score, p_qid, out_len, st = cache
        rc = ffi.otters_query_wait(self.ctx.handle, ticket, p_idx, p_score, p_qid, k, C.byref(out_len), C.byref(st) if want_stats else None)
        if rc != 0:
            from .types import OttersError

            raise OttersError(ffi.last_error())
        m = min(out_len.value, k)
        return (out[0][:m], out[1][:m], out[2][:m]), st

    def enqueue(self, vq, fp, k: int, want_stats: bool = False):
        """Enqueues local search + all-gather on the context's stream; returns the gathered record tensor."""
        ffi = self._ffi
        st = ffi.QueryStats()
        local = self.local[:k]
        if self.is_meta:
            rc = ffi.otters_query_local_device(None, self.store.handle, C.byref(vq), fp.byref() if fp else None, C.byref(self.map),
                                               C.c_void_p(local.data_ptr()), C.byref(st) if want_stats else None)
        else:
            self.store._flush()
            rc = ffi.otters_query_local_device(self.store._handle(), None, C.byref(vq), None, C.byref(self.map),
                                               C.c_void_p(local.data_ptr()), None)
        if rc != 0:
            from .types import OttersError

            raise OttersError(ffi.last_error())
        if self.world > 1:
            gathered = self.gathered[: self.world * k]
            if not self._stream_shared():
                self.ctx.synchronize()  # the records must be complete before NCCL (on torch's stream) reads them
            self._dist.all_gather_into_tensor(gathered, local, group=self.group)
        else:
            gathered = local
        return gathered, st

    def _stream_shared(self) -> bool:
        return self.ctx.stream is not None and self.ctx.stream == self._torch.cuda.current_stream().cuda_stream

    def merge(self, gathered, k: int, take_max: bool, fetch: bool = True):
        ffi = self._ffi
        n = gathered.shape[0]
        if self.world > 1 and not self._stream_shared():
            self._torch.cuda.current_stream().synchronize()  # the all-gather must have landed before the merge kernel reads it
        out_len = C.c_uint64(0)
        if not fetch:
            rc = ffi.otters_topk_merge_device(self.ctx.handle, C.c_void_p(gathered.data_ptr()), n, k, 1 if take_max else 0,
                                              None, None, None, 0, C.byref(out_len))
            if rc != 0:
                from .types import OttersError

                raise OttersError(ffi.last_error())
            return None
        idx = np.zeros(k, np.uint64)
        score = np.zeros(k, np.float32)
        qid = np.zeros(k, np.uint32)
        rc = ffi.otters_topk_merge_device(self.ctx.handle, C.c_void_p(gathered.data_ptr()), n, k, 1 if take_max else 0,
                                          idx.ctypes.data_as(ffi.c_u64p), score.ctypes.data_as(ffi.c_f32p),
                                          qid.ctypes.data_as(ffi.c_u32p), k, C.byref(out_len))
        if rc != 0:
            from .types import OttersError

            raise OttersError(ffi.last_error())
        m = min(out_len.value, k)
        return idx[:m], score[:m], qid[:m]

    def search(self, vq, fp, k: int, take_max: bool):
        gathered, _ = self.enqueue(vq, fp, k)
        return self.merge(gathered, k, take_max)
