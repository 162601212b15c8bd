#!/usr/bin/env python3
"""bench.py — headline benchmark of the otters exact-search hot path on B200.

A "step" is ONE query (one batch for --workload c2) through the hot path over the resident store: zonemap/Bloom chunk
pruning -> per-row predicate -> streaming scan -> top-k [-> exchange of k records per rank + merge at N > 1], all inside one
kernel launch per query for single queries.

Default workload (`--workload target`, BASELINE.json north_star "Target"): MetaStore 10M x 768 fp32, chunk_size 1024,
Cosine top-100 with meta_filter(price.gt & item.eq & ts.gte).  At N > 1 the SAME rows are dealt block-cyclically over the
ranks (strong scaling); the exchange of the k candidate records is fused into the query kernel (peer stores over NVLink),
with an NCCL all-gather + merge kernel as the portable path.

  value        : queries/s of the device pipeline: queries enqueued back to back on the context's two lanes
                 (otters_query_submit, no per-step host wait), store and scratch resident, timed with CUDA events
  e2e          : queries/s through the public non-blocking API with HOST buffers: otters_query_submit (host query in, one
                 H2D copy of the input image) / otters_query_wait (host top-k out), two queries in flight; the blocking
                 drop-in call (otters_metastore_query & co) is reported beside it as e2e.blocking_*
  roofline     : the scan kernel: algorithmic bytes (rows scored x (dim*4 [+4 cosine])) / its CUDA-event time, live
  cpu_baseline : the C oracle (port of the reference's CPU path) on this box's cores, bounded stratified row sample
  parity_check : every returned (row, score) re-derived by the oracle from that row alone, predicate / threshold / order
                 checks, completeness against the oracle's own top-k over a row sample, and a digest that must agree across N

`--impl reference` times that CPU path as its own arm (rank 0 only; never imports otters_b200).
"""
from __future__ import annotations

import argparse
import ctypes as C
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import bench_workloads as bw  # noqa: E402  (NumPy only)
from bench_workloads import DATA_SEED, QUERY_SEED, T0_MS, WORKLOADS, Workload, synth_fill_np  # noqa: E402,F401


# ---- compatibility helpers used by tests/test_gpu_fullsize.py ---------------------------------------------------------
def meta_columns(ob, rows, chunk):
    wl = Workload("target")
    wl.chunk = chunk
    return [c.to_ob(ob) for c in wl.columns(rows)]


def meta_expr(ob, rows_total):
    wl = Workload("target", rows_total)
    return wl.expr(ob), wl.filter_desc()


class ClockSampler:
    """Samples SM clocks and throttle reasons with nvidia-smi during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.proc, self.lines = device, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.device}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def measured_peak_tflops():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["bf16_tflops"]), "measured (MEASURED_PEAKS.json bf16_tflops, dense bf16 burst)"
    except Exception:
        return 1590.0, "fallback (B200_PROFILING.md 1.59 PFLOP/s dense bf16)"


def measured_traffic(workload, world):
    """dram bytes per scan launch from the committed ncu capture of this workload on ONE GPU (null for a shard)."""
    if world != 1:
        return None
    try:
        with open(os.path.join(ROOT, "profiles", "scan_traffic.json")) as f:
            return json.load(f).get(workload)
    except Exception:
        return None


def l2_note(wl):
    store_bytes = wl.rows * wl.dim * 4
    if store_bytes > 4 * 126e6:
        return "no flush needed: every step streams a multi-GB store (%.2f GB of rows), far larger than the 126 MB L2" % (store_bytes / 1e9)
    return "NOT flushed: the %.0f MB store is L2-resident between steps (latency-bound case; the HBM roofline does not apply)" % (store_bytes / 1e6)


def usable_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


# ---------------------------------------------------------------------------------------------------------
# CPU arm: the oracle (C port of the reference's CPU path) on a bounded, stratified row sample.
# Never imports otters_b200: columns are bench_workloads.SpecColumn, the filter is built from the CNF clauses.
# ---------------------------------------------------------------------------------------------------------
class CpuSample:
    """The oracle's store over a stratified sample of the workload's rows."""

    def __init__(self, wl, blocks=None):
        from oracle import oracle as ora

        self.ora, self.wl = ora, wl
        self.blocks = blocks or bw.stratified_sample_blocks(wl.rows, wl.chunk, wl.dim, wl.nq)
        self.global_rows = np.concatenate([np.arange(a, b, dtype=np.int64) for a, b in self.blocks])
        self.n = len(self.global_rows)
        self.vectors = np.concatenate([ora.synth_fill(a, b - a, wl.dim, DATA_SEED) for a, b in self.blocks], axis=0)
        prow, pvec = wl.planted()
        if len(prow):
            pos = np.searchsorted(self.global_rows, prow)
            hit = (pos < self.n) & (self.global_rows[np.minimum(pos, self.n - 1)] == prow)
            self.vectors[pos[hit]] = pvec[hit]
        self.vectors = wl.stored(self.vectors)
        self.queries = wl.queries()
        self.cores_all = usable_cores()
        self.kept_fraction = None
        if wl.meta:
            self.cols = wl.columns(self.global_rows)
            self.store = ora.MetaStore(self.vectors, self.cols, wl.chunk)
            idx = {c.name(): i for i, c in enumerate(self.cols)}
            self.fp = ora.FilterPack([[(idx[n], op, kind, val) for n, op, kind, val in cl] for cl in wl.clauses()])
            self.kept_fraction = float(wl.row_mask(self.cols).mean())
            self.cores = self.cores_all  # MetaQueryPlan::collect is rayon-parallel over chunks (src/meta.rs:678-691)
        else:
            self.inv = ora.inv_norms(self.vectors)
            self.cores = 1  # VecQueryPlan::collect is single-threaded (src/vec.rs:222-267)

    def query(self, variant, mode=None):
        """(global rows, scores, query ids) of the sample's top-k for query variant `variant`."""
        ora, wl = self.ora, self.wl
        mode = ora.FAITHFUL if mode is None else mode
        q = self.queries[variant % len(self.queries)]
        tt = 1 if wl.take_max else 0
        if wl.meta:
            idx, sc, qid, _ = self.store.query(q, wl.metric_code, tt, wl.k, wl.vec_filter, self.fp, mode, self.cores)
        else:
            idx, sc, qid = ora.vecstore_query(self.vectors, q, wl.metric_code, tt, wl.k, wl.vec_filter, None, mode, self.inv)
        return self.global_rows[np.asarray(idx, np.int64)], sc, qid


def store_kept_fraction(wl):
    """Fraction of ALL rows of the workload that pass the metadata filter (NumPy over the generators, 1M rows at a time)."""
    if not wl.meta:
        return None
    kept = 0
    for r0 in range(0, wl.rows, 1 << 20):
        r1 = min(r0 + (1 << 20), wl.rows)
        kept += int(wl.row_mask(wl.columns(np.arange(r0, r1))).sum())
    return kept / wl.rows


def cpu_arm(wl, budget_s, steps=None, warmup=1, sample=None):
    sample = sample or CpuSample(wl)
    for i in range(warmup):
        sample.query(i)
    times = []
    t_start = time.perf_counter()
    i = 0
    while True:
        t0 = time.perf_counter()
        sample.query(i)
        times.append(time.perf_counter() - t0)
        i += 1
        if steps is not None and i >= steps:
            break
        if steps is None and (time.perf_counter() - t_start > budget_s or i >= 200):
            break
    scale = wl.rows / sample.n
    per_step_sample = float(np.mean(times))
    qps_full = wl.nq / (per_step_sample * scale)
    kept_store = store_kept_fraction(wl)
    kept = ("" if kept_store is None else
            f"; rows passing the metadata filter: {sample.kept_fraction:.4f} of the sample, {kept_store:.4f} of the store")
    return dict(value=qps_full, unit="queries/s", cores=sample.cores, kind="port",
                sample=f"{len(times)} {'batches of %d queries' % wl.nq if wl.nq > 1 else 'queries'} over a stratified sample of {sample.n} of {wl.rows} rows "
                       f"({len(sample.blocks)} chunk-aligned blocks spread evenly over the store, same generators / filter / planted rows), "
                       f"time scaled x{scale:.2f}{kept}; oracle = C port of the reference CPU path (no Rust toolchain here), "
                       f"gcc -O3 -mavx2 -ffp-contract=off, {sample.cores} thread(s) of {sample.cores_all} usable cores "
                       f"(thread count pinned explicitly; OMP_NUM_THREADS is ignored)",
                ms_per_step_sample=per_step_sample * 1e3, n=len(times), kept_fraction_sample=sample.kept_fraction,
                kept_fraction_store=kept_store), sample


def run_reference(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # torchrun exports OMP_NUM_THREADS=1: the reference's MetaStore path is rayon-parallel over all cores, so is this arm
    os.environ["OMP_NUM_THREADS"] = str(usable_cores())
    os.environ.pop("OMP_THREAD_LIMIT", None)
    t0 = time.perf_counter()
    res, _ = cpu_arm(wl, budget_s=60.0, steps=args.steps, warmup=max(args.warmup, 1))
    ms = 1e3 / res["value"] * wl.nq
    line = {
        "impl": "reference", "metric": "queries_per_sec", "value": res["value"], "unit": "queries/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": dict(wl.config(), l2=l2_note(wl)),
        "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": res["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": time.perf_counter() - t0,
        "note": "ms_per_step is the sample's measured time scaled to the full store (see cpu_baseline.sample)",
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------------
# parity check of the GPU arm's results (rank 0, every N)
# ---------------------------------------------------------------------------------------------------------
def parity_check(wl, results, sample):
    """results: list of (variant, rows u64, scores f32, qids u32) as returned to the host by the timed API.
    sample: a CpuSample (its oracle answers the same queries over a stratified row sample)."""
    from oracle import oracle as ora

    out = {"ok": True, "queries_checked": len(results), "results_checked": 0, "failures": []}
    prow, pvec = wl.planted()
    digest = hashlib.sha1()

    def fail(msg):
        out["ok"] = False
        if len(out["failures"]) < 8:
            out["failures"].append(msg)

    def row_vectors(rows):
        v = np.concatenate([ora.synth_fill(int(r), 1, wl.dim, DATA_SEED) for r in rows], axis=0) if len(rows) else np.zeros((0, wl.dim), np.float32)
        if len(prow) and len(rows):
            pos = np.searchsorted(prow, rows)
            hit = (pos < len(prow)) & (prow[np.minimum(pos, len(prow) - 1)] == rows)
            v[hit] = pvec[pos[hit]]
        return wl.stored(v)

    queries = wl.queries()
    tt = 1 if wl.take_max else 0
    for variant, rows, scores, qids in results:
        rows = np.asarray(rows, np.int64)
        scores = np.asarray(scores, np.float32)
        qids = np.asarray(qids, np.int64)
        digest.update(rows.astype(np.uint64).tobytes())
        digest.update(scores.view(np.uint32).tobytes())
        out["results_checked"] += len(rows)
        qv = queries[variant % len(queries)]
        # (1) every (row, query, score) re-derived from that row alone, bit-identical
        v = row_vectors(rows)
        for qi in np.unique(qids) if len(rows) else []:
            sel = np.nonzero(qids == qi)[0]
            idx, sc, _ = ora.vecstore_query(v[sel], qv[qi:qi + 1], wl.metric_code, 1, len(sel), None, None, ora.CANONICAL)
            want = np.zeros(len(sel), np.float32)
            want[np.asarray(idx, np.int64)] = sc
            same = (want.view(np.uint32) == scores[sel].view(np.uint32)) | ((want == 0) & (scores[sel] == 0))
            if not same.all():
                fail(f"variant {variant}: {int((~same).sum())} scores differ from the per-row oracle")
        # (2) order, uniqueness, threshold, predicate
        if len(scores) > 1 and not (np.all(scores[:-1] >= scores[1:]) if wl.take_max else np.all(scores[:-1] <= scores[1:])):
            fail(f"variant {variant}: not ordered best-first")
        if wl.nq == 1 and len(set(rows.tolist())) != len(rows):
            fail(f"variant {variant}: duplicate rows")
        if wl.vec_filter and len(scores) and not np.all(scores > np.float32(wl.vec_filter[0])):
            fail(f"variant {variant}: a score fails vec_filter")
        if wl.meta and len(rows):
            order = np.argsort(rows, kind="stable")
            keep = wl.row_mask(wl.columns(rows[order]))
            if not keep.all():
                fail(f"variant {variant}: {int((~keep).sum())} returned rows fail the metadata filter")
        # (3) completeness against the oracle's own top-k over the row sample: whatever the oracle finds in the sample that
        #     beats our last entry must be in our list, and our entries that lie in the sample must be in the oracle's list
        #     (unless they fall behind its k-th)
        srows, sscores, sqids = sample.query(variant, ora.CANONICAL)
        ours = {(int(r), int(qd)): float(s) for r, qd, s in zip(rows, qids, scores)}
        full = len(rows) >= min(wl.k, 1 << 62)
        last = float(scores[-1]) if len(scores) else None
        for r, s, qd in zip(srows, sscores, sqids):
            better = last is None or not full or (s > last if wl.take_max else s < last)
            if better and (int(r), int(qd)) not in ours:
                fail(f"variant {variant}: row {int(r)} (score {float(s)}) found by the oracle in the sample is missing")
                break
        in_sample = np.isin(rows, sample.global_rows)
        theirs = {(int(r), int(qd)) for r, qd in zip(srows, sqids)}
        slast = float(sscores[-1]) if len(sscores) >= wl.k else None
        for r, qd, s in zip(rows[in_sample], qids[in_sample], scores[in_sample]):
            ahead = slast is None or (s > slast if wl.take_max else s < slast)
            if ahead and (int(r), int(qd)) not in theirs:
                fail(f"variant {variant}: row {int(r)} is in the sample but the oracle did not return it")
                break
        # (4) C5: the answer is known by construction — the best k planted rows that pass both filters
        if len(prow):
            pidx, psc, _ = ora.vecstore_query(wl.stored(pvec), qv[:1], wl.metric_code, 1, len(prow), wl.vec_filter, None, ora.CANONICAL)
            keep = wl.row_mask(wl.columns(prow))[np.asarray(pidx, np.int64)] if wl.meta else np.ones(len(pidx), bool)
            exp_rows, exp_sc = prow[np.asarray(pidx, np.int64)][keep][: wl.k], np.asarray(psc)[keep][: wl.k]
            if not (np.array_equal(exp_rows, rows) and np.array_equal(exp_sc.view(np.uint32), scores.view(np.uint32))):
                fail(f"variant {variant}: result differs from the planted set's top-{wl.k} ({len(rows)} rows vs {len(exp_rows)} expected)")
            out["planted_expected"] = int(len(exp_rows))
    out["sample_rows"] = int(sample.n)
    out["digest"] = digest.hexdigest()
    out["how"] = ("per-row oracle re-score (bit-identical), order / uniqueness / vec_filter / metadata predicate on every returned row, "
                  "completeness vs the oracle's top-k over a stratified row sample" + (", equality with the planted set's answer" if len(prow) else "")
                  + "; digest = sha1(rows, score bits) of the checked queries: identical at every N")
    return out


# ---------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------
def run_ours(args, wl):
    # libraries (NCCL's version banner, ...) may write to fd 1: park the real stdout and send everything else to stderr,
    # so that the ONE JSON line is all this process prints on stdout
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    import torch
    import torch.distributed as dist

    import otters_b200 as ob
    from otters_b200 import _ffi
    from otters_b200.sharded import CudaShard

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if rank == 0:
        # the CPU legs (cpu_baseline, parity check) use every core whatever torchrun exported
        os.environ["OMP_NUM_THREADS"] = str(usable_cores())
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    # a real (non-default) stream shared by torch (events, NCCL ordering) and the library's kernels
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    ctx = ob.Context(local_rank, stream.cuda_stream)
    tune = {}
    if args.tuning:
        w, s, kc, cps, ur = (int(x) for x in args.tuning.split(","))
        tune = dict(warps_per_cta=w, slots_per_warp=s, kc_floats=kc, ctas_per_sm=cps, unit_rows=ur)
    for name in ("scan_mode", "planners", "batch_passes", "separate_select", "lazy_prune", "disable_fused_predicate"):
        if getattr(args, name):
            tune[name] = getattr(args, name)
    ctx.set_tuning(**tune)

    rows, dim, chunk, k, nq = wl.rows, wl.dim, wl.chunk, wl.k, wl.nq
    metric = getattr(ob.Metric, wl.metric)
    tt = ob.TakeType.Max if wl.take_max else ob.TakeType.Min
    # block-cyclic row sharding: blocks of `block` rows are dealt round-robin to the ranks, so range filters
    # (ts >= cut) prune every shard equally
    block = wl.block
    n_local = bw.cyclic_local_rows(rows, block, world, rank)
    t_build = time.perf_counter()
    fp = None
    vfmt = ob.VectorFormat.Bf16 if wl.vector_format == "bf16" else ob.VectorFormat.F32
    if wl.meta:
        cols = [c.to_ob(ob) for c in wl.columns(bw.cyclic_global_rows(rows, block, world, rank))]
        store = (ob.MetaStore.from_columns(cols).with_synthetic_vectors(n_local, dim, DATA_SEED, 0, (world, rank, block))
                 .with_chunk_size(chunk).with_vector_format(vfmt).with_context(ctx).build())
        from otters_b200.meta import FilterPack

        fp = FilterPack(wl.expr(ob).compile(store.schema()), store.column_index())
    else:
        store = ob.VecStore(dim, ctx, vfmt)
        store.add_synthetic_sharded(world, rank, block, n_local, DATA_SEED)
    prow, pvec = wl.planted()
    if len(prow):
        mine, local_ids = bw.global_to_local(prow, block, world, rank)
        store.set_rows(local_ids, pvec[mine])
    build_s = time.perf_counter() - t_build

    queries = wl.queries()  # [variants][nq][dim]
    n_variants = len(queries)

    def build_vq(i):
        vq = _ffi.VecQuery()
        vq.queries = queries[i].ctypes.data_as(_ffi.c_f32p)
        vq.nq, vq.dim, vq.metric, vq.take_type, vq.k = nq, dim, int(metric), int(tt), k
        if wl.vec_filter:
            vq.has_filter, vq.thr, vq.cmp = 1, wl.vec_filter[0], wl.vec_filter[1]
        return vq

    # the query descriptors are built once: a step is the library call
    vqs = [build_vq(i) for i in range(n_variants)]

    def make_vq(i):
        return vqs[i % n_variants]

    shard = CudaShard(store, 0, k, block_rows=block)
    take_max = wl.take_max
    # N > 1: the exchange is fused into the query kernel (peer stores over NVLink + flags + merge, no NCCL call on the
    # query path) when the box can map peer memory; otherwise NCCL all-gather + merge kernel
    fused = world > 1 and nq == 1 and args.exchange != "nccl" and shard.enable_peer_exchange()
    # two queries in flight (submit / wait) wherever the exchange, if any, rides inside the query kernel
    pipelined = (world == 1 or fused) and not args.blocking

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- steps ---------------------------------------------------------------------------------------------
    def nccl_step(i, fetch):
        gathered, _ = shard.enqueue(make_vq(i), fp, k)
        if world > 1:
            return shard.merge(gathered, k, take_max, fetch=fetch)
        return None

    idx, sc = np.zeros(k, np.uint64), np.zeros(k, np.float32)
    qid_buf = np.zeros(k, np.uint32)
    qstats = _ffi.QueryStats()
    out_len = C.c_uint64()
    # output pointers are bound once: a step is the library call, not ctypes marshalling
    p_idx, p_sc, p_qid = idx.ctypes.data_as(_ffi.c_u64p), sc.ctypes.data_as(_ffi.c_f32p), qid_buf.ctypes.data_as(_ffi.c_u32p)
    p_len, p_stats = C.byref(out_len), C.byref(qstats)
    fp_ref = fp.byref() if fp else None
    vs_handle = None if wl.meta else store._handle()

    def blocking_step(i):
        """The blocking drop-in call: host query in -> host top-k out."""
        vq = make_vq(i)
        if world == 1:
            if wl.meta:
                rc = _ffi.otters_metastore_query(store.handle, C.byref(vq), fp_ref, p_idx, p_sc, p_qid, k, p_len, p_stats)
            else:
                rc = _ffi.otters_vecstore_query(vs_handle, C.byref(vq), p_idx, p_sc, p_qid, k, p_len)
            assert rc == 0, _ffi.last_error()
            m = min(out_len.value, k)
            return idx[:m], sc[:m], qid_buf[:m]
        if fused:
            out, _ = shard.search_fused(vq, fp, k)
            return out
        return nccl_step(i, True)

    def run_pipelined(first, count, collect=None):
        """count queries through submit / wait with two in flight; collect(i, result) sees every result."""
        t = shard.submit(make_vq(first), fp)
        for j in range(count):
            nxt = shard.submit(make_vq(first + j + 1), fp) if j + 1 < count else None
            res, _ = shard.wait(t, k)
            if collect is not None:
                collect(first + j, res)
            t = nxt

    if world > 1 and wl.meta:  # vectors_compared of this shard (summed over the ranks below)
        _, st0 = shard.enqueue(make_vq(0), fp, k, want_stats=True)
        qstats.vectors_compared = st0.vectors_compared
        torch.cuda.synchronize()

    # ---- parity: the first queries through the SAME calls the timed loops use ---------------------------------
    n_check = min(n_variants, 4)
    checked = []

    def keep_result(i, res):
        checked.append((i, res[0].copy(), res[1].copy(), res[2].copy()))

    if pipelined:
        run_pipelined(0, n_check, keep_result)
    else:
        for i in range(n_check):
            keep_result(i, blocking_step(i))
    barrier()

    # ---- e2e -------------------------------------------------------------------------------------------------
    sampler = ClockSampler(local_rank)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches = 0
    io_bytes = [0, 0]

    def count_work(_i=None, _res=None):
        nonlocal launches, io_bytes
        w = ctx.last_work()
        launches += int(w["kernel_launches"])
        io_bytes = [int(w["h2d_bytes"]), int(w["d2h_bytes"])]

    def timed(loop):
        barrier()
        t0 = time.perf_counter()
        ev0.record()
        loop()
        ctx.join()
        ev1.record()
        barrier()
        return ev0.elapsed_time(ev1) / args.steps, (time.perf_counter() - t0) * 1e3 / args.steps

    for i in range(args.warmup):
        blocking_step(i)
    if rank == 0:
        sampler.start()
    blk_ms, blk_wall_ms = timed(lambda: [(blocking_step(args.warmup + i), count_work()) for i in range(args.steps)])
    launches_blocking, launches = launches, 0
    if pipelined:
        run_pipelined(0, args.warmup)
        e2e_ms, e2e_wall_ms = timed(lambda: run_pipelined(args.warmup, args.steps, count_work))
    else:
        e2e_ms, e2e_wall_ms, launches = blk_ms, blk_wall_ms, launches_blocking

    # ---- roofline numerator: the same steps with the library's per-phase CUDA events switched on (they bracket the
    # kernels on the launching stream; kept out of the throughput loops because every event costs a few microseconds)
    scan_ms, scan_bytes, meta_bytes, rows_scored, batch_info, phase_ms = [], [], [], [], [], []
    ctx.set_tuning(**tune, timing=1)
    for i in range(min(args.steps, 30)):
        blocking_step(args.warmup + i)
        w = ctx.last_work()
        if w["scan_ms"] > 0:
            scan_ms.append(w["scan_ms"])
            scan_bytes.append(w["scan_bytes"])
        meta_bytes.append(w["meta_bytes"])
        phase_ms.append((w["prune_ms"], w["rowmask_ms"], w["scan_ms"], w["select_ms"]))
        rows_scored.append(w["rows_scored"])
        if nq > 1:
            batch_info.append((w["batch_used"], w["batch_fallback"], w["batch_max_err"], w["batch_delta"], w["select_ms"],
                               w["batch_passes"], w["batch_attempts"]))
    ctx.set_tuning(**tune)
    barrier()

    # ---- value: device pipeline only (no per-step host wait, results stay on the device) -------------------------
    def dev_loop(first, count):
        for j in range(count):
            if pipelined:
                shard.submit(make_vq(first + j), fp)
            elif fused:
                shard.search_fused(make_vq(first + j), fp, k, fetch=False)
            else:
                nccl_step(first + j, False)

    dev_loop(0, args.warmup)
    dev_ms, _ = timed(lambda: dev_loop(args.warmup, args.steps))
    ctx.synchronize()
    clocks = sampler.stop() if rank == 0 else None

    # max over ranks
    t = torch.tensor([dev_ms, e2e_ms, e2e_wall_ms, blk_ms, blk_wall_ms], dtype=torch.float64, device=dev)
    tot = torch.tensor([float(np.mean(rows_scored)) if rows_scored else 0.0, float(qstats.vectors_compared)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    dev_ms, e2e_ms, e2e_wall_ms, blk_ms, blk_wall_ms = (float(x) for x in t.tolist())
    rows_scored_total = float(tot[0].item())

    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        ach = (float(np.mean(scan_bytes)) / (float(np.mean(scan_ms)) * 1e-3) / 1e9) if scan_ms else 0.0
        cpu, sample = None, None
        if world == 1 and not args.no_cpu:
            c, sample = cpu_arm(wl, budget_s=args.cpu_budget)
            cpu = {kk: c[kk] for kk in ("value", "unit", "cores", "kind", "sample")}
        if sample is None:  # a small sample for the parity check alone
            budget = max(4_194_304 // nq // 4, 1) if nq > 1 else (131_072 if dim <= 768 else 65_536)
            sample = CpuSample(wl, bw.sample_blocks(rows, chunk, budget, 1 if nq > 1 else 10))
        parity = parity_check(wl, checked, sample) if not args.no_parity else None
        # bytes the library copied for one step (one input image: control block + lowered filter + queries; one result read)
        h2d, d2h = io_bytes
        roof = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak if peak else None,
                "traffic": measured_traffic(args.workload, world) if wl.vector_format == "f32" else None, "peak_source": peak_src,
                "kernel": "scan_kernel (K1: lazy chunk pruning + row predicate + scan + selection in its last CTA)",
                "scan_ms": float(np.mean(scan_ms)) if scan_ms else None,
                "algorithmic_bytes_per_launch": float(np.mean(scan_bytes)) if scan_bytes else None}
        if nq > 1:
            # tensor-bound: algorithmic flops = 2 * pairs * dim, counted ONCE (the 3xTF32 split issues 3x that on the pipe)
            tpeak, tsrc = measured_peak_tflops()
            flops = 2.0 * float(np.mean(rows_scored)) * dim
            tach = flops / (float(np.mean(scan_ms)) * 1e-3) / 1e12 if scan_ms else 0.0
            bi = np.array(batch_info, dtype=np.float64)
            roof = {"bound": "tensor", "achieved": tach, "peak": tpeak, "unit": "TFLOP/s", "frac": tach / tpeak if tpeak else None,
                    "traffic": measured_traffic(args.workload, world), "peak_source": tsrc,
                    "kernel": "batch_kernel (K2, tcgen05; rung = mma_passes: 2 = one kind::f16 MMA per product on bf16 shadow rows, "
                              "1 = one kind::tf32 MMA, 3 = 3xTF32 split)",
                    "scan_ms": float(np.mean(scan_ms)) if scan_ms else None, "algorithmic_flops_per_launch": flops,
                    "note": "peak is the measured dense bf16 rate: the bf16 rung (mma_passes 2) runs at it; kind::tf32 runs at half "
                            "of it, so the single-pass tf32 selection (mma_passes 1) has 1/2 of the peak as its ceiling, the 3xTF32 "
                            "split 1/6; every rung is selection only (certified + exactly re-scored from the fp32 rows)",
                    "mma_passes": float(bi[:, 5].mean()), "attempts_per_batch": float(bi[:, 6].mean()),
                    "tensor_path_used": float(bi[:, 0].mean()), "fallbacks": float(bi[:, 1].sum()),
                    "max_abs_err_vs_exact": float(bi[:, 2].max()), "assumed_err_bound": float(bi[:, 3].max()),
                    "rescore_sort_ms": float(bi[:, 4].mean())}
        api = ("otters_query_submit / otters_query_wait, two queries in flight" if pipelined else
               ("blocking otters_query_exchange" if fused else "blocking drop-in call" if world == 1 else "otters_query_local_device + NCCL all-gather + otters_topk_merge_device"))
        line = {
            "metric": "queries_per_sec", "value": nq * 1e3 / dev_ms, "unit": "queries/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32" if wl.vector_format == "f32" else "f32 arithmetic on bf16 rows", "data": "synthetic",
            "config": dict(wl.config(), l2=l2_note(wl)),
            "parallelism": f"rows block-cyclic ({block}-row blocks) over {world} GPU(s)" + (
                "" if world == 1 else (" + exchange fused into the query kernel (peer stores over NVLink, flags, merge)" if fused
                                       else " + NCCL all-gather of k records + device merge")),
            "rows_scored_per_query": rows_scored_total, "store_build_s": build_s,
            "e2e": {"value": nq * 1e3 / e2e_ms, "unit": "queries/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": e2e_ms, "wall_ms_per_step": e2e_wall_ms, "api": api,
                    "blocking_value": nq * 1e3 / blk_ms, "blocking_ms_per_step": blk_ms, "blocking_wall_ms_per_step": blk_wall_ms},
            "gpu_launches": launches,
            "phases_ms": dict(zip(("prune", "rowmask", "scan", "select"), (float(x) for x in np.mean(np.array(phase_ms), axis=0)))) if phase_ms else None,
            "roofline": roof,
            "rows_scored_per_sec": rows_scored_total * 1e3 / dev_ms,
            "vectors_compared_per_sec": (float(tot[1].item()) * 1e3 / dev_ms) if wl.meta else rows_scored_total * 1e3 / dev_ms,
            "cpu_baseline": cpu,
            "parity_check": parity,
            "clocks": clocks,
        }
        real_stdout.write(json.dumps(line) + "\n")
        real_stdout.flush()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="target", choices=sorted(WORKLOADS))
    ap.add_argument("--rows", type=int, default=0, help="override the row count (debugging only; invalid as a bench number)")
    ap.add_argument("--tuning", default="", help="warps,slots,kc,ctas_per_sm,unit_rows (0 = auto)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--blocking", action="store_true", help="e2e / value through the blocking calls only (no submit / wait)")
    ap.add_argument("--cpu-budget", type=float, default=15.0)
    ap.add_argument("--batch-passes", dest="batch_passes", type=int, default=0, help="K2: 0 auto (bf16 selection, then single-pass tf32, then 3xTF32), 1 (tf32 single pass), 2 (bf16) or 3")
    ap.add_argument("--scan-mode", dest="scan_mode", type=int, default=0, help="K1 front-end: 0 auto, 1 autonomous warps, 2 planner + workers")
    ap.add_argument("--planners", type=int, default=0, help="planner warps per CTA (planner front-end; 0 = auto)")
    ap.add_argument("--separate-select", dest="separate_select", type=int, default=0, help="1: K3 as its own kernel (A/B)")
    ap.add_argument("--lazy-prune", dest="lazy_prune", type=int, default=0, help="1: chunk pruning inside the scan kernel (A/B)")
    ap.add_argument("--unfused-predicate", dest="disable_fused_predicate", type=int, default=0,
                    help="1: row predicate in its own kernel (K0b row bitmask) instead of inside the scan (A/B)")
    ap.add_argument("--vector-format", dest="vector_format", default="f32", choices=["f32", "bf16"],
                    help="bf16: the store keeps bf16 rows (half the bytes per scan); both arms and the parity check score the rounded rows")
    ap.add_argument("--exchange", default="auto", choices=["auto", "nccl"], help="N > 1: fused peer-memory exchange when available, or force NCCL")
    args = ap.parse_args()
    wl = Workload(args.workload, args.rows, args.vector_format)
    if args.impl == "reference":
        run_reference(args, wl)
    else:
        run_ours(args, wl)


if __name__ == "__main__":
    main()
