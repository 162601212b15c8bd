// select_body.cuh — the body of K3 (final top-k selection over the per-CTA candidate lists of K1, optional running-list
// merge, optional fused peer exchange of the row-sharded search) as a device function, so that it can run either as its
// own kernel (select.cu) or in the LAST CTA of a scan kernel (one launch per query: prune + predicate + scan + select).
//
// Replaces TopKCollector::into_sorted_vec (reference src/vec_compute.rs:290-293) and the final merge of
// MetaQueryPlan::collect (src/meta.rs:699-708: concat, sort, truncate(k)).
// Canonical order everywhere: better score, then lower row, then lower query index.
#pragma once
#include "internal.h"

namespace otters {
namespace select_detail {

// element of the working set: a 64-bit candidate key plus a 32-bit source tag that both breaks ties
// (lower tag first) and lets the query index be recovered afterwards
__device__ __forceinline__ bool before(uint64_t ka, uint32_t sa, uint64_t kb, uint32_t sb) {
    return ka > kb || (ka == kb && sa < sb);
}

// block-wide bitonic sort of n (power of two) elements, best-first.  keys/src may live in shared or
// global memory (force-inlined so that the shared-memory instantiation compiles to LDS/STS).
// WITH_SRC = false sorts the keys alone (unique keys: no tie-break and no provenance needed).
template <bool WITH_SRC>
__device__ __forceinline__ void block_bitonic(uint64_t* keys, uint32_t* src, uint32_t n) {
    for (uint32_t size = 2; size <= n; size <<= 1) {
        for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
            for (uint32_t t = threadIdx.x; t < (n >> 1); t += blockDim.x) {
                uint32_t lo = 2 * t - (t & (stride - 1));
                uint32_t hi = lo + stride;
                bool fwd = (lo & size) == 0;
                uint64_t ka = keys[lo], kb = keys[hi];
                if (WITH_SRC) {
                    uint32_t sa = src[lo], sb = src[hi];
                    bool swap = fwd ? before(kb, sb, ka, sa) : before(ka, sa, kb, sb);
                    if (swap) {
                        keys[lo] = kb;
                        keys[hi] = ka;
                        src[lo] = sb;
                        src[hi] = sa;
                    }
                } else {
                    if (fwd ? (kb > ka) : (ka > kb)) {
                        keys[lo] = kb;
                        keys[hi] = ka;
                    }
                }
            }
            __syncthreads();
        }
    }
}

__device__ __forceinline__ uint32_t next_pow2(uint32_t x) {
    uint32_t p = 1;
    while (p < x) p <<= 1;
    return p;
}

// Source tags: list 0 is the running list of earlier queries (already ordered by key desc, qid asc,
// so position order == qid order among equal keys); lists 1.. are this query's per-CTA lists.
// tag = list << 11 | position  (position < 2048 because k <= 1024 in the fused path)
constexpr uint32_t kPosBits = 11;
constexpr uint32_t kRankSelectElems = 4096;  // largest working set served by the rank-counting fast path of the local selection
                                             // (keys in s_keys[0, 4096), compacted candidates in s_keys[4096, 8192), the ranked
                                             // result in the source-tag area, which the single-query path does not use)
constexpr uint32_t kRankMergeElems = 2048;   // stand-alone record merge (select.cu): all-pairs count below this size

__device__ __forceinline__ void st_release_sys_u32(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ otters_topk_record ld_record_volatile(const otters_topk_record* p) {
    uint32_t a, b, c, d;
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "l"(p) : "memory");
    otters_topk_record r;
    r.row = (uint64_t)a | ((uint64_t)b << 32);
    r.score = __uint_as_float(c);
    r.qid = d;
    return r;
}

__device__ __forceinline__ uint64_t ld_key_cg(const uint64_t* p) { return __ldcg(reinterpret_cast<const unsigned long long*>(p)); }

// WITH_PREV = false is the single-query hot path: no running list, all keys distinct, keys sorted alone.
// One CTA of any size runs this: either the stand-alone select_kernel (select.cu) or the LAST CTA of a scan kernel
// (scan_kernel.cuh / scan_planner.cu), which re-uses its dynamic shared memory (`sm`, at least kSelectSmemBytes).  The per-CTA
// lists were written by other CTAs (of this or an earlier kernel): they are read through L2 (ld.global.cg).
template <bool WITH_PREV>
__device__ __forceinline__ void select_body(const SelectParams& p, uint8_t* sm) {
    uint64_t* s_keys = reinterpret_cast<uint64_t*>(sm);
    uint32_t* s_src = reinterpret_cast<uint32_t*>(sm + (size_t)kSelectSmemElems * 8);
    __shared__ uint32_t s_total, s_nel, s_retry, s_maxcount, s_nc;
    __shared__ unsigned long long s_thr;

    const uint32_t first = WITH_PREV ? 0u : 1u;  // list 0 = running list
    const uint32_t n_lists = p.n_lists + 1;
    const uint32_t prev_n = (WITH_PREV && p.prev && p.prev_count) ? *p.prev_count : 0;
    auto list_count = [&](uint32_t l) -> uint32_t { return l == 0 ? prev_n : __ldcg(p.cta_counts + (l - 1)); };
    auto list_key = [&](uint32_t l, uint32_t i) -> uint64_t {
        return l == 0 ? p.prev[i].key : ld_key_cg(p.cta_keys + (size_t)(l - 1) * p.list_stride + i);
    };

    if (threadIdx.x == 0) {
        s_total = 0;
        s_maxcount = 0;
    }
    __syncthreads();
    {
        uint32_t loc = 0, mx = 0;
        for (uint32_t l = first + threadIdx.x; l < n_lists; l += blockDim.x) {
            uint32_t c = list_count(l);
            loc += c;
            mx = c > mx ? c : mx;
        }
        if (loc) atomicAdd(&s_total, loc);
        if (mx) atomicMax(&s_maxcount, mx);
    }
    __syncthreads();
    const uint32_t total = s_total;
    const uint32_t kk = total < p.k ? total : p.k;
    const uint32_t maxcount = s_maxcount;

    // Only a short prefix of every (sorted) list can reach the global top-k.  Start with a small
    // prefix length L, sort the union of prefixes, and grow L until no list was cut short.
    uint32_t L = (2 * p.k + n_lists - 1) / n_lists + 3;
    if (L > maxcount) L = maxcount;
    uint64_t* keys = s_keys;
    uint32_t* src = s_src;
    uint32_t P = 0;
    bool done = false;
    if (!WITH_PREV && L > 0 && p.n_lists * L <= kRankSelectElems) {
        // Fast path of the single-query case (all keys distinct): every thread loads one element of the union of
        // prefixes and finds its final position by counting the better elements — no sorting network, two barriers.
        const uint32_t nel_max = p.n_lists * L;
        uint64_t* ranked = reinterpret_cast<uint64_t*>(s_src);  // [kSelectSmemElems / 2]
        if (threadIdx.x == 0) {
            s_nel = 0;
            s_retry = 0;
        }
        __syncthreads();
        uint32_t loc = 0;
        for (uint32_t e = threadIdx.x; e < nel_max; e += blockDim.x) {
            const uint32_t l = e / L, i = e - l * L;
            const uint64_t key = i < __ldcg(p.cta_counts + l) ? ld_key_cg(p.cta_keys + (size_t)l * p.list_stride + i) : 0ull;
            s_keys[e] = key;
            loc += key != 0ull;
        }
        if (loc) atomicAdd(&s_nel, loc);
        __syncthreads();
        // (a) a lower bound on the k-th key from the list heads alone: if there are at least kk lists, the kk-th best
        //     HEAD has kk keys at or above it, so nothing below it can be part of the result
        uint64_t* cand = s_keys + kRankSelectElems;  // [kRankSelectElems] compacted candidates
        if (threadIdx.x == 0) {
            s_thr = 0ull;
            s_nc = 0;
        }
        __syncthreads();
        if (kk > 0 && kk <= p.n_lists) {
            for (uint32_t l = threadIdx.x; l < p.n_lists; l += blockDim.x) {
                const uint64_t key = s_keys[l * L];
                if (key == 0ull) continue;
                uint32_t rank = 0;
#pragma unroll 8
                for (uint32_t l2 = 0; l2 < p.n_lists; ++l2) rank += s_keys[l2 * L] > key;
                if (rank == kk - 1) s_thr = key;  // keys are distinct: exactly one head has this rank
            }
        }
        __syncthreads();
        // (b) compact the elements that can still matter
        const uint64_t thr = s_thr;
        for (uint32_t e = threadIdx.x; e < nel_max; e += blockDim.x) {
            const uint64_t key = s_keys[e];
            if (key != 0ull && key >= thr) cand[atomicAdd(&s_nc, 1u)] = key;
        }
        __syncthreads();
        // (c) final position = number of better candidates
        const uint32_t nc = s_nc;
        for (uint32_t e = threadIdx.x; e < nc; e += blockDim.x) {
            const uint64_t key = cand[e];
            uint32_t rank = 0;
#pragma unroll 8
            for (uint32_t j = 0; j < nc; ++j) rank += cand[j] > key;
            if (rank < kk) ranked[rank] = key;
        }
        __syncthreads();
        if (L < maxcount) {  // was any list cut short in a way that matters?
            if (s_nel < kk) {
                if (threadIdx.x == 0) s_retry = 1;
            } else if (kk > 0) {
                const uint64_t tk = ranked[kk - 1];
                for (uint32_t l = threadIdx.x; l < p.n_lists; l += blockDim.x)
                    if (__ldcg(p.cta_counts + l) > L && ld_key_cg(p.cta_keys + (size_t)l * p.list_stride + L) > tk) s_retry = 1;
            }
        }
        __syncthreads();
        if (!s_retry) {
            keys = ranked;
            done = true;
        } else {
            L = L * 4 < maxcount ? L * 4 : maxcount;
        }
        __syncthreads();
    }
    while (!done) {
        if (threadIdx.x == 0) {
            s_nel = 0;
            s_retry = 0;
        }
        __syncthreads();
        // gather prefixes
        uint64_t worst = (uint64_t)n_lists * L;
        bool use_global = worst > kSelectSmemElems;
        keys = use_global ? p.scratch_keys : s_keys;
        src = use_global ? p.scratch_src : s_src;
        for (uint32_t l = first + threadIdx.x; l < n_lists; l += blockDim.x) {
            uint32_t c = list_count(l);
            uint32_t take = c < L ? c : L;
            if (take) {
                uint32_t base = atomicAdd(&s_nel, take);
                for (uint32_t i = 0; i < take; ++i) {
                    keys[base + i] = list_key(l, i);
                    if (WITH_PREV) src[base + i] = (l << kPosBits) | i;
                }
            }
        }
        __syncthreads();
        const uint32_t nel = s_nel;
        P = next_pow2(nel < 2 ? 2 : nel);
        for (uint32_t i = nel + threadIdx.x; i < P; i += blockDim.x) {
            keys[i] = 0ull;
            if (WITH_PREV) src[i] = 0xFFFFFFFFu;
        }
        __syncthreads();
        if (use_global) block_bitonic<WITH_PREV>(p.scratch_keys, p.scratch_src, P);
        else block_bitonic<WITH_PREV>(s_keys, s_src, P);
        // was any list cut short in a way that matters?
        if (L < maxcount) {
            if (nel < kk) {
                if (threadIdx.x == 0) s_retry = 1;
            } else if (kk > 0) {
                const uint64_t tk = keys[kk - 1];
                const uint32_t ts = WITH_PREV ? src[kk - 1] : 0u;
                for (uint32_t l = first + threadIdx.x; l < n_lists; l += blockDim.x) {
                    uint32_t c = list_count(l);
                    if (c > L) {
                        uint64_t nk = list_key(l, L);
                        uint32_t ns = (l << kPosBits) | L;
                        if (WITH_PREV ? before(nk, ns, tk, ts) : (nk > tk)) s_retry = 1;
                    }
                }
            }
        }
        __syncthreads();
        if (!s_retry) break;
        L = L * 4 < maxcount ? L * 4 : maxcount;
        __syncthreads();
    }

    // emit the best kk, ordered
    Cand* const host_cands = p.host_out ? reinterpret_cast<Cand*>(p.host_out + sizeof(ResultHeader)) : nullptr;
    const bool exchange = p.ex_world > 1;
    const uint32_t par = p.ex_slot;  // which of the kExchangeSlots record / flag areas
    for (uint32_t i = threadIdx.x; i < (exchange ? p.ex_k : (p.records ? p.k : kk)); i += blockDim.x) {
        Cand c;
        c.key = 0ull;
        c.qid = p.qid;
        c.pad = 0;
        if (i < kk) {
            c.key = keys[i];
            if (WITH_PREV) {
                uint32_t s = src[i];
                uint32_t l = s >> kPosBits, pos = s & ((1u << kPosBits) - 1u);
                if (l == 0) c.qid = p.prev[pos].qid;
            }
            if (!exchange) {
                p.out[i] = c;
                if (host_cands) host_cands[i] = c;
            }
        }
        if (exchange || p.records) {
            otters_topk_record r;
            r.row = 0xFFFFFFFFFFFFFFFFull;
            r.score = 0.f;
            r.qid = 0;
            if (i < kk) {
                r.row = p.map.global_row(key_row(c.key));
                r.score = key_score(c.key, p.take_max != 0);
                r.qid = c.qid;
            }
            if (p.records) p.records[i] = r;
            if (exchange)  // peer stores over NVLink (the own area included)
                for (uint32_t pr = 0; pr < p.ex_world; ++pr) p.ex_records[pr][((size_t)par * p.ex_world + p.ex_rank) * p.ex_kmax + i] = r;
        }
    }
    uint32_t n_out = kk;
    if (exchange) {
        // publish: every record store above is ordered before the flag by the system-scope fence + release
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x < p.ex_world) st_release_sys_u32(p.ex_flags[threadIdx.x] + par * p.ex_world + p.ex_rank, p.ex_seq);
        // wait for the records of every rank for this query
        if (threadIdx.x < p.ex_world) {
            const uint32_t* f = p.ex_flags[p.ex_rank] + par * p.ex_world + threadIdx.x;
            const long long t0 = clock64();
            while (ld_acquire_sys_u32(f) != p.ex_seq) {
                if (clock64() - t0 > 200000000000ll) __trap();  // a rank never arrived (~100 s): fail instead of hanging forever
            }
        }
        __syncthreads();
        // merge world * k records (same order everywhere: better score, lower global row, lower query index)
        const otters_topk_record* mine = p.ex_records[p.ex_rank] + (size_t)par * p.ex_world * p.ex_kmax;
        const uint32_t n_in = p.ex_world * p.ex_k;
        if (threadIdx.x == 0) s_nel = 0;
        __syncthreads();
        uint32_t loc = 0;
        // every rank's k records arrive ordered, so a record's final position is a binary search per rank: this serves any
        // world * k that fits the working set (8 x 1024 records: ~20 us; the bitonic network it replaces took ~400 us with
        // the few hundred threads of a scan CTA)
        uint32_t P2 = n_in <= kSelectSmemElems ? n_in : next_pow2(n_in);
        for (uint32_t e = threadIdx.x; e < P2; e += blockDim.x) {
            uint64_t key = 0ull;
            uint32_t tag = 0xFFFFFFFFu;
            if (e < n_in) {
                const otters_topk_record r = ld_record_volatile(mine + (size_t)(e / p.ex_k) * p.ex_kmax + e % p.ex_k);
                if (r.row != 0xFFFFFFFFFFFFFFFFull) {
                    key = make_key(r.score, (uint32_t)r.row, p.take_max != 0);
                    tag = r.qid;
                    ++loc;
                }
            }
            s_keys[e] = key;
            s_src[e] = tag;
        }
        if (loc) atomicAdd(&s_nel, loc);
        __syncthreads();
        n_out = s_nel < p.ex_k ? s_nel : p.ex_k;
        if (n_in <= kSelectSmemElems) {
            for (uint32_t e = threadIdx.x; e < n_in; e += blockDim.x) {
                const uint64_t key = s_keys[e];
                if (key == 0ull) continue;
                const uint32_t tag = s_src[e];
                // every rank's records arrive ordered: binary search per rank instead of comparing against all records
                uint32_t rank = 0;
                for (uint32_t r2 = 0; r2 < p.ex_world; ++r2) {
                    const uint64_t* kp = s_keys + r2 * p.ex_k;
                    const uint32_t* tp = s_src + r2 * p.ex_k;
                    uint32_t lo = 0, hi = p.ex_k;
                    while (lo < hi) {
                        const uint32_t mid = (lo + hi) >> 1;
                        if (before(kp[mid], tp[mid], key, tag)) lo = mid + 1;
                        else hi = mid;
                    }
                    rank += lo;
                }
                if (rank < n_out) {
                    Cand c;
                    c.key = key;
                    c.qid = tag;
                    c.pad = 0;
                    p.out[rank] = c;
                    if (host_cands) host_cands[rank] = c;
                }
            }
        } else {
            block_bitonic<true>(s_keys, s_src, P2);
            for (uint32_t i = threadIdx.x; i < n_out; i += blockDim.x) {
                Cand c;
                c.key = s_keys[i];
                c.qid = s_src[i];
                c.pad = 0;
                p.out[i] = c;
                if (host_cands) host_cands[i] = c;
            }
        }
    }
    if (threadIdx.x == 0) {
        *p.out_count = n_out;
        if (p.tau_out) *p.tau_out = (!exchange && kk == p.k && kk > 0) ? keys[kk - 1] : 0ull;
        ResultHeader h;
        h.count = n_out;
        h.pad = 0;
        h.rows_scored = p.rows_scored_src ? __ldcg(p.rows_scored_src) : 0ull;
        h.stats[0] = p.stats_src ? __ldcg(p.stats_src) : 0ull;
        h.stats[1] = p.stats_src ? __ldcg(p.stats_src + 1) : 0ull;
        h.stats[2] = h.stats[3] = 0ull;
        h.extra[0] = h.extra[1] = 0ull;
        if (p.hdr) *p.hdr = h;
        if (p.host_out) *reinterpret_cast<ResultHeader*>(p.host_out) = h;
    }
}

}  // namespace select_detail
}  // namespace otters
