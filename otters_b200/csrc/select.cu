// select.cu — K3: final top-k selection over the per-CTA candidate lists of K1, the running-list
// merge for query batches, the record merge of the row-sharded multi-GPU path, and the full sort of
// the large-k (emit-all) path.
//
// Replaces TopKCollector::into_sorted_vec (reference src/vec_compute.rs:290-293) and the final merge
// of MetaQueryPlan::collect (src/meta.rs:699-708: concat, sort, truncate(k)).
//
// Canonical order everywhere: better score, then lower row, then lower query index.
#include "internal.h"
#include "select_body.cuh"

namespace otters {
namespace {

using namespace select_detail;

template <bool WITH_PREV>
__global__ void __launch_bounds__(1024, 1) select_kernel(const __grid_constant__ SelectParams p) {
    extern __shared__ __align__(16) uint8_t sm[];
    select_body<WITH_PREV>(p, sm);
}

// ---- sharded path: merge gathered records --------------------------------------------------------
__global__ void __launch_bounds__(1024, 1)
merge_records_kernel(const otters_topk_record* recs, uint32_t n, uint32_t k, uint32_t list_len, int take_max, Cand* out,
                     uint32_t* out_count, ResultHeader* hdr, const unsigned long long* rows_scored_src, uint64_t* scratch_keys,
                     uint32_t* scratch_src) {
    extern __shared__ __align__(16) uint8_t sm[];
    uint64_t* keys = reinterpret_cast<uint64_t*>(sm);
    uint32_t* src = reinterpret_cast<uint32_t*>(sm + (size_t)kSelectSmemElems * 8);
    __shared__ uint32_t s_valid;
    uint32_t P = next_pow2(n < 2 ? 2 : n);
    if (P > kSelectSmemElems) {
        keys = scratch_keys;
        src = scratch_src;
    }
    if (threadIdx.x == 0) s_valid = 0;
    __syncthreads();
    uint32_t loc = 0;
    for (uint32_t i = threadIdx.x; i < P; i += blockDim.x) {
        uint64_t key = 0ull;
        uint32_t tag = 0xFFFFFFFFu;
        if (i < n && recs[i].row != 0xFFFFFFFFFFFFFFFFull) {
            // global rows are < 2^32 (documented limit); ties: lower global row, then lower qid.
            key = make_key(recs[i].score, (uint32_t)recs[i].row, take_max != 0);
            tag = recs[i].qid;
            ++loc;
        }
        keys[i] = key;
        src[i] = tag;
    }
    if (loc) atomicAdd(&s_valid, loc);
    __syncthreads();
    if (n <= kRankMergeElems) {
        // few records (world * k): every thread counts the records ordered before its own — no sorting network.
        // When the input is a concatenation of ordered lists of list_len records (what otters_query_local_device writes,
        // one list per rank) the count is a binary search per list; the ordering is verified first.
        __shared__ uint32_t s_ordered;
        if (threadIdx.x == 0) s_ordered = (list_len && n % list_len == 0) ? 1u : 0u;
        __syncthreads();
        if (s_ordered)
            for (uint32_t e = threadIdx.x; e + 1 < n; e += blockDim.x)
                if ((e + 1) % list_len != 0 && before(keys[e + 1], src[e + 1], keys[e], src[e])) s_ordered = 0;
        __syncthreads();
        const bool ordered = s_ordered != 0;
        const uint32_t kr = s_valid < k ? s_valid : k;
        for (uint32_t e = threadIdx.x; e < n; e += blockDim.x) {
            const uint64_t key = keys[e];
            if (key == 0ull) continue;
            const uint32_t tag = src[e];
            uint32_t rank = 0;
            if (ordered) {
                // the records are `n / list_len` ordered lists (one per rank): binary search per list
                for (uint32_t l2 = 0; l2 < n / list_len; ++l2) {
                    const uint64_t* kp = keys + l2 * list_len;
                    const uint32_t* tp = src + l2 * list_len;
                    uint32_t lo = 0, hi = list_len;
                    while (lo < hi) {
                        const uint32_t mid = (lo + hi) >> 1;
                        if (before(kp[mid], tp[mid], key, tag)) lo = mid + 1;
                        else hi = mid;
                    }
                    rank += lo;
                }
            } else {
#pragma unroll 4
                for (uint32_t j = 0; j < n; ++j) rank += before(keys[j], src[j], key, tag);
            }
            if (rank < kr) {
                Cand c;
                c.key = key;
                c.qid = tag;
                c.pad = 0;
                out[rank] = c;
            }
        }
        if (threadIdx.x == 0) {
            *out_count = kr;
            if (hdr) {
                hdr->count = kr;
                hdr->rows_scored = rows_scored_src ? *rows_scored_src : 0ull;
                hdr->stats[0] = hdr->stats[1] = 0ull;
            }
        }
        return;
    }
    if (P > kSelectSmemElems) block_bitonic<true>(scratch_keys, scratch_src, P);
    else block_bitonic<true>(reinterpret_cast<uint64_t*>(sm), reinterpret_cast<uint32_t*>(sm + (size_t)kSelectSmemElems * 8), P);
    uint32_t kk = s_valid < k ? s_valid : k;
    for (uint32_t i = threadIdx.x; i < kk; i += blockDim.x) {
        Cand c;
        c.key = keys[i];
        c.qid = src[i];
        c.pad = 0;
        out[i] = c;
    }
    if (threadIdx.x == 0) {
        *out_count = kk;
        if (hdr) {
            hdr->count = kk;
            hdr->rows_scored = rows_scored_src ? *rows_scored_src : 0ull;
            hdr->stats[0] = hdr->stats[1] = 0ull;
        }
    }
}

// ---- large-k path: full sort of a candidate array -------------------------------------------------
__global__ void bitonic_step_kernel(Cand* buf, uint64_t n, uint64_t size, uint64_t stride) {
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (n >> 1)) return;
    uint64_t lo = 2 * t - (t & (stride - 1));
    uint64_t hi = lo + stride;
    bool fwd = (lo & size) == 0;
    Cand a = buf[lo], b = buf[hi];
    bool swap = fwd ? cand_before(b, a) : cand_before(a, b);
    if (swap) {
        buf[lo] = b;
        buf[hi] = a;
    }
}

// sorts 4096-element blocks entirely in shared memory for all (size, stride) with stride < 4096
__global__ void __launch_bounds__(1024) bitonic_local_kernel(Cand* buf, uint64_t n, uint64_t size_begin, uint64_t size_end) {
    __shared__ Cand s[2048];
    const uint64_t base = (uint64_t)blockIdx.x * 2048;
    for (uint32_t i = threadIdx.x; i < 2048; i += blockDim.x) s[i] = buf[base + i];
    __syncthreads();
    for (uint64_t size = size_begin; size <= size_end; size <<= 1) {
        uint64_t st0 = size >> 1;
        if (st0 > 1024) st0 = 1024;
        for (uint64_t stride = st0; stride > 0; stride >>= 1) {
            uint32_t t = threadIdx.x;
            uint32_t lo = 2 * t - (t & ((uint32_t)stride - 1));
            uint32_t hi = lo + (uint32_t)stride;
            bool fwd = ((base + lo) & size) == 0;
            Cand a = s[lo], b = s[hi];
            bool swap = fwd ? cand_before(b, a) : cand_before(a, b);
            if (swap) {
                s[lo] = b;
                s[hi] = a;
            }
            __syncthreads();
        }
    }
    for (uint32_t i = threadIdx.x; i < 2048; i += blockDim.x) buf[base + i] = s[i];
    (void)n;
}

__global__ void append_prev_kernel(Cand* buf, const uint32_t* emit_count, const Cand* prev, const uint32_t* prev_count,
                                   uint64_t n_pow2) {
    const uint64_t ne = *emit_count;
    const uint64_t np = (prev && prev_count) ? *prev_count : 0;
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t strideg = (uint64_t)gridDim.x * blockDim.x;
    for (; i < n_pow2; i += strideg) {
        if (i < ne) continue;
        Cand c;
        if (i < ne + np) c = prev[i - ne];
        else {
            c.key = 0ull;
            c.qid = 0xFFFFFFFFu;
            c.pad = 0;
        }
        buf[i] = c;
    }
}

__global__ void take_sorted_kernel(const Cand* buf, const uint32_t* emit_count, const uint32_t* prev_count, uint64_t k,
                                   Cand* out, uint32_t* out_count, uint64_t* tau_out, ResultHeader* hdr,
                                   const unsigned long long* rows_scored_src, const unsigned long long* stats_src,
                                   const unsigned long long* extra_src) {
    const uint64_t total = (uint64_t)*emit_count + (prev_count ? *prev_count : 0);
    const uint64_t kk = total < k ? total : k;
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t strideg = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t j = i; j < kk; j += strideg) out[j] = buf[j];
    if (i == 0) {
        *out_count = (uint32_t)kk;
        if (tau_out) *tau_out = (kk == k && kk > 0) ? buf[kk - 1].key : 0ull;
        if (hdr) {
            hdr->count = (uint32_t)kk;
            hdr->rows_scored = rows_scored_src ? *rows_scored_src : 0ull;
            hdr->stats[0] = stats_src ? stats_src[0] : 0ull;
            hdr->stats[1] = stats_src ? stats_src[1] : 0ull;
            hdr->extra[0] = extra_src ? extra_src[0] : 0ull;
            hdr->extra[1] = extra_src ? extra_src[1] : 0ull;
        }
    }
}

__global__ void cands_to_records_kernel(const Cand* cands, const uint32_t* count, uint32_t k, ShardMap map, int take_max,
                                        otters_topk_record* recs) {
    uint32_t n = *count;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < k; i += gridDim.x * blockDim.x) {
        otters_topk_record r;
        if (i < n) {
            r.row = map.global_row(key_row(cands[i].key));
            r.score = key_score(cands[i].key, take_max != 0);
            r.qid = cands[i].qid;
        } else {
            r.row = 0xFFFFFFFFFFFFFFFFull;
            r.score = 0.f;
            r.qid = 0;
        }
        recs[i] = r;
    }
}

}  // namespace

int launch_select(const SelectParams& p, cudaStream_t s) {
    static thread_local int configured_dev = -1;
    int dev = 0;
    cudaGetDevice(&dev);
    if (configured_dev != dev) {
        OTTERS_CUDA(cudaFuncSetAttribute(select_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSelectSmemBytes));
        OTTERS_CUDA(cudaFuncSetAttribute(select_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSelectSmemBytes));
        configured_dev = dev;
    }
    // block size ~ half the first working set (one compare-exchange per thread and step)
    const uint32_t n_lists = p.n_lists + 1;
    const uint32_t L = (2 * p.k + n_lists - 1) / n_lists + 3;
    uint32_t P = 2;
    while (P < n_lists * L) P <<= 1;
    uint32_t block = P / 2 < 128 ? 128 : (P / 2 > 1024 ? 1024 : P / 2);
    if (!p.prev || p.ex_world > 1) block = 1024;  // rank-counting paths: one element per thread
    if (p.prev) select_kernel<true><<<1, block, kSelectSmemBytes, s>>>(p);
    else select_kernel<false><<<1, block, kSelectSmemBytes, s>>>(p);
    OTTERS_CUDA(cudaGetLastError());
    return OTTERS_OK;
}

int launch_merge_records(const otters_topk_record* recs, uint32_t n, uint32_t k, int take_max, Cand* out, uint32_t* out_count,
                         ResultHeader* hdr, const unsigned long long* rows_scored_src, uint64_t* scratch_keys, uint32_t* scratch_src,
                         uint32_t scratch_elems, cudaStream_t s) {
    uint32_t P = 2;
    while (P < n) P <<= 1;
    if (P > kSelectSmemElems && P > scratch_elems) return fail(OTTERS_ERR_UNSUPPORTED, "merge: too many records");
    OTTERS_CUDA(
        cudaFuncSetAttribute(merge_records_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSelectSmemBytes));
    // k records per rank, each rank's records ordered best-first (otters_query_local_device): n == world * k
    merge_records_kernel<<<1, 1024, kSelectSmemBytes, s>>>(recs, n, k, k, take_max, out, out_count, hdr, rows_scored_src, scratch_keys,
                                                           scratch_src);
    OTTERS_CUDA(cudaGetLastError());
    return OTTERS_OK;
}

int launch_global_sort(Cand* buf, uint64_t n, cudaStream_t s) {
    // n is a power of two >= 2048
    const uint64_t blocks_local = n / 2048;
    bitonic_local_kernel<<<(unsigned)blocks_local, 1024, 0, s>>>(buf, n, 2, 2048);
    for (uint64_t size = 4096; size <= n; size <<= 1) {
        for (uint64_t stride = size >> 1; stride >= 2048; stride >>= 1) {
            uint64_t threads = n >> 1;
            bitonic_step_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, s>>>(buf, n, size, stride);
        }
        bitonic_local_kernel<<<(unsigned)blocks_local, 1024, 0, s>>>(buf, n, size, size);
    }
    OTTERS_CUDA(cudaGetLastError());
    return OTTERS_OK;
}

int launch_append_prev(Cand* buf, const uint32_t* emit_count, const Cand* prev, const uint32_t* prev_count, uint64_t n_pow2,
                       cudaStream_t s) {
    append_prev_kernel<<<1024, 256, 0, s>>>(buf, emit_count, prev, prev_count, n_pow2);
    OTTERS_CUDA(cudaGetLastError());
    return OTTERS_OK;
}

int launch_take_sorted(const Cand* buf, const uint32_t* emit_count, const uint32_t* prev_count, uint64_t k, Cand* out,
                       uint32_t* out_count, uint64_t* tau_out, ResultHeader* hdr, const unsigned long long* rows_scored_src,
                       const unsigned long long* stats_src, cudaStream_t s, const unsigned long long* extra_src) {
    take_sorted_kernel<<<256, 256, 0, s>>>(buf, emit_count, prev_count, k, out, out_count, tau_out, hdr, rows_scored_src, stats_src,
                                           extra_src);
    OTTERS_CUDA(cudaGetLastError());
    return OTTERS_OK;
}

int launch_cands_to_records(const Cand* cands, const uint32_t* count, uint32_t k, ShardMap map, int take_max,
                            otters_topk_record* recs, cudaStream_t s) {
    cands_to_records_kernel<<<(k + 255) / 256 ? (k + 255) / 256 : 1, 256, 0, s>>>(cands, count, k, map, take_max, recs);
    OTTERS_CUDA(cudaGetLastError());
    return OTTERS_OK;
}

}  // namespace otters
