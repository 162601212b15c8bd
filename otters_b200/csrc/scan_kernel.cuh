// scan_kernel.cuh — K1: the streaming exact-order scan kernel (sm_100a).
//
// Replaces the scan loop of VecQueryPlan::collect (reference src/vec.rs:222-303) and the scoring
// kernels of src/vec_compute.rs:9-54, fused with the vec_filter threshold (src/vec_compute.rs:56-74)
// and a per-CTA top-k (replacing TopKCollector, src/vec_compute.rs:77-294).
//
// Design:
//  * persistent CTAs; every warp is an autonomous pipeline: it claims 128-row work units from a
//    global counter, turns the unit's surviving-row bitmask into a compact row list, and streams the
//    surviving rows HBM -> shared memory with per-row TMA bulk copies (cp.async.bulk + mbarrier
//    complete_tx) into a ring of `slots` tiles of 16 rows x kc columns.  Masked rows are never read
//    (as in src/vec.rs:248-252).
//  * arithmetic is bit-identical to the reference's CPU path: two threads per row hold the eight
//    f32x8 lane accumulators (4 each), multiply and add are separate round-to-nearest operations
//    (no FMA), blocks of 8 columns are accumulated in order, lanes are reduced in wide's order
//    ((l0+l1)+l2)+l3 + ((l4+l5)+l6)+l7, the dim%8 tail is a serial sum added last.
//  * staged rows use a shared-memory pitch == 8 (mod 32) floats, so the 8 threads of a quarter warp
//    (4 rows x 2 halves) hit 8 distinct 16-byte bank groups: conflict-free LDS.128.
//  * candidates that beat the CTA's running threshold key are appended lock-free to a shared buffer
//    (one atomicAdd reserves the slots); the warp whose reservation crosses the capacity bitonic-sorts
//    the buffer, keeps the best k and raises the threshold.  At exit every CTA publishes its best k
//    keys (sorted) for K3.
#pragma once
#include "internal.h"
#include "predicate.cuh"
#include "scan_shared.cuh"
#include "select_body.cuh"

namespace otters {
namespace scan_impl {

using namespace scan_detail;

template <int METRIC, bool EMIT_ALL, bool HALF>
__global__ void __launch_bounds__(512, 1) scan_kernel(const __grid_constant__ ScanParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;

    CtaHdr* hdr = reinterpret_cast<CtaHdr*>(smem);
    uint64_t* cbuf = reinterpret_cast<uint64_t*>(smem + 16);
    float* qs = reinterpret_cast<float*>(smem + p.off_query);
    uint8_t* wbase = smem + p.off_warps + (size_t)warp * p.warp_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(wbase);
    uint32_t* slot_rows = reinterpret_cast<uint32_t*>(wbase + p.off_w_rows);
    uint32_t* slot_info = reinterpret_cast<uint32_t*>(wbase + p.off_w_info);
    float* slot_inv = reinterpret_cast<float*>(wbase + p.off_w_inv);  // per-slot inverse norms of the tile's rows
    uint8_t* rowlist = wbase + p.off_w_list;
    float* slot_base = reinterpret_cast<float*>(wbase + p.off_w_slots);
    const uint32_t slot_floats = kTileRows * p.pitch_s;

    const uint64_t tau0 = p.tau_in ? *p.tau_in : 0ull;
    if (tid == 0) {
        hdr->tau = tau0;
        hdr->count = 0;
        hdr->written = 0;
    }
    for (uint32_t i = tid; i < p.dim_pad; i += blockDim.x) qs[i] = p.query[i];
    // stage the lowered filter (leaves + clause offsets) in shared memory
    const DevLeaf* f_leaves = reinterpret_cast<const DevLeaf*>(smem + p.off_filter);
    const uint32_t* f_off = reinterpret_cast<const uint32_t*>(smem + p.off_filter + (size_t)p.flt_n_leaves * sizeof(DevLeaf));
    if (p.flt_leaves) {
        uint32_t* dst = reinterpret_cast<uint32_t*>(smem + p.off_filter);
        const uint32_t words = p.flt_n_leaves * (uint32_t)(sizeof(DevLeaf) / 4);
        for (uint32_t i = tid; i < words; i += blockDim.x) dst[i] = reinterpret_cast<const uint32_t*>(p.flt_leaves)[i];
        for (uint32_t i = tid; i <= p.flt_n_clauses; i += blockDim.x) dst[words + i] = p.flt_clause_off[i];
    }
    if (lane == 0) {
        for (uint32_t s = 0; s < p.slots; ++s) mbar_init(&bars[s], 1);
        fence_mbar_init();
    }
    __syncthreads();

    const bool take_max = p.take_max != 0;
    const float q_inv = p.q_inv;
    const uint32_t dim8 = p.dim & ~7u;
    const uint32_t ntail = p.dim & 7u;
    const uint64_t l2pol = policy_evict_first();

    // ---- producer state ----
    UnitClaims claims;  // lane 0: the next unit ids, claimed ahead of time (scan_shared.cuh)
    unsigned long long g_pref = 0;  // lane 0: grid-wide threshold read together with the unit id
    claims.prime(p.unit_counter, p.claim_depth, lane);
    uint32_t list_n = 0, list_pos = 0, unit_row0 = 0, kc_i = 0, tile_cnt = 0;
    bool prod_done = false;
    uint32_t prod_step = 0, cons_step = 0;
    unsigned long long scored = 0;
    unsigned long long st_chunks = 0, st_vecs = 0;  // lazy pruning: chunks kept / rows of kept chunks x queries (per lane)

    // The row list of the NEXT unit is built while a tile of the current unit is in flight: claiming the unit, reading its
    // chunk bits and evaluating the row predicate cost two to three dependent memory round trips, which used to sit between
    // the last tile of one unit and the first tile of the next with nothing in flight for this warp (32-row units of a small
    // shard: ~15 % of the kernel; narrow filtered rows: ~35 %).  Two row-list buffers alternate.
    uint32_t nlist_n = 0, nunit_row0 = 0, cur_buf = 0;
    bool nvalid = false, claim_done = false;
    auto prepare = [&]() {
        uint8_t* nlist = rowlist + (cur_buf ^ 1u) * kMaxUnitRows;
        while (!nvalid) {
            uint32_t u = claims.front();
            if (u >= p.n_units) {
                claim_done = true;
                return;
            }
            if (lane == 0) {
                // adopt what the other CTAs have found so far (once per round of claims; the value was read a round ago)
                if (!EMIT_ALL && !claims.phase && g_pref > ld_volatile_u64(&hdr->tau)) atomicMax(&hdr->tau, g_pref);
                claims.refill(p.unit_counter);
                if (!EMIT_ALL && !claims.phase && p.g_tau) g_pref = *reinterpret_cast<volatile unsigned long long*>(p.g_tau);
            }
            claims.advance(p.claim_depth);
            // guided schedule: the first n_big units are unit_rows long, the rest (the tail of the store, claimed last)
            // unit_small long, so that the warps run out of work within one SMALL unit of each other
            const bool big = u < p.n_big;
            const uint32_t urows = big ? p.unit_rows : p.unit_small;
            uint32_t row0 = big ? u * p.unit_rows : p.n_big * p.unit_rows + (u - p.n_big) * p.unit_small;
            const uint32_t rpl = urows >= 32 ? urows >> 5 : 1;  // rows per lane when building the row list (1, 2 or 4)
            uint32_t r = row0 + rpl * lane;  // this lane's first row; its rpl rows share one mask word
            uint32_t bits = (1u << rpl) - 1u;
            if (rpl * lane >= urows) bits = 0;  // 16-row units: the upper half of the warp has no row
            if (p.row_mask) {
                uint32_t w = (r >> 5) < p.row_mask_words ? __ldg(p.row_mask + (r >> 5)) : 0xFFFFFFFFu;
                bits &= w >> (r & 31);
            }
            if (r >= p.n_rows) bits = 0;
            else if (p.n_rows - r < rpl) bits &= (1u << (p.n_rows - r)) - 1u;
            if (p.flt_leaves && r < p.n_rows) {
                uint32_t out = 0;
                if (p.chunk_keep) {
                    // fused K0b: chunk bits from the prune kernel, then the CNF over the rows' metadata with EVERY value and
                    // null-word load of the lane's (up to 4) rows in flight together (rows_pass_mlp): evaluated row after row,
                    // the four rows cost four dependent HBM round trips per unit — a third of the kernel on narrow rows
                    uint32_t km = 0, ch_prev = 0xFFFFFFFFu, kp = 0;
#pragma unroll
                    for (uint32_t j = 0; j < 4; ++j) {
                        if (j < rpl && r + j < p.n_rows) {
                            const uint32_t ch = (r + j) / p.chunk_size;
                            if (ch != ch_prev) {
                                kp = (__ldg(p.chunk_keep + (ch >> 5)) >> (ch & 31)) & 1u;
                                ch_prev = ch;
                            }
                            km |= kp << j;
                        }
                    }
                    if (p.pred_seq) {
                        const uint32_t live = bits & km;
                        for (uint32_t j = 0; j < rpl; ++j)
                            if (((live >> j) & 1u) && row_passes(f_leaves, f_off, p.flt_n_clauses, r + j)) out |= 1u << j;
                    } else {
                        out = rows_pass_mlp(f_leaves, f_off, p.flt_n_clauses, p.flt_n_leaves, r, bits & km, rpl);
                    }
                } else if (rpl * lane < urows) {
                    // fused K0 + K0b (lazy pruning): the zonemap / Bloom rules of the chunk(s) this lane's rows fall into are
                    // evaluated right here (uniform addresses across the warp for the usual chunk >= unit case: broadcast
                    // loads that hit L2), and the lane holding a chunk's FIRST row accounts the chunk in the statistics —
                    // every row of the store belongs to exactly one lane of one unit, so every chunk is counted once
                    // (src/meta.rs:666-669, src/meta_compute.rs:166)
                    uint32_t km = 0, ch_prev = 0xFFFFFFFFu;
                    bool kp = false;
#pragma unroll
                    for (uint32_t j = 0; j < 4; ++j) {
                        if (j < rpl && r + j < p.n_rows) {
                            const uint32_t row = r + j;
                            const uint32_t ch = row / p.chunk_size;
                            if (ch != ch_prev) {
                                kp = chunk_passes(f_leaves, f_off, p.flt_n_clauses, ch);
                                ch_prev = ch;
                            }
                            if (kp) {
                                km |= 1u << j;
                                if ((uint64_t)row == (uint64_t)ch * p.chunk_size) {  // first row of its chunk
                                    const uint64_t ch_end = (uint64_t)(ch + 1) * p.chunk_size;
                                    st_chunks += 1ull;
                                    st_vecs += (unsigned long long)(ch_end <= p.n_rows ? p.chunk_size : p.n_rows - row) * p.nq_stats;
                                }
                            }
                        }
                    }
                    out = rows_pass_mlp(f_leaves, f_off, p.flt_n_clauses, p.flt_n_leaves, r, bits & km, rpl);
                }
                bits = out;
            }
            uint32_t c = __popc(bits);
            uint32_t incl = c;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                uint32_t t = __shfl_up_sync(FULL, incl, d);
                if (lane >= d) incl += t;
            }
            uint32_t pos = incl - c;
            __syncwarp();
            while (bits) {
                int b = __ffs(bits) - 1;
                bits &= bits - 1;
                nlist[pos++] = (uint8_t)(rpl * lane + b);
            }
            const uint32_t n_keep = __shfl_sync(FULL, incl, 31);
            __syncwarp();
            if (n_keep) {
                nlist_n = n_keep;
                nunit_row0 = row0;
                nvalid = true;
            }
        }
    };

    auto issue = [&]() {
        if (prod_done) return;
        if (kc_i == 0) {
            if (list_pos >= list_n) {
                if (!nvalid && !claim_done) prepare();  // nothing prefetched yet (kernel start, or only empty units so far)
                if (!nvalid) {
                    prod_done = true;
                    return;
                }
                cur_buf ^= 1u;
                list_n = nlist_n;
                list_pos = 0;
                unit_row0 = nunit_row0;
                nvalid = false;
            }
            tile_cnt = list_n - list_pos < kTileRows ? list_n - list_pos : kTileRows;
        }
        const uint32_t slot = prod_step % p.slots;
        uint32_t row = 0xFFFFFFFFu;
        if (lane < (int)tile_cnt) row = unit_row0 + rowlist[cur_buf * kMaxUnitRows + list_pos + lane];
        if (lane < (int)kTileRows) slot_rows[slot * kTileRows + lane] = row;
        const uint32_t c0 = kc_i * p.kc;
        const uint32_t ncols = p.dim_pad - c0 < p.kc ? p.dim_pad - c0 : p.kc;
        const uint32_t bytes = ncols * (HALF ? 2u : 4u);
        if (lane == 0) {
            slot_info[slot] = tile_cnt | (kc_i << 8);
            mbar_arrive_expect_tx(&bars[slot], tile_cnt * bytes);
        }
        __syncwarp();
        if (lane < (int)tile_cnt) {
            bulk_g2s_hint(slot_base + (size_t)slot * slot_floats + (size_t)lane * p.pitch_s,
                          row_src<HALF>(p.vectors, p.pitch_g, row, c0), bytes, &bars[slot], l2pol);
            // the row's precomputed inverse norm rides along as a 4-byte cp.async (LDGSTS): its latency
            // overlaps the bulk copy instead of being exposed in the epilogue
            if (METRIC == OTTERS_METRIC_COSINE && kc_i == 0) cp_async_4(&slot_inv[slot * kTileRows + lane], p.inv_norms + row);
        }
        if (++kc_i == p.nkc) {
            kc_i = 0;
            list_pos += kTileRows;
        }
        ++prod_step;
    };

    // ---- consumer state ----
    const int r = lane >> 1;  // row of the tile handled by this thread pair
    const int h = lane & 1;   // which half of the 8 lanes: h=0 -> l0..l3, h=1 -> l4..l7
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    uint32_t my_row = 0xFFFFFFFFu;
    float rinv = 0.f;

    for (uint32_t s = 0; s < p.slots; ++s) issue();
    if (!p.no_prefetch && !nvalid && !claim_done) prepare();

    while (cons_step < prod_step) {
        const uint32_t slot = cons_step % p.slots;
        mbar_wait(&bars[slot], (cons_step / p.slots) & 1u);
        const uint32_t info = slot_info[slot];
        const uint32_t cnt = info & 0xFFu;
        const uint32_t kci = info >> 8;
        if (kci == 0) {
            a0 = a1 = a2 = a3 = 0.f;
            my_row = slot_rows[slot * kTileRows + r];
            if (METRIC == OTTERS_METRIC_COSINE) {
                cp_async_wait_all();
                __syncwarp();
                rinv = (r < (int)cnt) ? slot_inv[slot * kTileRows + r] : 0.f;
            }
        }
        const uint32_t c0 = kci * p.kc;
        const uint32_t cend = c0 + p.kc < dim8 ? c0 + p.kc : dim8;
        const uint32_t nblk = cend > c0 ? (cend - c0) >> 3 : 0;
        const float* vrow = slot_base + (size_t)slot * slot_floats + (size_t)r * p.pitch_s;
        const float4* qp = reinterpret_cast<const float4*>(qs + c0) + h;
#pragma unroll 4
        for (uint32_t j = 0; j < nblk; ++j) {
            const float4 v = load_row4<HALF>(vrow, 2 * j + h);
            const float4 q = qp[2 * j];
            if (METRIC == OTTERS_METRIC_EUCLIDEAN) {
                // src/vec_compute.rs:35-54: diff = query - row; acc += diff*diff
                const float d0 = __fsub_rn(q.x, v.x), d1 = __fsub_rn(q.y, v.y), d2 = __fsub_rn(q.z, v.z), d3 = __fsub_rn(q.w, v.w);
                a0 = __fadd_rn(a0, __fmul_rn(d0, d0));
                a1 = __fadd_rn(a1, __fmul_rn(d1, d1));
                a2 = __fadd_rn(a2, __fmul_rn(d2, d2));
                a3 = __fadd_rn(a3, __fmul_rn(d3, d3));
            } else {
                // src/vec_compute.rs:9-22: acc += q*v (multiply, then add)
                a0 = __fadd_rn(a0, __fmul_rn(q.x, v.x));
                a1 = __fadd_rn(a1, __fmul_rn(q.y, v.y));
                a2 = __fadd_rn(a2, __fmul_rn(q.z, v.z));
                a3 = __fadd_rn(a3, __fmul_rn(q.w, v.w));
            }
        }
        if (kci + 1 == p.nkc) {
            // wide f32x8::reduce_add (non-AVX build): (((l0+l1)+l2)+l3) + (((l4+l5)+l6)+l7)
            float sdot = __fadd_rn(__fadd_rn(__fadd_rn(a0, a1), a2), a3);
            float other = __shfl_xor_sync(FULL, sdot, 1);
            float tot = h == 0 ? __fadd_rn(sdot, other) : __fadd_rn(other, sdot);
            // serial remainder (src/vec_compute.rs:15-21), Rust's f32 Sum starts at -0.0
            float tail = -0.0f;
            if (ntail) {
                const float* qt = qs + dim8;
                for (uint32_t e = 0; e < ntail; ++e) {
                    const float ve = load_row1<HALF>(vrow, dim8 - c0 + e);
                    if (METRIC == OTTERS_METRIC_EUCLIDEAN) {
                        float d = __fsub_rn(qt[e], ve);
                        tail = __fadd_rn(tail, __fmul_rn(d, d));
                    } else {
                        tail = __fadd_rn(tail, __fmul_rn(qt[e], ve));
                    }
                }
            }
            float score = __fadd_rn(tot, tail);
            if (METRIC == OTTERS_METRIC_COSINE) score = __fmul_rn(__fmul_rn(score, q_inv), rinv);  // src/vec_compute.rs:31
            bool ok = (h == 0) && (r < (int)cnt) && !(score != score);  // NaN never returned (src/vec_compute.rs:237-239)
            if (p.has_filter) ok = ok && score_passes(score, p.thr, p.cmp);
            const uint64_t key = make_key(score, my_row, take_max);
            scored += (h == 0 && r < (int)cnt) ? 1ull : 0ull;
            if (EMIT_ALL) {
                ok = ok && key > tau0;
                unsigned m = __ballot_sync(FULL, ok);
                if (m) {
                    uint32_t base = 0;
                    if (lane == 0) base = atomicAdd(p.emit_count, (uint32_t)__popc(m));
                    base = __shfl_sync(FULL, base, 0);
                    if (ok) {
                        uint32_t at = base + __popc(m & ((1u << lane) - 1u));
                        if (at < p.emit_cap) {
                            Cand c;
                            c.key = key;
                            c.qid = p.qid;
                            c.pad = 0;
                            p.emit[at] = c;
                        }
                    }
                }
            } else {
                ok = ok && key > ld_volatile_u64(&hdr->tau);
                if (__ballot_sync(FULL, ok)) warp_push(hdr, cbuf, p.cap, p.k, ok, key, lane, p.g_tau);
            }
        }
        __syncwarp();
        ++cons_step;
        issue();
        if (!p.no_prefetch && !nvalid && !claim_done) prepare();  // overlaps the copy that was just issued
    }

    if (p.rows_scored) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) scored += __shfl_xor_sync(FULL, scored, d);
        if (lane == 0 && scored) atomicAdd(p.rows_scored, scored);
    }
    if (p.flt_leaves && !p.chunk_keep && p.stats) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            st_chunks += __shfl_xor_sync(FULL, st_chunks, d);
            st_vecs += __shfl_xor_sync(FULL, st_vecs, d);
        }
        if (lane == 0 && st_chunks) {
            atomicAdd(&p.stats[0], st_chunks);
            atomicAdd(&p.stats[1], st_vecs);
        }
    }

    if (!EMIT_ALL) {
        if (p.fuse_select && lane == 0)
            for (uint32_t s = 0; s < p.slots; ++s) mbar_inval(&bars[s]);  // the shared memory is about to be repurposed
        __syncthreads();
        if (warp == 0) {
            const uint32_t cnt = hdr->count;  // every push has completed: cnt <= cap and written == cnt
            uint32_t sort_n = 32;  // only as much of the buffer as holds candidates (k = 1000: 2048 slots, usually a few dozen used)
            while (sort_n < cnt) sort_n <<= 1;
            warp_sort(cbuf, cnt, sort_n, lane);
            const uint32_t n = cnt < p.k ? cnt : p.k;
            for (uint32_t i = lane; i < n; i += 32) p.cta_keys[(size_t)blockIdx.x * p.k + i] = cbuf[i];
            if (lane == 0) p.cta_counts[blockIdx.x] = n;
        }
        if (p.fuse_select) {
            // K3 fused into the scan: the CTA that publishes its list LAST (ticket from a global counter; the lists, the
            // row / chunk counters and the ticket are ordered by the fences) selects the final top-k — and, in a row-sharded
            // search, exchanges it with the peers — with the threads and the shared memory it already owns.  The other
            // CTAs have exited by then, so the next query's scan (enqueued on the context's other lane) already fills
            // their SMs: selection and exchange overlap the next scan instead of sitting between two launches.
            __shared__ uint32_t s_last;
            __threadfence();
            __syncthreads();
            if (tid == 0) s_last = atomicAdd(p.done_counter, 1u) == gridDim.x - 1u;
            __syncthreads();
            if (s_last) {
                __threadfence();
                select_detail::select_body<false>(p.sel, smem);
            }
        }
    }
}

template <int METRIC, bool EMIT, bool HALF>
int launch_one_fmt(const ScanParams& p, const ScanLaunch& l, uint32_t* smem_configured, cudaStream_t s) {
    auto kern = scan_kernel<METRIC, EMIT, HALF>;
    // the opt-in shared-memory limit is sticky per function and device: raise it only when it grows
    static uint32_t limits[64];
    uint32_t& have = smem_limit_slot(limits);
    (void)smem_configured;
    if (l.smem_bytes > have) {
        OTTERS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)l.smem_bytes));
        have = l.smem_bytes;
    }
    kern<<<l.grid, l.block, l.smem_bytes, s>>>(p);
    OTTERS_CUDA(cudaGetLastError());
    return OTTERS_OK;
}

// one row format per translation unit (scan.cu: fp32 rows, scan_bf16.cu: bf16 rows) so that the twelve instantiations of this
// kernel compile in parallel
template <bool HALF>
int launch_scan_fmt(const ScanParams& p, const ScanLaunch& l, int metric, bool emit_all, uint32_t* smem_configured, cudaStream_t s) {
    switch (metric) {
    case OTTERS_METRIC_COSINE:
        return emit_all ? launch_one_fmt<OTTERS_METRIC_COSINE, true, HALF>(p, l, smem_configured, s)
                        : launch_one_fmt<OTTERS_METRIC_COSINE, false, HALF>(p, l, smem_configured, s);
    case OTTERS_METRIC_EUCLIDEAN:
        return emit_all ? launch_one_fmt<OTTERS_METRIC_EUCLIDEAN, true, HALF>(p, l, smem_configured, s)
                        : launch_one_fmt<OTTERS_METRIC_EUCLIDEAN, false, HALF>(p, l, smem_configured, s);
    case OTTERS_METRIC_DOT:
        return emit_all ? launch_one_fmt<OTTERS_METRIC_DOT, true, HALF>(p, l, smem_configured, s)
                        : launch_one_fmt<OTTERS_METRIC_DOT, false, HALF>(p, l, smem_configured, s);
    }
    return fail(OTTERS_ERR_INVALID, "Search metric is not set");
}

}  // namespace scan_impl
}  // namespace otters
