// scan_planner.cu — K1, planner front-end: the streaming exact-order scan with the row selection done AHEAD of the
// streaming warps (sm_100a).
//
// Same contract, arithmetic and outputs as scan_kernel.cuh (reference src/vec.rs:222-303, src/vec_compute.rs:9-294): the
// difference is who decides which rows are read.  In scan_kernel.cuh every warp is autonomous: it claims a 128-row unit,
// evaluates the row mask / CNF predicate for it, then streams the survivors; while it evaluates, its TMA slot is
// idle, and the last units of the dynamic schedule are finished by single warps.  Here one or two PLANNER warps per
// CTA claim the units, evaluate the chunk bit + predicate (all metadata loads of a unit in flight together) and
// publish 16-row tiles of surviving rows into a shared-memory ring; the WORKER warps only pop tiles and stream:
//   * no unit-boundary bubble in the streaming warps (this capped 128-d stores at 70 % of the HBM peak);
//   * the rows of a unit are spread over all warps of the CTA, so the tail of the schedule is one unit per CTA
//     (8 µs at 768 d) instead of one unit per warp (100 µs) — this is what a small shard of a row-sharded search needs.
// Ring protocol: planners reserve ticket ranges with one shared atomicAdd, wait until the previous occupant of each
// ticket's slot has been consumed (freed[slot] == ticket / ring size), fill it and publish state[slot] = ticket + 1;
// workers take tickets in order with one atomicAdd, wait for their slot's state, copy the 16 row ids and bump freed.  Producers only ever wait on OLDER tickets being consumed and
// consumers only on their own ticket being produced, so the protocol cannot deadlock.
#include "internal.h"
#include "predicate.cuh"
#include "scan_shared.cuh"
#include "select_body.cuh"

namespace otters {

namespace {

using namespace scan_detail;

constexpr uint32_t kRing = kPlannerRingTiles;

struct Ring {
    uint32_t head;      // next ticket handed to a worker
    uint32_t tail;      // next ticket handed to a planner
    uint32_t done;      // planner warps that have finished
    uint32_t pad;
    uint32_t state[kRing];  // ticket + 1 of the tile published in the slot
    uint32_t freed[kRing];  // how many tiles of this slot have been consumed: ticket t may be written when freed == t / kRing
    uint32_t cnt[kRing];    // rows in the tile (1..16)
    uint32_t rows[kRing * kTileRows];
};
static_assert(sizeof(Ring) <= kPlannerRingBytes, "ring does not fit its shared-memory reservation");

__device__ __forceinline__ void st_volatile_u32(uint32_t* p, uint32_t v) { *reinterpret_cast<volatile uint32_t*>(p) = v; }

template <int METRIC, bool HALF>
__global__ void __launch_bounds__(640, 1) scan_planner_kernel(const __grid_constant__ ScanParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const uint32_t np = p.planners;

    CtaHdr* hdr = reinterpret_cast<CtaHdr*>(smem);
    uint64_t* cbuf = reinterpret_cast<uint64_t*>(smem + 16);
    float* qs = reinterpret_cast<float*>(smem + p.off_query);
    Ring* ring = reinterpret_cast<Ring*>(smem + p.off_ring);

    const uint64_t tau0 = p.tau_in ? *p.tau_in : 0ull;
    if (tid == 0) {
        hdr->tau = tau0;
        hdr->count = 0;
        hdr->written = 0;
        ring->head = ring->tail = ring->done = 0;
    }
    for (uint32_t i = tid; i < kRing; i += blockDim.x) {
        ring->state[i] = 0;
        ring->freed[i] = 0;
    }
    for (uint32_t i = tid; i < p.dim_pad; i += blockDim.x) qs[i] = p.query[i];
    const DevLeaf* f_leaves = reinterpret_cast<const DevLeaf*>(smem + p.off_filter);
    const uint32_t* f_off = reinterpret_cast<const uint32_t*>(smem + p.off_filter + (size_t)p.flt_n_leaves * sizeof(DevLeaf));
    if (p.flt_leaves) {
        uint32_t* dst = reinterpret_cast<uint32_t*>(smem + p.off_filter);
        const uint32_t words = p.flt_n_leaves * (uint32_t)(sizeof(DevLeaf) / 4);
        for (uint32_t i = tid; i < words; i += blockDim.x) dst[i] = reinterpret_cast<const uint32_t*>(p.flt_leaves)[i];
        for (uint32_t i = tid; i <= p.flt_n_clauses; i += blockDim.x) dst[words + i] = p.flt_clause_off[i];
    }
    uint64_t* bar = nullptr;
    uint32_t* slot_rows = nullptr;
    float* slot_inv = nullptr;
    float* slot = nullptr;
    if (warp >= (int)np) {
        uint8_t* wbase = smem + p.off_warps + (size_t)(warp - np) * p.warp_bytes;
        bar = reinterpret_cast<uint64_t*>(wbase);
        slot_rows = reinterpret_cast<uint32_t*>(wbase + p.off_w_rows);
        slot_inv = reinterpret_cast<float*>(wbase + p.off_w_inv);
        slot = reinterpret_cast<float*>(wbase + p.off_w_slots);
        if (lane == 0) {
            mbar_init(bar, 1);
            fence_mbar_init();
        }
    }
    __syncthreads();

    unsigned long long scored = 0;
    if (warp < (int)np) {
        // =============================== planner ===============================
        UnitClaims claims;  // lane 0: the next unit ids, claimed ahead of time (scan_shared.cuh)
        unsigned long long g_pref = 0;  // grid-wide threshold, read together with the unit id
        unsigned long long st_chunks = 0, st_vecs = 0;  // lazy pruning: chunks kept / rows of kept chunks x queries (per lane)
        claims.prime(p.unit_counter, p.claim_depth, lane);
        for (;;) {
            const uint32_t u = claims.front();
            if (u >= p.n_units) break;
            if (lane == 0) {
                // adopt what the other CTAs have found so far (once per round of claims; the value was read a round ago)
                if (!claims.phase && g_pref > ld_volatile_u64(&hdr->tau)) atomicMax(&hdr->tau, g_pref);
                claims.refill(p.unit_counter);
                if (!claims.phase && p.g_tau) g_pref = *reinterpret_cast<volatile unsigned long long*>(p.g_tau);
            }
            claims.advance(p.claim_depth);
            // guided schedule (see scan_kernel.cuh): big units first, small units for the tail of the store
            const bool big = u < p.n_big;
            const uint32_t urows = big ? p.unit_rows : p.unit_small;
            const uint32_t row0 = big ? u * p.unit_rows : p.n_big * p.unit_rows + (u - p.n_big) * p.unit_small;
            const uint32_t rpl = urows >= 32 ? urows >> 5 : 1;  // rows per lane (1, 2 or 4); a lane's rows share one mask word
            const uint32_t r = row0 + rpl * lane;
            uint32_t bits = (1u << rpl) - 1u;
            if (rpl * lane >= urows) bits = 0;
            if (p.row_mask) {
                const uint32_t w = (r >> 5) < p.row_mask_words ? __ldg(p.row_mask + (r >> 5)) : 0xFFFFFFFFu;
                bits &= w >> (r & 31);
            }
            if (r >= p.n_rows) bits = 0;
            else if (p.n_rows - r < rpl) bits &= (1u << (p.n_rows - r)) - 1u;
            if (p.flt_leaves && r < p.n_rows && rpl * lane < urows) {  // (lanes past the unit's rows own nothing)
                // chunk bits from the prune kernel AND the CNF over the row's metadata (fused K0b).  The chunk words are
                // requested first but only used after the predicate, so that all loads of the unit — chunk words, null
                // words, column values — are in flight together: under a saturated HBM every dependent round trip costs
                // microseconds.  (Rows of pruned chunks get their metadata read for nothing: ~20 B against 0.5-6 KB per row.)
                uint32_t kwm = 0;
                if (p.chunk_keep) {
#pragma unroll
                    for (uint32_t j = 0; j < 4; ++j) {
                        if (j < rpl && r + j < p.n_rows) {
                            const uint32_t ch = (r + j) / p.chunk_size;
                            kwm |= ((__ldg(p.chunk_keep + (ch >> 5)) >> (ch & 31)) & 1u) << j;
                        }
                    }
                } else {
                    // lazy pruning (see scan_kernel.cuh): zonemap / Bloom rules of the chunks of this lane's rows, evaluated here; the
                    // lane holding a chunk's first row accounts it in the statistics
                    uint32_t ch_prev = 0xFFFFFFFFu;
                    bool keep_ch = false;
#pragma unroll
                    for (uint32_t j = 0; j < 4; ++j) {
                        if (j < rpl && r + j < p.n_rows) {
                            const uint32_t row = r + j;
                            const uint32_t ch = row / p.chunk_size;
                            if (ch != ch_prev) {
                                keep_ch = chunk_passes(f_leaves, f_off, p.flt_n_clauses, ch);
                                ch_prev = ch;
                            }
                            if (keep_ch) {
                                kwm |= 1u << j;
                                if ((uint64_t)row == (uint64_t)ch * p.chunk_size) {  // first row of its chunk
                                    st_chunks += 1ull;
                                    const uint64_t ch_end = (uint64_t)(ch + 1) * p.chunk_size;
                                    st_vecs += (unsigned long long)(ch_end <= p.n_rows ? p.chunk_size : p.n_rows - row) * p.nq_stats;
                                }
                            }
                        }
                    }
                }
                bits = rows_pass_mlp(f_leaves, f_off, p.flt_n_clauses, p.flt_n_leaves, r, bits, rpl);
                bits &= kwm;
            }
            const uint32_t c = __popc(bits);
            uint32_t incl = c;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t t = __shfl_up_sync(FULL, incl, d);
                if (lane >= d) incl += t;
            }
            uint32_t pos = incl - c;
            const uint32_t n = __shfl_sync(FULL, incl, 31);
            if (n == 0) continue;
            if (lane == 0) scored += n;
            const uint32_t need = (n + kTileRows - 1) / kTileRows;  // <= 8
            uint32_t t0 = 0;
            if (lane == 0) t0 = atomicAdd(&ring->tail, need);
            t0 = __shfl_sync(FULL, t0, 0);
            // wait for the slots of my tickets to be free (the tickets one ring earlier have been consumed)
            if (lane < (int)need) {
                const uint32_t s = (t0 + lane) % kRing;
                const uint32_t cycle = (t0 + lane) / kRing;
                const long long tw = clock64();
                while (ld_volatile_u32(&ring->freed[s]) != cycle) {
                    __nanosleep(20);
                    if (clock64() - tw > 4000000000ll) __trap();  // watchdog
                }
            }
            __syncwarp();
            while (bits) {
                const int b = __ffs(bits) - 1;
                bits &= bits - 1;
                ring->rows[((t0 + pos / kTileRows) % kRing) * kTileRows + pos % kTileRows] = r + b;
                ++pos;
            }
            if (lane < (int)need) {
                const uint32_t left = n - lane * kTileRows;
                ring->cnt[(t0 + lane) % kRing] = left < kTileRows ? left : kTileRows;
            }
            __syncwarp();
            __threadfence_block();
            if (lane < (int)need) st_volatile_u32(&ring->state[(t0 + lane) % kRing], t0 + lane + 1u);
        }
        __syncwarp();
        if (lane == 0) {
            __threadfence_block();
            atomicAdd(&ring->done, 1u);
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
    } else {
        // =============================== worker ===============================
        const bool take_max = p.take_max != 0;
        const float q_inv = p.q_inv;
        const uint32_t dim8 = p.dim & ~7u;
        const uint32_t ntail = p.dim & 7u;
        const uint64_t l2pol = policy_evict_first();
        const int r = lane >> 1;  // row of the tile handled by this thread pair
        const int h = lane & 1;   // which half of the 8 lanes: h=0 -> l0..l3, h=1 -> l4..l7
        uint32_t phase = 0;
        for (;;) {
            // ---- take the next tile ----
            uint32_t ticket = 0;
            if (lane == 0) ticket = atomicAdd(&ring->head, 1u);
            ticket = __shfl_sync(FULL, ticket, 0);
            const uint32_t rs = ticket % kRing;
            uint32_t have = 0;
            if (lane == 0) {
                const long long tw = clock64();
                for (;;) {
                    if (ld_volatile_u32(&ring->state[rs]) == ticket + 1u) {
                        have = 1;
                        break;
                    }
                    // every planner has finished and no ticket this high was ever handed out: no more work
                    if (ld_volatile_u32(&ring->done) == np && ticket >= ld_volatile_u32(&ring->tail)) break;
                    __nanosleep(20);
                    if (clock64() - tw > 4000000000ll) __trap();  // watchdog
                }
            }
            have = __shfl_sync(FULL, have, 0);
            if (!have) break;
            __threadfence_block();
            const uint32_t cnt = ring->cnt[rs];
            uint32_t row = 0xFFFFFFFFu;
            if (lane < (int)cnt) row = ring->rows[rs * kTileRows + lane];
            if (lane < (int)kTileRows) slot_rows[lane] = row;
            __syncwarp();
            if (lane == 0) st_volatile_u32(&ring->freed[rs], ticket / kRing + 1u);  // slot free for the next ring cycle

            // ---- stream and score it, kc columns at a time ----
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
            const uint32_t my_row = slot_rows[r];
            float rinv = 0.f;
            for (uint32_t kci = 0; kci < p.nkc; ++kci) {
                const uint32_t c0 = kci * p.kc;
                const uint32_t ncols = p.dim_pad - c0 < p.kc ? p.dim_pad - c0 : p.kc;
                const uint32_t bytes = ncols * (HALF ? 2u : 4u);
                if (lane == 0) mbar_arrive_expect_tx(bar, cnt * bytes);
                __syncwarp();
                if (lane < (int)cnt) {
                    bulk_g2s_hint(slot + (size_t)lane * p.pitch_s, row_src<HALF>(p.vectors, p.pitch_g, row, c0), bytes, bar, l2pol);
                    if (METRIC == OTTERS_METRIC_COSINE && kci == 0) cp_async_4(&slot_inv[lane], p.inv_norms + row);
                }
                mbar_wait(bar, phase);
                phase ^= 1u;
                if (kci == 0 && METRIC == OTTERS_METRIC_COSINE) {
                    cp_async_wait_all();
                    __syncwarp();
                    rinv = (r < (int)cnt) ? slot_inv[r] : 0.f;
                }
                const uint32_t cend = c0 + p.kc < dim8 ? c0 + p.kc : dim8;
                const uint32_t nblk = cend > c0 ? (cend - c0) >> 3 : 0;
                const float* vrow = slot + (size_t)r * p.pitch_s;
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
                    const float sdot = __fadd_rn(__fadd_rn(__fadd_rn(a0, a1), a2), a3);
                    const float other = __shfl_xor_sync(FULL, sdot, 1);
                    const float tot = h == 0 ? __fadd_rn(sdot, other) : __fadd_rn(other, sdot);
                    // serial remainder (src/vec_compute.rs:15-21), Rust's f32 Sum starts at -0.0
                    float tail = -0.0f;
                    if (ntail) {
                        const float* qt = qs + dim8;
                        for (uint32_t e = 0; e < ntail; ++e) {
                            const float ve = load_row1<HALF>(vrow, dim8 - c0 + e);
                            if (METRIC == OTTERS_METRIC_EUCLIDEAN) {
                                const float d = __fsub_rn(qt[e], ve);
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
                    ok = ok && key > ld_volatile_u64(&hdr->tau);
                    if (__ballot_sync(FULL, ok)) warp_push(hdr, cbuf, p.cap, p.k, ok, key, lane, p.g_tau);
                }
                __syncwarp();  // every lane is done with the slot before the next copy lands in it
            }
        }
    }

    if (p.rows_scored) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) scored += __shfl_xor_sync(FULL, scored, d);
        if (lane == 0 && scored) atomicAdd(p.rows_scored, scored);
    }
    if (p.fuse_select && bar && lane == 0) mbar_inval(bar);  // the shared memory is about to be repurposed
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
        // K3 fused into the scan (see scan_kernel.cuh): the CTA that publishes its list last selects the final top-k
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

template <int METRIC, bool HALF>
int launch_one_fmt(const ScanParams& p, const ScanLaunch& l, uint32_t* smem_configured, cudaStream_t s) {
    auto kern = scan_planner_kernel<METRIC, HALF>;
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

template <int METRIC>
int launch_one(const ScanParams& p, const ScanLaunch& l, uint32_t* smem_configured, cudaStream_t s) {
    return p.half ? launch_one_fmt<METRIC, true>(p, l, smem_configured, s) : launch_one_fmt<METRIC, false>(p, l, smem_configured, s);
}

}  // namespace

int launch_scan_planner(const ScanParams& p, const ScanLaunch& l, int metric, uint32_t* smem_configured, cudaStream_t s) {
    switch (metric) {
    case OTTERS_METRIC_COSINE: return launch_one<OTTERS_METRIC_COSINE>(p, l, smem_configured, s);
    case OTTERS_METRIC_EUCLIDEAN: return launch_one<OTTERS_METRIC_EUCLIDEAN>(p, l, smem_configured, s);
    case OTTERS_METRIC_DOT: return launch_one<OTTERS_METRIC_DOT>(p, l, smem_configured, s);
    }
    return fail(OTTERS_ERR_INVALID, "Search metric is not set");
}

}  // namespace otters
