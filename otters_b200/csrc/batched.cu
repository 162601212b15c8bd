// batched.cu — K2: query batches as a TMA-fed tcgen05 (UMMA) 3xTF32 tile contraction (sm_100a).
//
// Replaces, for nq > 1, the serial per-query inner loop of VecQueryPlan::collect (reference
// src/vec.rs:243-266: every 8-row block is scored against every query) together with the scoring
// kernels of src/vec_compute.rs:9-54.  A batch feeds ONE global collector (src/vec.rs:217-219), so the
// kernel keeps one candidate list per CTA over all (row, query) pairs it scores.
//
// Structure:
//   S[row, q] = sum_d V[row, d] * Q[q, d] is a dense contraction: 128 store rows (UMMA M, the TMEM
//   lanes) x 256 queries (UMMA N, TMEM columns) per tile, K streamed in blocks of 32 fp32 columns
//   (one 128-byte swizzle atom).  fp32 accuracy on tf32 tensor cores comes from the 3xTF32 split
//   x = hi + lo with hi = rna_tf32(x), lo = rna_tf32(x - hi):  S ~= Vlo*Qhi + Vhi*Qlo + Vhi*Qhi
//   (three tcgen05.mma kind::tf32 per k-step, fp32 accumulators in TMEM).
//   Warp roles: warp 0 TMA producer (cp.async.bulk.tensor 2D, 128B swizzle, mbarrier complete_tx),
//   warp 1 MMA issuer (one thread), warp 2 TMEM allocator, warps 4-7 split the landed V tile into
//   hi/lo in shared memory, warps 8-11 epilogue (tcgen05.ld -> metric -> loosened vec_filter ->
//   lock-free per-CTA top-k of APPROXIMATE scores).  The accumulator is double buffered in TMEM
//   (2 x 256 columns) so the epilogue of tile i overlaps the MMAs of tile i+1.
//   Exactness: tensor-core scores differ from the reference's 8-lane fp32 order in the last bits, so
//   K2 only SELECTS candidates; rescore_kernel then recomputes every surviving (row, query) pair in the
//   reference's exact arithmetic order (same code shape as K1), applies the exact vec_filter, and the
//   final order comes from the exact keys.  The host verifies that no excluded pair can reach the
//   result (approximate cut + error bound < exact k-th score) and otherwise falls back to K1 per query.
//   Rungs (BatchParams::passes): 2 = one kind::f16 MMA per k-step on a bf16 SHADOW of the store rows and bf16 queries
//   (half the operand bytes, twice the MMA rate; bound 2^-7 |q||v|), 1 = one kind::tf32 MMA on the fp32 rows (2^-9),
//   3 = the 3xTF32 split (2^-15).  A k-block is one 128-byte swizzle atom in every rung: 32 fp32 or 64 bf16 columns,
//   so stages, barriers and descriptors are byte-identical between rungs 1 and 2.
#include <cuda.h>  // CUtensorMap types only; the encoder is resolved through cudaGetDriverEntryPoint
#include <float.h>

#include "internal.h"

namespace otters {
namespace {

constexpr unsigned FULL = 0xFFFFFFFFu;
constexpr uint32_t BM = kBatchRows;     // 128 store rows per tile
constexpr uint32_t BN = kBatchQueries;  // 256 queries per tile
#ifndef OTTERS_BK
#define OTTERS_BK 32
#endif
constexpr uint32_t BK = OTTERS_BK;      // fp32 columns per k-block: 32 (one 128-byte swizzle atom) or 16 (64-byte atoms, twice the stages)
static_assert(BK == 32 || BK == 16, "k-block must be one 128-byte or one 64-byte swizzle atom");
constexpr uint32_t kSwizzleBytes = BK * 4;
// auxiliary shared memory behind the stages: mbarriers | TMEM base address | candidate-buffer header | candidates
constexpr uint32_t kAuxTmemSlot = 448, kAuxHdr = 512, kAuxCand = 640;
constexpr uint32_t UK = 8;              // K of one tcgen05.mma kind::tf32
constexpr uint32_t A_BYTES = BM * BK * 4;  // 16 KB: the V tile of one CTA and one k-block
constexpr uint32_t TMEM_COLS = 512;
constexpr uint32_t NUM_THREADS = 512;   // launch bound; the kernel is launched with 384 + 32 * (extra epilogue warps) threads
constexpr uint32_t BASE_WARPS = 8;      // warps 0-7: producer, MMA issuer, TMEM allocator, relay, four split warps

// ---- PTX wrappers ------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                     smem_u32(dst)),
                 "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
                 : "memory");
}
// CTA pairs: the same load, but its transaction bytes are credited to the barrier at this shared-memory offset in the
// pair's LEADER (bit 24 of a shared::cluster address is the CTA's rank inside the pair; clearing it addresses rank 0)
__device__ __forceinline__ void tma_load_2d_pair(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                     smem_u32(dst)),
                 "l"(map), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, both operands K-major, kind::tf32, fp32 accumulate
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrives on the mbarrier once every tcgen05.mma issued so far by this thread has completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// 32 consecutive TMEM columns of this thread's lane (lane = 32 * (warp % 4) + laneid)
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ float rna_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// Shared-memory matrix descriptor of a K-major tile whose rows are one 128-byte swizzle atom wide
// (rows 128 bytes apart, 8-row groups 1024 bytes apart; tile base 1024-byte aligned):
//   [0,14) start address >> 4 | [16,30) leading byte offset >> 4 (unused for swizzled K-major) |
//   [32,46) stride byte offset >> 4 = 1024 >> 4 | [46,48) version = 1 (sm_100) | [61,64) layout = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
    // stride byte offset = 8 rows of one swizzle atom; layout type 2 = SWIZZLE_128B, 4 = SWIZZLE_64B
    return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)((8u * kSwizzleBytes) >> 4) << 32) | (1ull << 46) |
           ((kSwizzleBytes == 128 ? 2ull : 4ull) << 61);
}
// Instruction descriptor, kind::tf32 (Geo<CG>::IDESC): [4,6) D format = 1 (f32) | [7,10) A format = 2 (tf32) |
// [10,13) B format = 2 | [15] A major = 0 (K) | [16] B major = 0 (K) | [17,23) N >> 3 | [24,29) M >> 4

struct CtaHdr {
    unsigned long long tau;  // candidates must have key > tau
    uint32_t count;          // slots reserved in the candidate buffer (may transiently exceed cap)
    uint32_t written;        // slots whose entry has been stored
    uint32_t excl;           // ord_f32 of the best signed score this CTA excluded from its list (0 = nothing excluded)
};
// The upper half of a key ("goodness": ord(score) for take Max, ~ord(score) for take Min) orders candidates by
// score alone; excl tracks its maximum over everything a CTA left out of its list.

__device__ __forceinline__ uint64_t ld_volatile_u64(const unsigned long long* p) {
    return *reinterpret_cast<const volatile unsigned long long*>(p);
}
__device__ __forceinline__ uint32_t ld_volatile_u32(const uint32_t* p) { return *reinterpret_cast<const volatile uint32_t*>(p); }

__device__ __forceinline__ bool before(uint64_t ka, uint32_t qa, uint64_t kb, uint32_t qb) {
    return ka > kb || (ka == kb && qa < qb);
}

// one warp: sort (key, qid) pairs best-first; entries >= cnt are cleared first
__device__ void warp_sort_pairs(uint64_t* keys, uint32_t* qids, uint32_t cnt, uint32_t cap, int lane) {
    for (uint32_t i = cnt + lane; i < cap; i += 32) {
        keys[i] = 0ull;
        qids[i] = 0xFFFFFFFFu;
    }
    __syncwarp();
    for (uint32_t size = 2; size <= cap; size <<= 1) {
        for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
            for (uint32_t t = lane; t < (cap >> 1); t += 32) {
                uint32_t lo = 2 * t - (t & (stride - 1));
                uint32_t hi = lo + stride;
                bool fwd = (lo & size) == 0;
                uint64_t ka = keys[lo], kb = keys[hi];
                uint32_t qa = qids[lo], qb = qids[hi];
                bool swap = fwd ? before(kb, qb, ka, qa) : before(ka, qa, kb, qb);
                if (swap) {
                    keys[lo] = kb;
                    keys[hi] = ka;
                    qids[lo] = qb;
                    qids[hi] = qa;
                }
            }
            __syncwarp();
        }
    }
}

// Lock-free append (same protocol as K1's warp_push, scan_shared.cuh) with a query-id payload.  After a compaction
// the CTA's new threshold is published to the grid-wide threshold (any key below some CTA's k-th best key
// cannot be among the global best k), and the grid-wide value is adopted when it is higher.
__device__ __noinline__ void warp_push_pairs(CtaHdr* hdr, uint64_t* keys, uint32_t* qids, uint32_t cap, uint32_t k, bool has, uint64_t key,
                                uint32_t qid, unsigned long long* g_tau, int lane) {
    for (;;) {
        const uint64_t tau = ld_volatile_u64(&hdr->tau);
        if (has && key <= tau) {  // rejected by the exact key test (rare): it counts as excluded
            atomicMax(&hdr->excl, (uint32_t)(key >> 32));
            has = false;
        }
        const unsigned m = __ballot_sync(FULL, has);
        const uint32_t n = __popc(m);
        if (!n) return;
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(&hdr->count, n);
        base = __shfl_sync(FULL, base, 0);
        if (base + n <= cap) {
            if (has) {
                const uint32_t at = base + __popc(m & ((1u << lane) - 1u));
                keys[at] = key;
                qids[at] = qid;
            }
            __syncwarp();
            if (lane == 0) {
                __threadfence_block();
                atomicAdd(&hdr->written, n);
            }
            return;
        }
        if (base <= cap) {
            // compactor: valid entries are [0, base)
            if (lane == 0) {
                const long long t0 = clock64();
                while (ld_volatile_u32(&hdr->written) != base) {
                    if (clock64() - t0 > 4000000000ll) __trap();  // watchdog
                }
            }
            __syncwarp();
            __threadfence_block();
            warp_sort_pairs(keys, qids, base, cap, lane);
            if (lane == 0) {
                if (base > k) {  // entries [k, base) are dropped: remember the best of them
                    atomicMax(&hdr->excl, (uint32_t)(keys[k] >> 32));
                }
                unsigned long long t = keys[k - 1];  // base > cap - 32 >= k
                const unsigned long long g = atomicMax(g_tau, t);
                if (g > t) t = g;
                if (t > hdr->tau) *reinterpret_cast<volatile unsigned long long*>(&hdr->tau) = t;
                *reinterpret_cast<volatile uint32_t*>(&hdr->written) = k;
                __threadfence_block();
                atomicExch(&hdr->count, k);
            }
            __syncwarp();
        } else {
            if (lane == 0) {
                const long long t0 = clock64();
                while (ld_volatile_u32(&hdr->count) > cap) {
                    __nanosleep(32);
                    if (clock64() - t0 > 4000000000ll) __trap();  // watchdog
                }
            }
            __syncwarp();
        }
    }
}

// loosened vec_filter on an approximate score: never rejects a pair whose exact score passes (|a - e| <= delta)
__device__ __forceinline__ bool score_passes_loose(float s, float thr, int cmp, float delta) {
    switch (cmp) {
    case 0: return s < thr + delta;
    case 1: return s > thr - delta;
    case 2: return s <= thr + delta;
    case 3: return s >= thr - delta;
    default: return fabsf(s - thr) <= delta;
    }
}

// metric epilogue on a tensor-core dot product a = <v, q>
template <int METRIC>
__device__ __forceinline__ float batch_score(float a, const float* q_scal, uint32_t qi, float rs) {
    if (METRIC == OTTERS_METRIC_COSINE) return a * __ldg(q_scal + qi) * rs;          // rs = 1/|v|, q_scal = 1/|q|
    if (METRIC == OTTERS_METRIC_EUCLIDEAN) return (__ldg(q_scal + qi) + rs) - 2.0f * a;  // rs = |v|^2, q_scal = |q|^2
    return a;
}
__device__ __forceinline__ uint32_t tmem_ld_32x32b_x1(uint32_t taddr) {
    uint32_t r;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    return r;
}

// Per-configuration geometry.  CG = 1: one CTA per tile of 128 rows x 256 queries.  CG = 2: a CTA pair
// (cluster of two, tcgen05 cta_group::2) per tile of 256 rows x 256 queries; each CTA stages its own 128 rows
// and HALF of the query tile, so its stage shrinks to 64 KB (3 stages instead of 2) and the shared-memory
// operand traffic per MMA halves.
template <int CG>
struct Geo {
    static constexpr uint32_t BN_LOAD = BN / CG;              // query rows staged per CTA
    static constexpr uint32_t B_BYTES = BN_LOAD * BK * 4;     // 32 KB / 16 KB
    static constexpr uint32_t STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;  // Vhi | Vlo | Qhi | Qlo
    static constexpr uint32_t STAGES = (192u * 1024u) / STAGE_BYTES;  // 2 / 3 at 128-byte k-blocks, 4 / 6 at 64-byte ones
    static constexpr uint32_t TILE_ROWS = BM * CG;
    static constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((BN >> 3) << 17) | ((TILE_ROWS >> 4) << 24);
    // kind::f16: same fields, A / B format = 1 (bf16), fp32 accumulate
    static constexpr uint32_t IDESC_BF16 = (1u << 4) | (1u << 7) | (1u << 10) | ((BN >> 3) << 17) | ((TILE_ROWS >> 4) << 24);
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the same barrier of CTA `cta` of the cluster (release at cluster scope)
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t cta) {
    asm volatile(
        "{\n\t.reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}"
        ::"r"(smem_u32(bar)), "r"(cta)
        : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait_cluster(bar, parity)) return;
    const long long t0 = clock64();
    uint32_t spins = 0;
    while (!mbar_try_wait_cluster(bar, parity)) {
        if ((++spins & 0x3FFFu) == 0 && clock64() - t0 > 4000000000ll) __trap();
    }
}
template <int CG>
__device__ __forceinline__ void umma_tf32_cg(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    if constexpr (CG == 1) {
        umma_tf32(tmem_d, adesc, bdesc, idesc, accumulate);
    } else {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
            : "memory");
    }
}
// the bf16 rung: kind::f16 with bf16 operands (K = 16 per instruction = the same 32 bytes per row as 8 tf32 columns)
template <int CG>
__device__ __forceinline__ void umma_bf16_cg(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    if constexpr (CG == 1) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
            : "memory");
    } else {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
            : "memory");
    }
}
// CG = 2: the arrival is multicast to the same barrier of both CTAs of the pair
template <int CG>
__device__ __forceinline__ void umma_commit_cg(uint64_t* bar) {
    if constexpr (CG == 1) {
        umma_commit(bar);
    } else {
        const uint16_t mask = 3;
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                         smem_u32(bar)),
                     "h"(mask)
                     : "memory");
    }
}
template <int CG>
__device__ __forceinline__ void tmem_alloc_cg(uint32_t* dst_smem, uint32_t ncols) {
    if constexpr (CG == 1) {
        tmem_alloc(dst_smem, ncols);
    } else {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
}
template <int CG>
__device__ __forceinline__ void tmem_dealloc_cg(uint32_t taddr, uint32_t ncols) {
    if constexpr (CG == 1) tmem_dealloc(taddr, ncols);
    else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// tile t -> (row tile, query tile): consecutive tiles share the row tile so that the CTAs working on it at the
// same time find the V tile in L2 after the first HBM read.  A row tile whose rows are all masked is skipped
// by every warp role (the test reads the same mask words everywhere).
template <int CG>
__device__ __forceinline__ bool tile_live(const BatchParams& p, uint32_t rt) {
    if (!p.row_mask) return true;
    const uint32_t w0 = rt * (Geo<CG>::TILE_ROWS / 32);
    uint32_t any = 0;
#pragma unroll
    for (uint32_t i = 0; i < Geo<CG>::TILE_ROWS / 32; ++i) any |= (w0 + i) < p.row_mask_words ? __ldg(p.row_mask + w0 + i) : 0xFFFFFFFFu;
    return any != 0;
}

template <int METRIC, int CG>
__global__ void __launch_bounds__(NUM_THREADS, 1)
batch_kernel(const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_qh,
             const __grid_constant__ CUtensorMap tm_ql, const __grid_constant__ BatchParams p) {
    using G = Geo<CG>;
    constexpr uint32_t B_BYTES = G::B_BYTES;
    constexpr uint32_t RING_BYTES = G::STAGES * G::STAGE_BYTES;
    constexpr uint32_t MAX_STAGES = RING_BYTES / (A_BYTES + B_BYTES);  // single-pass stages hold V | Qhi only
    // the ring is cut at run time: 3xTF32 stages are Vhi | Vlo | Qhi | Qlo, single-pass stages V | Qhi (twice as many)
    const bool single = p.passes != 3;  // one MMA per product (selection with a wider error bound) instead of the 3xTF32 split
    const bool half = p.passes == 2;    // ... on bf16 operands (tm_v / tm_qh then describe the bf16 shadow arrays)
    const uint32_t kcols = half ? 2u * BK : BK;  // columns of one k-block: one swizzle atom of bf16 or of fp32
    // single pass only: KPS k-blocks share one stage = one barrier round trip and one tcgen05.commit;
    // DIRECT (pairs): both CTAs' loads credit the leader's barrier themselves instead of going through the relay warp
    const uint32_t KPS = !single ? 1u : (p.kps < 1u ? 1u : (p.kps > MAX_STAGES / 2u ? MAX_STAGES / 2u : p.kps));  // at least two stages
    const bool direct = CG == 2 && single && p.pair_direct != 0u;
    const uint32_t KB_BYTES = A_BYTES + B_BYTES;  // one single-pass k-block: V | Qhi
    const uint32_t STAGE_BYTES = single ? KPS * KB_BYTES : G::STAGE_BYTES;
    const uint32_t STAGES = single ? MAX_STAGES / KPS : G::STAGES;
    const uint32_t Q_OFF = single ? A_BYTES : 2 * A_BYTES;  // Qhi inside a stage
    extern __shared__ uint8_t smem_raw[];
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const uint32_t rank = CG == 2 ? cluster_ctarank() : 0u;  // position in the CTA pair; rank 0 issues the MMAs
    const uint32_t unit = blockIdx.x / CG, n_units = gridDim.x / CG;

    // the operand tiles need 1024-byte alignment (128B swizzle atoms); both CTAs of a pair compute the same offsets
    const uint32_t raw_addr = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
    uint8_t* aux = smem + RING_BYTES;
    uint64_t* bar_full = reinterpret_cast<uint64_t*>(aux);       // [MAX_STAGES] this CTA's TMA loads landed
    uint64_t* bar_cast = bar_full + MAX_STAGES;                   // [MAX_STAGES] (rank 0) V split into hi/lo in every CTA of the pair
    uint64_t* bar_empty = bar_cast + MAX_STAGES;                  // [MAX_STAGES] MMAs reading the stage completed
    uint64_t* bar_tfull = bar_empty + MAX_STAGES;                 // [2] accumulator complete
    uint64_t* bar_tempty = bar_tfull + 2;                         // [2] (rank 0) accumulator drained by every epilogue thread
    uint64_t* bar_full2 = bar_tempty + 2;                         // [MAX_STAGES] (rank 0, CG = 2) the loads of BOTH CTAs landed
    static_assert((4 * MAX_STAGES + 4) * 8 <= kAuxTmemSlot, "barriers overflow their shared-memory area");
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(aux + kAuxTmemSlot);
    CtaHdr* hdr = reinterpret_cast<CtaHdr*>(aux + kAuxHdr);
    uint64_t* cand_keys = reinterpret_cast<uint64_t*>(aux + kAuxCand);
    uint32_t* cand_qids = reinterpret_cast<uint32_t*>(aux + kAuxCand + (size_t)p.cap * 8);

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tm_v);
        prefetch_tmap(&tm_qh);
        prefetch_tmap(&tm_ql);
    }
    if (warp == 1 && lane == 0) {
        for (uint32_t s = 0; s < STAGES; ++s) {
            mbar_init(&bar_full[s], 1);
            mbar_init(&bar_cast[s], CG);  // one elected arrival per CTA of the pair
            mbar_init(&bar_empty[s], 1);
            mbar_init(&bar_full2[s], direct ? 1 : CG);  // direct: the leader's producer arrives once and expects both CTAs' bytes
        }
        for (uint32_t b = 0; b < 2; ++b) {
            mbar_init(&bar_tfull[b], 1);
            mbar_init(&bar_tempty[b], CG);
        }
        fence_mbar_init();
        hdr->tau = 0ull;
        hdr->count = 0;
        hdr->written = 0;
        hdr->excl = 0;
    }
    if constexpr (CG == 2) {
        __syncthreads();
        cluster_sync_all();  // the peer's barriers are initialised before anything arrives on them
    }
    if (warp == 2) tmem_alloc_cg<CG>(tmem_slot, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    if constexpr (CG == 2) cluster_sync_all();  // both halves of the pair's tensor memory are allocated
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);

    const bool raw_hi = single || (p.dbg & 16u) == 0;  // default; bit 16 of the debug word restores the in-place rounded-hi split for A/B runs
    const uint32_t n_rowtiles = (p.n_rows + G::TILE_ROWS - 1) / G::TILE_ROWS;
    const uint32_t n_tiles = n_rowtiles * p.n_qtiles;
    const uint32_t nkb = p.nkb;
    const uint32_t nst = (nkb + KPS - 1) / KPS;  // stage fills per tile

    if (warp == 0) {
        // ===== TMA producer (every CTA stages its own rows and its share of the query tile) =====
        if (lane == 0) {
            uint32_t it = 0;
            for (uint32_t t = unit; t < n_tiles; t += n_units) {
                const uint32_t rt = t / p.n_qtiles, qt = t % p.n_qtiles;
                if (!tile_live<CG>(p, rt)) continue;
                for (uint32_t ks = 0; ks < nst; ++ks, ++it) {
                    const uint32_t s = it % STAGES, ph = (it / STAGES) & 1u;
                    mbar_wait(&bar_empty[s], ph ^ 1u);
                    uint8_t* st = smem + s * STAGE_BYTES;
                    if (p.dbg & 1u) {  // timing experiment: no loads
                        if (!direct) mbar_arrive_expect_tx(&bar_full[s], 0);
                        else if (rank == 0) mbar_arrive_expect_tx(&bar_full2[s], 0);
                        continue;
                    }
                    if (single) {
                        const uint32_t nk = nkb - ks * KPS < KPS ? nkb - ks * KPS : KPS;  // k-blocks in this fill
                        if (direct) {
                            if (rank == 0) mbar_arrive_expect_tx(&bar_full2[s], 2u * nk * KB_BYTES);
                        } else {
                            mbar_arrive_expect_tx(&bar_full[s], nk * KB_BYTES);
                        }
                        for (uint32_t j = 0; j < nk; ++j) {
                            const uint32_t kb = ks * KPS + j;
                            uint8_t* blk = st + j * KB_BYTES;
                            if (direct) {
                                tma_load_2d_pair(blk, &tm_v, (int)(kb * kcols), (int)(rt * G::TILE_ROWS + rank * BM), &bar_full2[s]);
                                tma_load_2d_pair(blk + A_BYTES, &tm_qh, (int)(kb * kcols), (int)(qt * BN + rank * G::BN_LOAD), &bar_full2[s]);
                            } else {
                                tma_load_2d(blk, &tm_v, (int)(kb * kcols), (int)(rt * G::TILE_ROWS + rank * BM), &bar_full[s]);
                                tma_load_2d(blk + A_BYTES, &tm_qh, (int)(kb * kcols), (int)(qt * BN + rank * G::BN_LOAD), &bar_full[s]);
                            }
                        }
                        continue;
                    }
                    const uint32_t kb = ks;
                    mbar_arrive_expect_tx(&bar_full[s], A_BYTES + 2u * B_BYTES);
                    tma_load_2d(st, &tm_v, (int)(kb * BK), (int)(rt * G::TILE_ROWS + rank * BM), &bar_full[s]);
                    tma_load_2d(st + Q_OFF, &tm_qh, (int)(kb * BK), (int)(qt * BN + rank * G::BN_LOAD), &bar_full[s]);
                    tma_load_2d(st + 2 * A_BYTES + B_BYTES, &tm_ql, (int)(kb * BK), (int)(qt * BN + rank * G::BN_LOAD), &bar_full[s]);
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (one thread of rank 0) =====
        if (lane == 0 && rank == 0) {
            uint32_t it = 0, tn = 0;
            for (uint32_t t = unit; t < n_tiles; t += n_units) {
                const uint32_t rt = t / p.n_qtiles;
                if (!tile_live<CG>(p, rt)) continue;
                const uint32_t buf = tn & 1u, bph = (tn >> 1) & 1u;
                if constexpr (CG == 2) mbar_wait_cluster(&bar_tempty[buf], bph ^ 1u);
                else mbar_wait(&bar_tempty[buf], bph ^ 1u);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + buf * BN;
                for (uint32_t kb = 0; kb < nst; ++kb, ++it) {  // kb counts stage fills (= k-blocks unless KPS > 1)
                    const uint32_t s = it % STAGES, ph = (it / STAGES) & 1u;
                    const uint32_t sa = smem_u32(smem + s * STAGE_BYTES);
                    const uint64_t d_vh = umma_desc_sw128(sa), d_vl = umma_desc_sw128(sa + A_BYTES);
                    const uint64_t d_qh = umma_desc_sw128(sa + Q_OFF), d_ql = umma_desc_sw128(sa + 2 * A_BYTES + B_BYTES);
                    if (single) {
                        // single-pass selection: one MMA per k-step on the landed tiles (tf32: V read at 19 bits, Q rounded;
                        // bf16: both operands were rounded when their shadow arrays were written)
                        if constexpr (CG == 2) mbar_wait_cluster(&bar_full2[s], ph);
                        else mbar_wait(&bar_full[s], ph);
                        tc_fence_after();
                        const uint32_t nk = nkb - kb * KPS < KPS ? nkb - kb * KPS : KPS;
                        for (uint32_t j = 0; j < nk; ++j) {
                            const uint64_t blk = (uint64_t)((j * KB_BYTES) >> 4);  // descriptor start addresses count 16-byte units
#pragma unroll
                            for (uint32_t kk = 0; kk < ((p.dbg & 8u) ? 0u : BK / UK); ++kk) {
                                const uint64_t adv = blk + (uint64_t)((kk * UK * 4) >> 4);  // 32 bytes per k-step in both formats
                                if (half) umma_bf16_cg<CG>(d_tmem, d_vh + adv, d_qh + adv, G::IDESC_BF16, (kb | j | kk) != 0 ? 1u : 0u);
                                else umma_tf32_cg<CG>(d_tmem, d_vh + adv, d_qh + adv, G::IDESC, (kb | j | kk) != 0 ? 1u : 0u);
                            }
                        }
                    } else if (raw_hi) {
                        // the landed fp32 tile itself is the hi operand (the tensor core reads its upper 19 bits), so two
                        // thirds of the stage's MMAs start as soon as the loads land and hide the split of the lo tile
                        if constexpr (CG == 2) mbar_wait_cluster(&bar_full2[s], ph);
                        else mbar_wait(&bar_full[s], ph);
                        tc_fence_after();
#pragma unroll
                        for (uint32_t kk = 0; kk < BK / UK; ++kk) {
                            const uint64_t adv = (uint64_t)((kk * UK * 4) >> 4);  // K advance inside the swizzle atom
                            umma_tf32_cg<CG>(d_tmem, d_vh + adv, d_ql + adv, G::IDESC, (kb | kk) != 0 ? 1u : 0u);
                            umma_tf32_cg<CG>(d_tmem, d_vh + adv, d_qh + adv, G::IDESC, 1u);
                        }
                        if constexpr (CG == 2) mbar_wait_cluster(&bar_cast[s], ph);
                        else mbar_wait(&bar_cast[s], ph);
                        tc_fence_after();
#pragma unroll
                        for (uint32_t kk = 0; kk < BK / UK; ++kk) {
                            const uint64_t adv = (uint64_t)((kk * UK * 4) >> 4);
                            umma_tf32_cg<CG>(d_tmem, d_vl + adv, d_qh + adv, G::IDESC, 1u);
                        }
                    } else {
                        mbar_wait(&bar_full[s], ph);
                        if constexpr (CG == 2) mbar_wait_cluster(&bar_cast[s], ph);
                        else mbar_wait(&bar_cast[s], ph);
                        tc_fence_after();
#pragma unroll
                        for (uint32_t kk = 0; kk < ((p.dbg & 8u) ? 0u : BK / UK); ++kk) {
                            const uint64_t adv = (uint64_t)((kk * UK * 4) >> 4);  // K advance inside the swizzle atom
                            umma_tf32_cg<CG>(d_tmem, d_vl + adv, d_qh + adv, G::IDESC, (kb | kk) != 0 ? 1u : 0u);
                            umma_tf32_cg<CG>(d_tmem, d_vh + adv, d_ql + adv, G::IDESC, 1u);
                            umma_tf32_cg<CG>(d_tmem, d_vh + adv, d_qh + adv, G::IDESC, 1u);
                        }
                    }
                    umma_commit_cg<CG>(&bar_empty[s]);
                }
                umma_commit_cg<CG>(&bar_tfull[buf]);
                ++tn;
            }
        }
    } else if (warp == 3) {
        // ===== relay (CTA pairs with raw hi operands): tells rank 0 that this CTA's loads of a stage have landed =====
        if (CG == 2 && raw_hi && !direct && lane == 0) {
            uint32_t it = 0;
            for (uint32_t t = unit; t < n_tiles; t += n_units) {
                const uint32_t rt = t / p.n_qtiles;
                if (!tile_live<CG>(p, rt)) continue;
                for (uint32_t kb = 0; kb < nst; ++kb, ++it) {
                    const uint32_t s = it % STAGES, ph = (it / STAGES) & 1u;
                    mbar_wait(&bar_full[s], ph);
                    if (rank != 0) mbar_arrive_remote(&bar_full2[s], 0);
                    else mbar_arrive(&bar_full2[s]);
                }
            }
        }
    } else if (warp >= 4 && warp < 8) {
        // ===== split warps: V tile -> hi (in place) and lo =====
        const int tt = tid - 128;
        uint32_t it = 0;
        for (uint32_t t = unit; t < n_tiles && !single; t += n_units) {  // single-pass selection needs no lo tile
            const uint32_t rt = t / p.n_qtiles;
            if (!tile_live<CG>(p, rt)) continue;
            for (uint32_t kb = 0; kb < nkb; ++kb, ++it) {
                const uint32_t s = it % STAGES, ph = (it / STAGES) & 1u;
                mbar_wait(&bar_full[s], ph);
                float4* hi = reinterpret_cast<float4*>(smem + s * STAGE_BYTES);
                float4* lo = reinterpret_cast<float4*>(smem + s * STAGE_BYTES + A_BYTES);
                if (raw_hi) {
                    // lo = x - (what the tensor core will read of x: its upper 19 bits), rounded to tf32; x stays as it is
#pragma unroll
                    for (uint32_t i = 0; i < A_BYTES / 16 / 128; ++i) {
                        const uint32_t idx = tt + i * 128;
                        const float4 x = hi[idx];
                        float4 l;
                        l.x = rna_tf32(x.x - __uint_as_float(__float_as_uint(x.x) & 0xFFFFE000u));
                        l.y = rna_tf32(x.y - __uint_as_float(__float_as_uint(x.y) & 0xFFFFE000u));
                        l.z = rna_tf32(x.z - __uint_as_float(__float_as_uint(x.z) & 0xFFFFE000u));
                        l.w = rna_tf32(x.w - __uint_as_float(__float_as_uint(x.w) & 0xFFFFE000u));
                        lo[idx] = l;
                    }
                } else
#pragma unroll
                for (uint32_t i = 0; i < ((p.dbg & 2u) ? 0u : A_BYTES / 16 / 128); ++i) {
                    const uint32_t idx = tt + i * 128;
                    const float4 x = hi[idx];
                    float4 h, l;
                    h.x = rna_tf32(x.x);
                    h.y = rna_tf32(x.y);
                    h.z = rna_tf32(x.z);
                    h.w = rna_tf32(x.w);
                    l.x = rna_tf32(x.x - h.x);
                    l.y = rna_tf32(x.y - h.y);
                    l.z = rna_tf32(x.z - h.z);
                    l.w = rna_tf32(x.w - h.w);
                    hi[idx] = h;
                    lo[idx] = l;
                }
                fence_proxy_async();  // generic-proxy writes -> visible to the tensor core's async-proxy reads
                named_bar_sync(2, 128);  // the four split warps; then ONE thread signals (a cluster-scope release per
                                         // thread costs a gpu-wide fence each: profiles/r1_batch_c2_hot_sass.txt)
                if (tt == 0) {
                    if (CG == 2 && rank != 0) mbar_arrive_remote(&bar_cast[s], 0);
                    else mbar_arrive(&bar_cast[s]);
                }
            }
        }
    } else if (warp >= 8) {
        // ===== epilogue: TMEM -> metric -> loosened filter -> per-CTA top-k of approximate scores =====
        // Four or EIGHT warps (blockDim): a warp may only read the TMEM lanes of its quarter (warp % 4), so with eight warps the
        // two warps of a quarter take the even / odd 32-column chunks.  One epilogue warp per scheduler runs its dependent
        // chains at ~0.25 instructions per cycle (ncu: 330 instructions and ~1400 cycles per chunk, profiles/r2b_batch_c2_bf16_*):
        // once the bf16 rung halved the MMA time the epilogue became the bound, and a second warp per scheduler hides that latency.
        const uint32_t quarter = (uint32_t)warp & 3u;  // TMEM lanes [32*quarter, 32*quarter+32)
        const uint32_t epi_threads = blockDim.x - BASE_WARPS * 32u;       // 128 or 256
        const uint32_t cstep = epi_threads >> 7;                          // column chunks are dealt round-robin to 1 or 2 warps per quarter
        const uint32_t cfirst = ((uint32_t)warp - BASE_WARPS) >> 2;       // 0 or 1
        const bool take_max = p.take_max != 0;
        unsigned long long scored = 0;
        uint32_t tn = 0;
        bool nonfinite = false;
        const float sgn = take_max ? 1.0f : -1.0f;
        const float delta = p.has_filter ? __ldg(p.delta) : 0.f;
        float xbest = -INFINITY;  // best signed score among the pairs this thread excluded by the top-k test
        for (uint32_t t = unit; t < n_tiles; t += n_units) {
            const uint32_t rt = t / p.n_qtiles, qt = t % p.n_qtiles;
            if (!tile_live<CG>(p, rt)) continue;
            const uint32_t buf = tn & 1u, bph = (tn >> 1) & 1u;
            const uint32_t row = rt * G::TILE_ROWS + rank * BM + quarter * 32 + lane;
            bool live = row < p.n_rows;
            if (live && p.row_mask) {
                const uint32_t w = (row >> 5) < p.row_mask_words ? __ldg(p.row_mask + (row >> 5)) : 0xFFFFFFFFu;
                live = (w >> (row & 31)) & 1u;
            }
            float rs = 0.f;  // cosine: 1/|v|; euclidean: |v|^2
            if (METRIC != OTTERS_METRIC_DOT && row < p.n_rows) {
                const float inv = __ldg(p.inv_norms + row);
                rs = METRIC == OTTERS_METRIC_COSINE ? inv : (inv > 0.f ? 1.0f / (inv * inv) : 0.f);
            }
            const uint32_t q_base = qt * BN;
            const uint32_t nq_tile = p.nq - q_base < BN ? p.nq - q_base : BN;
            if (live && cfirst == 0) scored += nq_tile;  // (one warp of the quarter accounts the row)
            mbar_wait(&bar_tfull[buf], bph);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((quarter * 32u) << 16) + buf * BN;
            for (uint32_t c = cfirst; c < BN / 32; c += cstep) {
                if (c * 32 >= nq_tile || (p.dbg & 4u)) break;  // warp-uniform
                uint32_t v[32];
                tmem_ld_32x32b_x32(taddr + c * 32, v);
                tmem_ld_wait();
                // threshold of this chunk as a float pre-test (the exact key test is inside the push)
                const uint64_t tau = ld_volatile_u64(&hdr->tau);
                const float tau_s = tau ? key_score(tau, take_max) : (take_max ? -INFINITY : INFINITY);
                if (!p.has_filter && c * 32 + 32 <= nq_tile && !(p.dbg & 32u)) {  // warp-uniform
                    // common case (unfiltered, full chunk): two or three instructions per pair.  In the signed domain
                    // (larger = better) only the chunk maximum matters: below the threshold nothing can enter the list and
                    // the maximum is the best excluded score.  z turns NaN iff some score is inf / NaN.
                    const float tau_t = tau ? tau_s * sgn : -INFINITY;
                    float cmax = -INFINITY, zz[4] = {0.f, 0.f, 0.f, 0.f};  // four short chains instead of one long one
#pragma unroll
                    for (uint32_t j = 0; j < 32; ++j) {
                        const float tj = batch_score<METRIC>(__uint_as_float(v[j]), p.q_scal, q_base + c * 32 + j, rs) * sgn;
                        cmax = fmaxf(cmax, tj);
                        zz[j & 3] = fmaf(tj, 0.f, zz[j & 3]);
                    }
                    const float z = (zz[0] + zz[1]) + (zz[2] + zz[3]);
                    const bool hit = live && (!(z == 0.f) || cmax >= tau_t);
                    if (!__any_sync(FULL, hit)) {
                        if (live) xbest = fmaxf(xbest, cmax);
                        continue;
                    }
                    // some lane has a candidate (or an odd value).  redo_general: take the general path below — straight-line
                    // scoring of the 32 columns already in registers, then only the columns where some lane passed are
                    // re-read from TMEM; otherwise column by column, every column re-read from TMEM (32 dependent tcgen05.ld
                    // round trips: ~3 µs per chunk, which made tiles with several such chunks outlast their MMAs)
                    if (p.redo_general) goto general_path;
#pragma unroll 1
                    for (uint32_t j = 0; j < 32; ++j) {
                        const uint32_t qi = q_base + c * 32 + j;
                        const float sj = batch_score<METRIC>(__uint_as_float(tmem_ld_32x32b_x1(taddr + c * 32 + j)), p.q_scal, qi, rs);
                        if (live && !(fabsf(sj) <= FLT_MAX)) nonfinite = true;
                        const bool top = take_max ? sj >= tau_s : sj <= tau_s;
                        if (live && !top) xbest = fmaxf(xbest, sj * sgn);
                        const bool psh = live && top;
                        if (__any_sync(FULL, psh))
                            warp_push_pairs(hdr, cand_keys, cand_qids, p.cap, p.k, psh, make_key(sj, row, take_max), qi, p.g_tau, lane);
                    }
                    continue;
                }
                // general path (vec_filter, ragged last chunk): straight-line scoring of the 32 columns, one bit per passing column
            general_path:
                uint32_t pass = 0;
#pragma unroll
                for (uint32_t j = 0; j < 32; ++j) {
                    const bool inq = c * 32 + j < nq_tile;  // columns past the batch hold zero queries: never read their scalars
                    const float s = batch_score<METRIC>(__uint_as_float(v[j]), p.q_scal, inq ? q_base + c * 32 + j : 0u, rs);
                    bool ok = live && inq;
                    if (ok && !(fabsf(s) <= FLT_MAX)) nonfinite = true;  // inf / NaN: let the exact path decide
                    if (p.has_filter) ok = ok && score_passes_loose(s, p.thr, p.cmp, delta);
                    const bool top = take_max ? s >= tau_s : s <= tau_s;
                    if (ok && !top) xbest = fmaxf(xbest, s * sgn);
                    pass |= (ok && top) ? (1u << j) : 0u;
                }
                // rare path: columns where some lane passed are re-read one at a time and pushed (kept out of the
                // unrolled loop so that the hot code stays small)
                uint32_t cols = __reduce_or_sync(FULL, pass);
                while (cols) {
                    const uint32_t j = __ffs(cols) - 1;
                    cols &= cols - 1;
                    const uint32_t qi = q_base + c * 32 + j;
                    const float s = batch_score<METRIC>(__uint_as_float(tmem_ld_32x32b_x1(taddr + c * 32 + j)), p.q_scal, qi, rs);
                    warp_push_pairs(hdr, cand_keys, cand_qids, p.cap, p.k, (pass >> j) & 1u, make_key(s, row, take_max), qi, p.g_tau, lane);
                }
            }
            tc_fence_before();
            named_bar_sync(3, epi_threads);  // every epilogue thread has drained its TMEM lanes; one thread signals
            if (warp == 8 && lane == 0) {
                if (CG == 2 && rank != 0) mbar_arrive_remote(&bar_tempty[buf], 0);
                else mbar_arrive(&bar_tempty[buf]);
            }
            ++tn;
            // adopt the grid-wide threshold once per tile
            if (warp == 8 && lane == 0) {
                const unsigned long long g = *reinterpret_cast<volatile unsigned long long*>(p.g_tau);
                if (g > hdr->tau) atomicMax(&hdr->tau, g);
            }
        }
        if (__any_sync(FULL, nonfinite) && lane == 0) atomicOr(p.g_flags, 1u);
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) xbest = fmaxf(xbest, __shfl_xor_sync(FULL, xbest, d));
        if (lane == 0 && xbest > -INFINITY) atomicMax(&hdr->excl, (uint32_t)(make_key(xbest * sgn, 0, take_max) >> 32));
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) scored += __shfl_xor_sync(FULL, scored, d);
        if (lane == 0 && scored) atomicAdd(p.pairs_scored, scored);
        named_bar_sync(1, epi_threads);  // the epilogue warps: every push has completed
        if (warp == 8) {
            const uint32_t cnt = hdr->count;
            warp_sort_pairs(cand_keys, cand_qids, cnt, p.cap, lane);
            const uint32_t n = cnt < p.k ? cnt : p.k;
            for (uint32_t i = lane; i < n; i += 32) {
                p.cta_keys[(size_t)blockIdx.x * p.k + i] = cand_keys[i];
                p.cta_qids[(size_t)blockIdx.x * p.k + i] = cand_qids[i];
            }
            if (lane == 0) {
                p.cta_counts[blockIdx.x] = n;
                // best score this CTA left out: rejected by the top-k test, dropped by a compaction, or cut here
                uint32_t x = hdr->excl;
                if (cnt > p.k && (uint32_t)(cand_keys[p.k] >> 32) > x) x = (uint32_t)(cand_keys[p.k] >> 32);
                if (x) atomicMax(p.g_excl, x);
            }
        }
    }

    __syncwarp();
    tc_fence_before();
    __syncthreads();
    if constexpr (CG == 2) cluster_sync_all();  // no arrival on the peer's barriers is still in flight; both CTAs are done with TMEM
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc_cg<CG>(tmem_base, TMEM_COLS);
    }
}

// ---- query split: hi = rna_tf32(q), lo = rna_tf32(q - hi); rows >= nq are zero -----------------------------
// One warp per (padded) query row.  Also leaves |q|^2 per query (the euclidean epilogue's scalar) and the
// largest |q|^2 of the batch (for the error bound).
__global__ void split_queries_kernel(const float* q, uint32_t nq, uint32_t nq_pad, uint32_t dim_pad, float* qh, float* ql,
                                     float* qn2, uint32_t* qmax2_bits) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t wpb = blockDim.x >> 5;
    for (uint32_t r = blockIdx.x * wpb + (threadIdx.x >> 5); r < nq_pad; r += gridDim.x * wpb) {
        float acc = 0.f;
        for (uint32_t c = lane; c < dim_pad; c += 32) {
            const size_t i = (size_t)r * dim_pad + c;
            const float x = r < nq ? q[i] : 0.f;
            const float h = rna_tf32(x);
            qh[i] = h;
            ql[i] = rna_tf32(x - h);
            acc = fmaf(x, x, acc);
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(FULL, acc, d);
        if (lane == 0 && r < nq) {
            qn2[r] = acc;
            if (acc > 0.f && acc <= FLT_MAX) atomicMax(qmax2_bits, __float_as_uint(acc));  // positive floats order like uints
            else if (!(acc <= FLT_MAX)) atomicMax(qmax2_bits, 0x7F800000u);                   // inf / NaN query
        }
    }
}

// Bound on |tensor-core score - exact score| for this batch: kappa * |q|max * |v|max for dot
// products, kappa for cosine, kappa * (|q|max + |v|max)^2 for squared distances; kappa depends on the arithmetic rung (below).
__global__ void batch_delta_kernel(int metric, uint32_t dim, uint32_t passes, const uint32_t* qmax2_bits, const uint32_t* vmin_inv_bits,
                                   float* delta) {
    // Bounds relative to |q||v| (tests/test_tf32_bound.py re-derives them on the host).
    // 3xTF32: the dropped Vlo.Qlo term and the lo roundings stay below 2^-21; every MMA step adds into the fp32
    //   accumulator with at most one truncated ulp (2^-23 of a partial sum that never exceeds |q||v|), 3*dim/8 steps:
    //   3*(dim/8+1)*2^-23 + 2^-21 <= 2^-15 * max(1, dim/640).
    // single pass: V is read at 19 bits (truncation, 2^-10), Q is rounded to tf32 (2^-11): 0.75 * 2^-9 when every error
    //   lines up, plus (dim/8+1)*2^-23 of accumulation <= 2^-9 * max(1, dim/1024).
    // Safety factor: the accumulator model above (one truncated ulp per MMA step) is a datasheet-level assumption about the
    // tensor core's internal alignment, not something the runtime check can see for the EXCLUDED pairs.  The measured error
    // is 20-90x below the bound, so doubling it (single pass) / quadrupling it (3xTF32, whose margin was only ~5 %) costs
    // almost no extra fallbacks.
    // bf16 rung: both operands were rounded to nearest-even at 8 significant bits by convert_bf16_kernel — an error this library
    //   produces itself, (1 + 2^-8)^2 - 1 = 2^-7 + 2^-16 relative to sum |q_i v_i| <= |q||v| — and the products are exact in
    //   the fp32 accumulator up to (dim/16 + 1) truncating additions: 1.01 * 2^-7 + 4 * (dim/16 + 1) * 2^-23 (safety factor 4
    //   on the accumulator model only; the operand term is exact arithmetic, re-derived by tests/test_tf32_bound.py).
    const double kappa = passes == 2   ? 1.01 * ldexp(1.0, -7) + ((double)dim / 16.0 + 1.0) * ldexp(1.0, -21)
                         : passes == 1 ? 2.0 * ldexp(1.0, -9) * fmax(1.0, (double)dim / 1024.0)
                                       : 4.0 * ldexp(1.0, -15) * fmax(1.0, (double)dim / 640.0);
    double d;
    if (metric == OTTERS_METRIC_COSINE) {
        d = kappa;
    } else {
        const double qmax = sqrt((double)__uint_as_float(*qmax2_bits));
        const float vinv = __uint_as_float(*vmin_inv_bits);  // smallest positive inverse row norm (huge when all rows are zero)
        const double vmax = (vinv > 0.f && vinv < 1e30f) ? 1.0 / (double)vinv : 0.0;
        d = metric == OTTERS_METRIC_DOT ? kappa * qmax * vmax : kappa * (qmax + vmax) * (qmax + vmax);
    }
    *delta = (float)d;  // inf / NaN propagate: the host then takes the exact path
}

// ---- exact re-scoring of the selected (row, query) pairs ---------------------------------------------------
// Two threads per pair hold the eight f32x8 lane accumulators (4 each) exactly like K1 (scan_kernel.cuh): multiply and
// add are separate round-to-nearest operations, 8-column blocks in order, wide's non-AVX reduce_add order,
// serial dim%8 tail added last, cosine as (dot * q_inv) * row_inv (reference src/vec_compute.rs:9-54).
// HALF: the store's rows are bf16 (OTTERS_VECTORS_FMT_BF16): widened exactly, same arithmetic (as K1's load_row4<HALF>).
template <int METRIC, bool HALF>
__global__ void __launch_bounds__(256) rescore_kernel(const __grid_constant__ RescoreParams p) {
    const uint32_t slot = blockIdx.x * (blockDim.x >> 1) + (threadIdx.x >> 1);
    const int h = threadIdx.x & 1;
    const int lane = threadIdx.x & 31;
    const uint32_t total = p.n_lists * p.k;
    const uint32_t list = slot < total ? slot / p.k : 0, pos = slot < total ? slot % p.k : 0;
    const bool valid = slot < total && pos < p.cta_counts[list];
    uint64_t key_a = 0;
    uint32_t qid = 0, row = 0;
    if (valid) {
        key_a = p.cta_keys[(size_t)list * p.k + pos];
        qid = p.cta_qids[(size_t)list * p.k + pos];
        row = key_row(key_a);
    }
    const bool take_max = p.take_max != 0;
    const uint32_t dim8 = p.dim & ~7u, ntail = p.dim & 7u;
    const float* vrow = p.vectors + (size_t)row * p.pitch_g;                                      // fp32 rows
    const uint16_t* vrow_h = reinterpret_cast<const uint16_t*>(p.vectors) + (size_t)row * p.pitch_g;  // bf16 rows
    const float* qrow = p.queries + (size_t)qid * p.dim_pad;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    if (valid) {
        const float4* vp = reinterpret_cast<const float4*>(vrow) + h;
        const uint2* vph = reinterpret_cast<const uint2*>(vrow_h) + h;
        const float4* qp = reinterpret_cast<const float4*>(qrow) + h;
        const uint32_t nblk = dim8 >> 3;
#pragma unroll 4
        for (uint32_t j = 0; j < nblk; ++j) {
            float4 v;
            if constexpr (HALF) {
                const uint2 w = __ldg(vph + 2 * j);
                v = make_float4(__uint_as_float(w.x << 16), __uint_as_float(w.x & 0xFFFF0000u), __uint_as_float(w.y << 16),
                                __uint_as_float(w.y & 0xFFFF0000u));
            } else {
                v = __ldg(vp + 2 * j);
            }
            const float4 q = __ldg(qp + 2 * j);
            if (METRIC == OTTERS_METRIC_EUCLIDEAN) {
                const float d0 = __fsub_rn(q.x, v.x), d1 = __fsub_rn(q.y, v.y), d2 = __fsub_rn(q.z, v.z), d3 = __fsub_rn(q.w, v.w);
                a0 = __fadd_rn(a0, __fmul_rn(d0, d0));
                a1 = __fadd_rn(a1, __fmul_rn(d1, d1));
                a2 = __fadd_rn(a2, __fmul_rn(d2, d2));
                a3 = __fadd_rn(a3, __fmul_rn(d3, d3));
            } else {
                a0 = __fadd_rn(a0, __fmul_rn(q.x, v.x));
                a1 = __fadd_rn(a1, __fmul_rn(q.y, v.y));
                a2 = __fadd_rn(a2, __fmul_rn(q.z, v.z));
                a3 = __fadd_rn(a3, __fmul_rn(q.w, v.w));
            }
        }
    }
    const float sdot = __fadd_rn(__fadd_rn(__fadd_rn(a0, a1), a2), a3);
    const float other = __shfl_xor_sync(FULL, sdot, 1);
    const float tot = h == 0 ? __fadd_rn(sdot, other) : __fadd_rn(other, sdot);
    float tail = -0.0f;
    if (valid && ntail) {
        for (uint32_t e = 0; e < ntail; ++e) {
            const float ve = HALF ? __uint_as_float((uint32_t)vrow_h[dim8 + e] << 16) : vrow[dim8 + e];
            if (METRIC == OTTERS_METRIC_EUCLIDEAN) {
                const float d = __fsub_rn(qrow[dim8 + e], ve);
                tail = __fadd_rn(tail, __fmul_rn(d, d));
            } else {
                tail = __fadd_rn(tail, __fmul_rn(qrow[dim8 + e], ve));
            }
        }
    }
    float score = __fadd_rn(tot, tail);
    if (METRIC == OTTERS_METRIC_COSINE && valid) score = __fmul_rn(__fmul_rn(score, p.q_inv[qid]), p.inv_norms[row]);
    bool ok = valid && h == 0 && !(score != score);  // NaN never returned (src/vec_compute.rs:237-239)
    if (p.has_filter) ok = ok && score_passes(score, p.thr, p.cmp);
    if (h == 0 && slot < p.out_slots) {
        Cand c;
        c.key = ok ? make_key(score, row, take_max) : 0ull;
        c.qid = ok ? qid : 0xFFFFFFFFu;
        c.pad = 0;
        p.out[slot] = c;
    }
    // telemetry: largest |approximate - exact| over the re-scored pairs (non-negative floats order like uints)
    float err = 0.f;
    if (valid && h == 0 && !(score != score)) {
        const float e = fabsf(key_score(key_a, take_max) - score);
        if (e <= FLT_MAX) err = e;
    }
    const unsigned m = __ballot_sync(FULL, ok);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) err = fmaxf(err, __shfl_xor_sync(FULL, err, d));
    if (lane == 0) {
        if (m) atomicAdd(p.out_count, (uint32_t)__popc(m));
        if (err > 0.f) atomicMax(p.max_err_bits, __float_as_uint(err));
    }
}

__global__ void pad_cands_kernel(Cand* buf, uint32_t from, uint32_t to) {
    for (uint32_t i = from + blockIdx.x * blockDim.x + threadIdx.x; i < to; i += gridDim.x * blockDim.x) {
        Cand c;
        c.key = 0ull;
        c.qid = 0xFFFFFFFFu;
        c.pad = 0;
        buf[i] = c;
    }
}

// largest row norm of the store: 1 / min positive inverse norm (0 when every row is zero)
__global__ void min_inv_norm_kernel(const float* inv, uint64_t n, uint32_t* out_bits) {
    float m = INFINITY;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const float v = inv[i];
        if (v > 0.f && v < m) m = v;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) m = fminf(m, __shfl_xor_sync(FULL, m, d));
    if ((threadIdx.x & 31) == 0) atomicMin(out_bits, __float_as_uint(m));  // positive floats order like uints
}

// ---- bf16 shadow arrays for the bf16 rung: dst[r][c] = bf16_rn(src[r][c]) for r < n_src, c < dim; everything else 0 ------
// One thread converts 8 columns (two 16-byte loads, one 16-byte store); pitches are multiples of 4 floats / 8 bf16.
__global__ void convert_bf16_kernel(const float* __restrict__ src, uint64_t src_pitch, uint64_t n_src, uint32_t dim,
                                    uint16_t* __restrict__ dst, uint64_t dst_pitch, uint64_t n_dst) {
    const uint64_t groups = dst_pitch >> 3;
    const uint64_t total = n_dst * groups;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t r = i / groups;
        const uint32_t c0 = (uint32_t)(i - r * groups) << 3;
        float x[8];
#pragma unroll
        for (uint32_t j = 0; j < 8; ++j) x[j] = 0.f;
        if (r < n_src) {
            const float* sp = src + r * src_pitch + c0;
            if (c0 + 8 <= dim) {
                const float4 a = __ldg(reinterpret_cast<const float4*>(sp)), b = __ldg(reinterpret_cast<const float4*>(sp) + 1);
                x[0] = a.x, x[1] = a.y, x[2] = a.z, x[3] = a.w, x[4] = b.x, x[5] = b.y, x[6] = b.z, x[7] = b.w;
            } else {
#pragma unroll
                for (uint32_t j = 0; j < 8; ++j)
                    if (c0 + j < dim) x[j] = __ldg(sp + j);
            }
        }
        uint32_t w[4];
#pragma unroll
        for (uint32_t j = 0; j < 4; ++j) {
            // round to nearest even; NaN stays NaN and |x| >= 2^128 (1 - 2^-9) becomes inf (the kernel then flags the batch)
            // (the first source operand lands in the upper half of the pair)
            asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(w[j]) : "f"(x[2 * j + 1]), "f"(x[2 * j]));
        }
        *reinterpret_cast<uint4*>(dst + r * dst_pitch + c0) = make_uint4(w[0], w[1], w[2], w[3]);
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn tensor_map_encoder() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) != cudaSuccess ||
            qres != cudaDriverEntryPointSuccess)
            sym = nullptr;
        fn = reinterpret_cast<EncodeTiledFn>(sym);
    }
    return fn;
}

// 2D fp32 tensor [rows][cols] with a row pitch of pitch_floats, boxes of box_rows x 32 columns, 128B swizzle;
// out-of-bounds elements are filled with zeros
// (half: a bf16 tensor with a row pitch of pitch_elems bf16 and boxes of box_rows x 64 columns — the same 128 bytes per box row)
int make_tensor_map(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t pitch_elems, uint32_t box_rows, bool half = false) {
    EncodeTiledFn enc = tensor_map_encoder();
    if (!enc) return fail(OTTERS_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {pitch_elems * (half ? 2 : sizeof(float))};
    cuuint32_t box[2] = {half ? 2 * BK : BK, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(out, half ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, kSwizzleBytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(OTTERS_ERR_CUDA, "cuTensorMapEncodeTiled failed with code " + std::to_string((int)r));
    return OTTERS_OK;
}

template <int METRIC, int CG>
int launch_batch_one(const CUtensorMap& tv, const CUtensorMap& tqh, const CUtensorMap& tql, const BatchParams& p, uint32_t grid,
                     uint32_t smem, uint32_t* configured, cudaStream_t s) {
    auto kern = batch_kernel<METRIC, CG>;
    static uint32_t limits[64];
    uint32_t& have = smem_limit_slot(limits);
    (void)configured;
    if (smem > have) {
        OTTERS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        have = smem;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3((BASE_WARPS + (p.epi_warps == 4u ? 4u : 8u)) * 32u);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CG;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    OTTERS_CUDA(cudaLaunchKernelEx(&cfg, kern, tv, tqh, tql, p));
    return OTTERS_OK;
}

template <int CG>
int launch_batch_cg(const CUtensorMap& tv, const CUtensorMap& tqh, const CUtensorMap& tql, const BatchParams& p, uint32_t grid,
                    uint32_t smem, int metric, uint32_t* configured, cudaStream_t s) {
    switch (metric) {
    case OTTERS_METRIC_COSINE: return launch_batch_one<OTTERS_METRIC_COSINE, CG>(tv, tqh, tql, p, grid, smem, configured, s);
    case OTTERS_METRIC_EUCLIDEAN: return launch_batch_one<OTTERS_METRIC_EUCLIDEAN, CG>(tv, tqh, tql, p, grid, smem, configured, s);
    case OTTERS_METRIC_DOT: return launch_batch_one<OTTERS_METRIC_DOT, CG>(tv, tqh, tql, p, grid, smem, configured, s);
    }
    return fail(OTTERS_ERR_INVALID, "Search metric is not set");
}

}  // namespace

uint32_t batch_smem_bytes(uint32_t cap) {
    static_assert(Geo<1>::STAGES * Geo<1>::STAGE_BYTES == Geo<2>::STAGES * Geo<2>::STAGE_BYTES, "both geometries stage 192 KB");
    return Geo<1>::STAGES * Geo<1>::STAGE_BYTES + 1024 + kAuxCand + cap * 12;
}

int launch_split_queries(const float* q, uint32_t nq, uint32_t nq_pad, uint32_t dim_pad, float* qh, float* ql, float* qn2,
                         uint32_t* qmax2_bits, cudaStream_t s) {
    const unsigned blocks = (unsigned)std::min<uint32_t>((nq_pad + 7) / 8, 1184);
    split_queries_kernel<<<blocks ? blocks : 1, 256, 0, s>>>(q, nq, nq_pad, dim_pad, qh, ql, qn2, qmax2_bits);
    OTTERS_CUDA(cudaGetLastError());
    return OTTERS_OK;
}

int launch_batch_delta(int metric, uint32_t dim, uint32_t passes, const uint32_t* qmax2_bits, const uint32_t* vmin_inv_bits, float* delta,
                       cudaStream_t s) {
    batch_delta_kernel<<<1, 1, 0, s>>>(metric, dim, passes, qmax2_bits, vmin_inv_bits, delta);
    OTTERS_CUDA(cudaGetLastError());
    return OTTERS_OK;
}

int launch_batch(const BatchLaunch& l, BatchParams p, int metric, uint32_t* smem_configured, cudaStream_t s) {
    const uint32_t cg = l.cta_group == 2 ? 2 : 1;
    CUtensorMap tv, tqh, tql;
    int rc;
    if (p.passes == 2) {
        if (!l.v_half || !l.q_half) return fail(OTTERS_ERR_INVALID, "the bf16 rung needs the bf16 shadow arrays");
        rc = make_tensor_map(&tv, l.v_half, l.n_rows, l.dim, l.pitch_h, BM, true);
        if (rc) return rc;
        rc = make_tensor_map(&tqh, l.q_half, l.nq_pad, l.dim, l.q_pitch_h, BN / cg, true);
        if (rc) return rc;
        tql = tqh;  // unused by single-pass rungs
    } else {
        rc = make_tensor_map(&tv, l.vectors, l.n_rows, l.dim, l.pitch_g, BM);
        if (rc) return rc;
        rc = make_tensor_map(&tqh, l.q_hi, l.nq_pad, l.dim, l.dim_pad, BN / cg);
        if (rc) return rc;
        rc = make_tensor_map(&tql, l.q_lo, l.nq_pad, l.dim, l.dim_pad, BN / cg);
        if (rc) return rc;
    }
    p.n_qtiles = l.nq_pad / BN;
    const uint32_t kcols = p.passes == 2 ? 2 * BK : BK;
    p.nkb = (l.dim + kcols - 1) / kcols;
    const uint32_t smem = batch_smem_bytes(p.cap);
    if (cg == 2) return launch_batch_cg<2>(tv, tqh, tql, p, l.grid, smem, metric, smem_configured, s);
    return launch_batch_cg<1>(tv, tqh, tql, p, l.grid, smem, metric, smem_configured, s);
}

int launch_rescore(const RescoreParams& p, int metric, uint32_t n_sort, cudaStream_t s) {
    const uint32_t total = p.n_lists * p.k;
    const unsigned blocks = (total + 127) / 128;
    if (blocks && p.half) {
        switch (metric) {
        case OTTERS_METRIC_COSINE: rescore_kernel<OTTERS_METRIC_COSINE, true><<<blocks, 256, 0, s>>>(p); break;
        case OTTERS_METRIC_EUCLIDEAN: rescore_kernel<OTTERS_METRIC_EUCLIDEAN, true><<<blocks, 256, 0, s>>>(p); break;
        default: rescore_kernel<OTTERS_METRIC_DOT, true><<<blocks, 256, 0, s>>>(p); break;
        }
    } else if (blocks) {
        switch (metric) {
        case OTTERS_METRIC_COSINE: rescore_kernel<OTTERS_METRIC_COSINE, false><<<blocks, 256, 0, s>>>(p); break;
        case OTTERS_METRIC_EUCLIDEAN: rescore_kernel<OTTERS_METRIC_EUCLIDEAN, false><<<blocks, 256, 0, s>>>(p); break;
        default: rescore_kernel<OTTERS_METRIC_DOT, false><<<blocks, 256, 0, s>>>(p); break;
        }
    }
    if (n_sort > total) pad_cands_kernel<<<64, 256, 0, s>>>(p.out, total, n_sort);
    OTTERS_CUDA(cudaGetLastError());
    return OTTERS_OK;
}

int launch_convert_bf16(const float* src, uint64_t src_pitch, uint64_t n_src, uint32_t dim, uint16_t* dst, uint64_t dst_pitch, uint64_t n_dst,
                        cudaStream_t s) {
    if (dst_pitch & 7u) return fail(OTTERS_ERR_INVALID, "bf16 row pitch must be a multiple of 8");
    const uint64_t total = n_dst * (dst_pitch >> 3);
    if (!total) return OTTERS_OK;
    const unsigned blocks = (unsigned)std::min<uint64_t>((total + 255) / 256, 148 * 16);
    convert_bf16_kernel<<<blocks, 256, 0, s>>>(src, src_pitch, n_src, dim, dst, dst_pitch, n_dst);
    OTTERS_CUDA(cudaGetLastError());
    return OTTERS_OK;
}

int launch_min_inv_norm(const float* inv, uint64_t n, uint32_t* out_bits, cudaStream_t s) {
    OTTERS_CUDA(cudaMemsetAsync(out_bits, 0x7F, 4, s));  // 0x7F7F7F7F: a huge positive float
    min_inv_norm_kernel<<<296, 256, 0, s>>>(inv, n, out_bits);
    OTTERS_CUDA(cudaGetLastError());
    return OTTERS_OK;
}

}  // namespace otters
