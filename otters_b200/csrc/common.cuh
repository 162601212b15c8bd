// common.cuh — shared device helpers for libotters_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#ifndef __CUDA_ARCH__
#define OTTERS_HOST_ONLY 1
#endif

namespace otters {

// ---------------------------------------------------------------------------------------------
// Candidate keys.  A candidate (score, row) is packed into one 64-bit key so that a LARGER key is
// always the BETTER candidate:
//   take Max: key = ord(score) << 32 | ~row        (higher score first, then lower row)
//   take Min: key = ~ord(score) << 32 | ~row       (lower score first, then lower row)
// ord() is the usual monotone float->uint map.  Scores are canonicalised with +0.0f first so that
// -0.0 and +0.0 tie (the reference admits with IEEE compares, src/vec_compute.rs:243-246).
// key == 0 never encodes a real candidate (NaN scores are dropped, src/vec_compute.rs:237-239) and
// is used as the empty/"no threshold" value.
// ---------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint32_t ord_f32(float f) {
#ifdef __CUDA_ARCH__
    uint32_t b = __float_as_uint(f);
#else
    union { float f; uint32_t u; } c; c.f = f; uint32_t b = c.u;
#endif
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__host__ __device__ __forceinline__ float unord_f32(uint32_t o) {
    uint32_t b = (o & 0x80000000u) ? (o & 0x7FFFFFFFu) : ~o;
#ifdef __CUDA_ARCH__
    return __uint_as_float(b);
#else
    union { float f; uint32_t u; } c; c.u = b; return c.f;
#endif
}
__host__ __device__ __forceinline__ uint64_t make_key(float score, uint32_t row, bool take_max) {
    uint32_t o = ord_f32(score + 0.0f);
    if (!take_max) o = ~o;
    return ((uint64_t)o << 32) | (uint64_t)(~row);
}
__host__ __device__ __forceinline__ float key_score(uint64_t key, bool take_max) {
    uint32_t o = (uint32_t)(key >> 32);
    if (!take_max) o = ~o;
    return unord_f32(o);
}
__host__ __device__ __forceinline__ uint32_t key_row(uint64_t key) { return ~(uint32_t)key; }

// Sortable candidate: order = key descending, then qid ascending.
struct __align__(16) Cand {
    uint64_t key;
    uint32_t qid;
    uint32_t pad;
};
__host__ __device__ __forceinline__ bool cand_before(const Cand& a, const Cand& b) {
    return a.key > b.key || (a.key == b.key && a.qid < b.qid);
}

// IEEE compare of a score against the vec_filter threshold (src/vec_compute.rs:56-74, :216-228)
__host__ __device__ __forceinline__ bool score_passes(float s, float thr, int cmp) {
    switch (cmp) {
    case 0: return s < thr;
    case 1: return s > thr;
    case 2: return s <= thr;
    case 3: return s >= thr;
    default: return s == thr;
    }
}

#ifdef __CUDACC__
// ---------------------------------------------------------------------------------------------
// PTX wrappers: mbarrier + TMA bulk copy (cp.async.bulk, SASS UBLKCP)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_inval(uint64_t* bar) {
    asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Spins until the phase with the given parity completes.  A watchdog turns a lost transaction (a bug)
// into a trap after ~2 s instead of a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if ((++spins & 0x3FFFu) == 0 && clock64() - t0 > 4000000000ll) __trap();
    }
}
// global -> shared bulk async copy; bytes % 16 == 0, both addresses 16-byte aligned
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void bulk_g2s_hint(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar,
                                              uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
            smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}
// 4-byte asynchronous global->shared copy (LDGSTS)
__device__ __forceinline__ void cp_async_4(void* dst_smem, const void* src_gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst_smem)), "l"(src_gmem) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ uint64_t policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
#endif  // __CUDACC__

}  // namespace otters
