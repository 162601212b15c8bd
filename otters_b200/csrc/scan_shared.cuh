// scan_shared.cuh — pieces shared by the two front-ends of K1 (scan_kernel.cuh: autonomous warps; scan_planner.cu: planner
// warps feeding worker warps): the per-CTA lock-free candidate buffer and its compaction.
#pragma once
#include "internal.h"

namespace otters {
namespace scan_detail {

constexpr unsigned FULL = 0xFFFFFFFFu;

struct CtaHdr {
    unsigned long long tau;  // candidates must have key > tau
    uint32_t count;          // slots reserved in the candidate buffer (may transiently exceed cap)
    uint32_t written;        // slots whose key has been stored
};

__device__ __forceinline__ uint64_t ld_volatile_u64(const unsigned long long* p) {
    return *reinterpret_cast<const volatile unsigned long long*>(p);
}
__device__ __forceinline__ uint32_t ld_volatile_u32(const uint32_t* p) { return *reinterpret_cast<const volatile uint32_t*>(p); }

// Staged rows are fp32, or — stores created with OTTERS_VECTORS_FMT_BF16 — bf16 widened to fp32 on the way into the same
// arithmetic (a bf16 value IS the fp32 value with sixteen zero bits appended, so the widening is exact).
// load_row4: elements [4*idx4, 4*idx4 + 4) of a staged row; load_row1: element e.
template <bool HALF>
__device__ __forceinline__ float4 load_row4(const float* vrow, uint32_t idx4) {
    if constexpr (HALF) {
        const uint2 w = reinterpret_cast<const uint2*>(vrow)[idx4];
        return make_float4(__uint_as_float(w.x << 16), __uint_as_float(w.x & 0xFFFF0000u), __uint_as_float(w.y << 16),
                           __uint_as_float(w.y & 0xFFFF0000u));
    } else {
        return reinterpret_cast<const float4*>(vrow)[idx4];
    }
}
template <bool HALF>
__device__ __forceinline__ float load_row1(const float* vrow, uint32_t e) {
    if constexpr (HALF) return __uint_as_float((uint32_t)reinterpret_cast<const uint16_t*>(vrow)[e] << 16);
    else return vrow[e];
}
// address of column c0 of stored row `row` (pitch_g counts elements)
template <bool HALF>
__device__ __forceinline__ const void* row_src(const float* vectors, uint64_t pitch_g, uint32_t row, uint32_t c0) {
    if constexpr (HALF) return reinterpret_cast<const uint16_t*>(vectors) + (size_t)row * pitch_g + c0;
    else return vectors + (size_t)row * pitch_g + c0;
}

// one warp: sort buf[0..cap) best-first (entries >= cnt are zeroed first); the best min(cnt,k) end up in front
__device__ inline void warp_sort(uint64_t* buf, uint32_t cnt, uint32_t cap, int lane) {
    for (uint32_t i = cnt + lane; i < cap; i += 32) buf[i] = 0ull;
    __syncwarp();
    for (uint32_t size = 2; size <= cap; size <<= 1) {
        for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
            for (uint32_t t = lane; t < (cap >> 1); t += 32) {
                uint32_t lo = 2 * t - (t & (stride - 1));
                uint32_t hi = lo + stride;
                bool desc = (lo & size) == 0;
                uint64_t a = buf[lo], b = buf[hi];
                if ((a < b) == desc) {
                    buf[lo] = b;
                    buf[hi] = a;
                }
            }
            __syncwarp();
        }
    }
}

// Lock-free append of this warp's passing candidates to the CTA buffer.  A warp reserves slots with one
// shared-memory atomicAdd, stores its keys and bumps `written`.  The single warp whose reservation crosses
// the capacity becomes the compactor: it waits until every earlier reservation has been written, sorts,
// keeps the best k, raises the threshold and reopens the buffer; later arrivals wait for the reopen and
// retry against the new threshold.
// g_tau (nullable): grid-wide threshold.  A CTA's k-th best key is a lower bound of the global k-th best key, so it is
// published after every compaction and the largest published value is adopted: all CTAs tighten together instead of
// each warming up its own threshold (which costs a small store several compactions per CTA).
__device__ inline void warp_push(CtaHdr* hdr, uint64_t* buf, uint32_t cap, uint32_t k, bool has, uint64_t key, int lane,
                                 unsigned long long* g_tau = nullptr) {
    for (;;) {
        const uint64_t tau = ld_volatile_u64(&hdr->tau);
        has = has && key > tau;
        const unsigned m = __ballot_sync(FULL, has);
        const uint32_t n = __popc(m);
        if (!n) return;
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(&hdr->count, n);
        base = __shfl_sync(FULL, base, 0);
        if (base + n <= cap) {
            if (has) buf[base + __popc(m & ((1u << lane) - 1u))] = key;
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
            warp_sort(buf, base, cap, lane);
            if (lane == 0) {
                unsigned long long t = buf[k - 1];  // base > cap - 32 >= k
                if (g_tau) {
                    const unsigned long long g = atomicMax(g_tau, t);
                    if (g > t) t = g;
                }
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


// Unit ids claimed ahead of time.  The atomic on the (single, hot) unit counter takes 1-3 us under load, and a run of dead
// units is worked off faster than one claim returns: with a single id in flight the shuffle that broadcasts it was ~10 % of
// the stall samples on the filtered 10M x 128 workload (profiles/r2_scan_c3_hot_sass.txt).  Up to four ids are kept in flight
// in four registers used round-robin (a uniform switch: no instruction touches a register whose atomic is still pending);
// the depth is chosen on the host so that small stores do not starve late warps (ScanParams::claim_depth).
struct UnitClaims {
    uint32_t u0 = 0, u1 = 0, u2 = 0, u3 = 0, phase = 0;
    __device__ __forceinline__ void prime(uint32_t* counter, uint32_t depth, int lane) {
        if (lane != 0) return;
        u0 = atomicAdd(counter, 1u);
        if (depth > 1) u1 = atomicAdd(counter, 1u);
        if (depth > 2) u2 = atomicAdd(counter, 1u);
        if (depth > 3) u3 = atomicAdd(counter, 1u);
    }
    // the oldest claimed id, broadcast to the warp
    __device__ __forceinline__ uint32_t front() const {
        switch (phase) {
        case 0: return __shfl_sync(FULL, u0, 0);
        case 1: return __shfl_sync(FULL, u1, 0);
        case 2: return __shfl_sync(FULL, u2, 0);
        default: return __shfl_sync(FULL, u3, 0);
        }
    }
    // lane 0: replace the id just taken by a new claim
    __device__ __forceinline__ void refill(uint32_t* counter) {
        switch (phase) {
        case 0: u0 = atomicAdd(counter, 1u); break;
        case 1: u1 = atomicAdd(counter, 1u); break;
        case 2: u2 = atomicAdd(counter, 1u); break;
        default: u3 = atomicAdd(counter, 1u); break;
        }
    }
    __device__ __forceinline__ void advance(uint32_t depth) { phase = phase + 1u >= depth ? 0u : phase + 1u; }
};

}  // namespace scan_detail
}  // namespace otters
