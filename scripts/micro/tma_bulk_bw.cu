// Microbenchmark: HBM read bandwidth of cp.async.bulk (UBLKCP) as a function of the bytes per copy.
// Every warp streams its own sequence of 16-"row" tiles: mode 0 = 16 copies of B bytes each (one per lane),
// mode 1 = one copy of 16*B bytes (same bytes).  Rows of a tile are adjacent in memory in both modes.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}

__global__ void __launch_bounds__(512, 1) bw_kernel(const uint8_t* __restrict__ src, uint64_t total_tiles, uint32_t B, int mode,
                                                     uint32_t slots, unsigned long long* counter, float* sink) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t tile_bytes = 16 * B;
    uint8_t* wbase = smem + (size_t)warp * (slots * tile_bytes + 128);
    uint64_t* bars = (uint64_t*)wbase;
    uint8_t* data = wbase + 128;
    if (lane == 0) {
        for (uint32_t s = 0; s < slots; ++s)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bars[s])), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    float acc = 0.f;
    uint32_t issued = 0, consumed = 0;
    unsigned long long my_tiles[8];
    auto issue = [&]() -> bool {
        unsigned long long t = 0;
        if (lane == 0) t = atomicAdd(counter, 1ull);
        t = __shfl_sync(0xFFFFFFFFu, t, 0);
        if (t >= total_tiles) return false;
        uint32_t slot = issued % slots;
        my_tiles[slot] = t;
        if (lane == 0)
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bars[slot])), "r"(tile_bytes) : "memory");
        __syncwarp();
        const uint8_t* g = src + t * tile_bytes;
        uint8_t* d = data + (size_t)slot * tile_bytes;
        if (mode == 0) {
            if (lane < 16)
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(d + lane * B)),
                             "l"(g + (size_t)lane * B), "r"(B), "r"(smem_u32(&bars[slot])) : "memory");
        } else {
            if (lane == 0)
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(d)),
                             "l"(g), "r"(tile_bytes), "r"(smem_u32(&bars[slot])) : "memory");
        }
        ++issued;
        return true;
    };
    for (uint32_t s = 0; s < slots; ++s) if (!issue()) break;
    while (consumed < issued) {
        uint32_t slot = consumed % slots;
        while (!try_wait(&bars[slot], (consumed / slots) & 1u)) {}
        const float4* p = (const float4*)(data + (size_t)slot * tile_bytes);
        for (uint32_t i = lane; i < tile_bytes / 16; i += 32) { float4 v = p[i]; acc += v.x + v.y + v.z + v.w; }
        __syncwarp();
        ++consumed;
        issue();
    }
    if (acc == 123.456f) sink[0] = acc;
}

int main(int argc, char** argv) {
    const size_t bytes = (size_t)8 << 30;  // 8 GiB source
    uint8_t* src; cudaMalloc(&src, bytes); cudaMemset(src, 1, bytes);
    unsigned long long* counter; cudaMalloc(&counter, 8);
    float* sink; cudaMalloc(&sink, 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaFuncSetAttribute(bw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    printf("mode B_bytes warps slots smemKB  ms  GB/s\n");
    for (int mode = 0; mode < 2; ++mode)
        for (uint32_t B : {256u, 512u, 1024u, 1536u, 2048u, 3072u, 4096u, 8192u})
            for (uint32_t warps : {4u, 8u, 13u, 16u})
                for (uint32_t slots : {1u, 2u}) {
                    size_t smem = (size_t)warps * (slots * 16 * B + 128);
                    if (smem > 226 * 1024) continue;
                    uint64_t tiles = bytes / (16 * B);
                    float best = 1e9;
                    for (int rep = 0; rep < 3; ++rep) {
                        cudaMemset(counter, 0, 8);
                        cudaEventRecord(e0);
                        bw_kernel<<<148, warps * 32, smem>>>(src, tiles, B, mode, slots, counter, sink);
                        cudaEventRecord(e1); cudaEventSynchronize(e1);
                        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
                    }
                    cudaError_t err = cudaGetLastError();
                    if (err != cudaSuccess) { printf("err %s\n", cudaGetErrorString(err)); return 1; }
                    printf("%d %5u %2u %u %4zu %8.3f %8.0f\n", mode, B, warps, slots, smem / 1024, best, bytes / (best * 1e-3) / 1e9);
                    fflush(stdout);
                }
    return 0;
}
